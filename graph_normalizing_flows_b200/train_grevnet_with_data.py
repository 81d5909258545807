"""Input side of the reference's embedding-flow trainer (train_grevnet_with_data.py): the on-disk
training format and the fully-connected batch construction, plus the sampling -> decode tail, on
top of the B200 hot path.  SURVEY §8 rows f3 / f4.

On-disk format (written by generate_grevnet_training_data.py:90-92,115-119): every file is a
pickle of `(node_embeddings [sum(n_node), D] float32, n_node [G] int)`.
"""
from __future__ import annotations

import os
import pickle
import random

import numpy as np
import torch

from . import loss as loss_lib
from .graphs import GraphsTuple
from .utils import senders_receivers


def _load(path):
    with open(path, "rb") as f:
        d = pickle.load(f)
    return np.asarray(d[0], dtype=np.float32), np.asarray(d[1], dtype=np.int64)


class GrevnetDatasetFixed:
    """train_grevnet_with_data.py:145-182: `train_batch_size` graphs per batch, files in os.listdir
    order repeated `train_epochs` times, next file when the current one cannot fill a batch."""

    def __init__(self, train_data_dir, train_batch_size, train_epochs=1):
        self.files = sorted(os.listdir(train_data_dir)) * train_epochs
        self.file_ind = 0
        self.prev_graph_ind = 0
        self.prev_node_embedding_ind = 0
        self.train_batch_size = train_batch_size
        self.train_data_dir = train_data_dir
        self._open()

    def _open(self):
        self.node_embeddings, self.n_node = _load(os.path.join(self.train_data_dir, self.files[self.file_ind]))
        self.n_node_cs = np.cumsum(self.n_node)
        self.prev_graph_ind = 0
        self.prev_node_embedding_ind = 0

    def train_batch(self):
        new_ind = self.prev_graph_ind + self.train_batch_size
        if new_ind > len(self.n_node):
            self.file_ind += 1
            self._open()
            new_ind = self.train_batch_size
        node_embeddings = self.node_embeddings[self.prev_node_embedding_ind:self.n_node_cs[new_ind - 1]]
        n_node = self.n_node[self.prev_graph_ind:new_ind]
        self.prev_graph_ind = new_ind
        self.prev_node_embedding_ind = self.n_node_cs[new_ind - 1]
        return node_embeddings, n_node


class GrevnetDatasetVariable:
    """train_grevnet_with_data.py:185-234: as many consecutive graphs as stay below `max_nodes`."""

    def __init__(self, train_data_dir, max_nodes):
        self.files = sorted(os.listdir(train_data_dir))
        self.file_ind = 0
        self.max_nodes = max_nodes
        self.train_data_dir = train_data_dir
        self._open()

    def _open(self):
        self.node_embeddings, self.n_node = _load(os.path.join(self.train_data_dir, self.files[self.file_ind]))
        self.n_node_cs = np.cumsum(self.n_node)
        self.graph_ind = 0
        self.prev_graph_ind = 0
        self.prev_node_embedding_ind = 0

    def train_batch(self):
        total_nodes = 0
        while True:
            if self.graph_ind >= len(self.n_node):       # file exhausted: flush the tail, open the next
                node_embeddings = self.node_embeddings[self.prev_node_embedding_ind:self.n_node_cs[self.graph_ind - 1]]
                n_node = self.n_node[self.prev_graph_ind:self.graph_ind]
                self.file_ind += 1
                self._open()
                return node_embeddings, n_node
            if total_nodes + self.n_node[self.graph_ind] < self.max_nodes:
                total_nodes += self.n_node[self.graph_ind]
                self.graph_ind += 1
            else:
                break
        node_embeddings = self.node_embeddings[self.prev_node_embedding_ind:self.n_node_cs[self.graph_ind - 1]]
        n_node = self.n_node[self.prev_graph_ind:self.graph_ind]
        self.prev_graph_ind = self.graph_ind
        self.prev_node_embedding_ind = self.n_node_cs[self.graph_ind - 1]
        return node_embeddings, n_node


def transform_example(node_embeddings, n_node) -> GraphsTuple:
    """train_grevnet_with_data.py:237-271: fully connected graphs from the sizes (all n^2 ordered
    pairs incl. self pairs, sender-major; utils.py:164-183), n_edge = n_node^2."""
    n_node = np.asarray(n_node, dtype=np.int32)
    senders, receivers = senders_receivers(n_node)
    return GraphsTuple(nodes=np.ascontiguousarray(node_embeddings, dtype=np.float32), edges=None,
                       receivers=receivers, senders=senders, globals=None, n_node=n_node,
                       n_edge=(n_node.astype(np.int64) ** 2).astype(np.int32))


def sample_graphs(grevnet, sample_n_node, device="cuda", generator=None):
    """train_grevnet_with_data.py:397-416,529-548: z ~ N(0, I) on FC graphs of the requested sizes ->
    x = grevnet(z, inverse=False) -> pred_adj(scaled_hacky_sigmoid_l2) -> threshold 0.5 -> networkx.
    Returns (graphs, per-graph mean prior log-prob of the sampled nodes)."""
    sample_n_node = np.asarray(sample_n_node, dtype=np.int32)
    d = grevnet.node_embedding_dim
    z = torch.randn(int(sample_n_node.sum()), d, generator=generator)
    g = transform_example(z.numpy(), sample_n_node).to(device)
    top = grevnet(g, inverse=False)
    blocks, adj_off = loss_lib.pred_adj(top)
    log_prob = -0.5 * (z * z).sum(1) - 0.5 * d * loss_lib.LOG_2PI           # mvn.log_prob(sample_nodes)
    cs = np.concatenate([[0], np.cumsum(sample_n_node)])
    per_graph = [float(log_prob[cs[i]:cs[i + 1]].mean()) for i in range(len(sample_n_node))]
    return loss_lib.sampled_graphs(blocks, adj_off, sample_n_node), per_graph
