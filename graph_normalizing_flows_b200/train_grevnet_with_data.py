"""Input side of the reference's embedding-flow trainer (train_grevnet_with_data.py): the on-disk
training format and the fully-connected batch construction, plus the sampling -> decode tail, on
top of the B200 hot path.  SURVEY §8 rows f3 / f4.

On-disk format (written by generate_grevnet_training_data.py:90-92,115-119): every file is a
pickle of `(node_embeddings [sum(n_node), D] float32, n_node [G] int)`.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import loss as loss_lib
from .graphs import GraphsTuple
from .utils import senders_receivers


class _EmbeddingFile:
    """One pickle `(node_embeddings, n_node)` with a node-offset table: graphs [g0, g1) are rows
    [node_off[g0], node_off[g1]) of the embedding matrix."""

    def __init__(self, path):
        with open(path, "rb") as f:
            d = pickle.load(f)
        self.node_embeddings = np.asarray(d[0], dtype=np.float32)
        self.n_node = np.asarray(d[1], dtype=np.int64)
        self.node_off = np.concatenate([[0], np.cumsum(self.n_node)])

    def graphs(self, g0, g1):
        return self.node_embeddings[self.node_off[g0]:self.node_off[g1]], self.n_node[g0:g1]


class _TableReader:
    """Batches of consecutive graphs, file after file.  Every file is cut ONCE into a table of graph ranges
    (`_cut`); `train_batch` walks the table and moves to the next file when it is used up.  Running past the last
    file raises IndexError, as indexing the reference's file list does."""

    def __init__(self, train_data_dir, files):
        self.train_data_dir, self.files = train_data_dir, files
        self.file_ind = -1
        self._next_file()

    def _next_file(self):
        self.file_ind += 1
        self._file = _EmbeddingFile(os.path.join(self.train_data_dir, self.files[self.file_ind]))
        self._table = self._cut(self._file.n_node)
        self._row = 0

    def train_batch(self):
        if self._row >= len(self._table):
            self._next_file()
        g0, g1 = self._table[self._row]
        self._row += 1
        return self._file.graphs(g0, g1)


class GrevnetDatasetFixed(_TableReader):
    """train_grevnet_with_data.py:145-182: `train_batch_size` graphs per batch, the file list repeated
    `train_epochs` times; the graphs of a file that do not fill a batch are skipped (the reference opens the
    next file when `prev_graph_ind + train_batch_size > len(n_node)`)."""

    def __init__(self, train_data_dir, train_batch_size, train_epochs=1):
        self.train_batch_size = int(train_batch_size)
        super().__init__(train_data_dir, sorted(os.listdir(train_data_dir)) * train_epochs)

    def _cut(self, n_node):
        b = self.train_batch_size
        if len(n_node) < b:          # the reference indexes n_node_cs[b - 1] of the freshly opened file
            raise IndexError(f"file {self.files[self.file_ind]} holds {len(n_node)} graphs, fewer than one batch of {b}")
        return [(k * b, (k + 1) * b) for k in range(len(n_node) // b)]


class GrevnetDatasetVariable(_TableReader):
    """train_grevnet_with_data.py:185-234: as many consecutive graphs as keep the node count BELOW `max_nodes`
    (strict: a graph is added while total + n < max_nodes); the tail of a file is flushed as its own batch.
    (A graph with >= max_nodes nodes makes the reference return empty batches forever; here it is an error.)"""

    def __init__(self, train_data_dir, max_nodes):
        self.max_nodes = int(max_nodes)
        super().__init__(train_data_dir, sorted(os.listdir(train_data_dir)))

    def _cut(self, n_node):
        table, g, count = [], 0, len(n_node)
        while g < count:
            g0, total = g, 0
            while g < count and total + n_node[g] < self.max_nodes:
                total += n_node[g]
                g += 1
            if g == g0:
                raise ValueError(f"graph {g0} of {self.files[self.file_ind]} has {int(n_node[g0])} >= max_nodes nodes")
            table.append((g0, g))
        return table


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
