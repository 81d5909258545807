"""Batch construction for real graph families (input contract of the hot path).

Mirrors /root/reference/graph_data.py:24-122,283-300.  The pickles are read once and converted
to packed integer structures; every batch is then assembled with vectorised numpy (no per-call
networkx work), bit-identical in layout to `gn.utils_np.networkxs_to_graphs_tuple` applied to the
reference's `convert_nx_repr` graphs:  nodes relabelled 0..n-1 in iteration order, one self-loop
per node FIRST, edges sender-major, indices offset by the cumulative node count.
"""
from __future__ import annotations

import os
import pickle
import random
from functools import partial

import numpy as np

from .graphs import GraphsTuple, concat_structures

FILENAME_MAP = {     # graph_data.py:283-300
    "graph_rnn_grid": "training_graphs/GraphRNN_RNN_grid_4_128_train_0.dat",
    "graph_rnn_protein": "training_graphs/GraphRNN_RNN_protein_4_128_train_0.dat",
    "graph_rnn_ego": "training_graphs/GraphRNN_RNN_citeseer_4_128_train_0.dat",
    "graph_rnn_community": "training_graphs/GraphRNN_RNN_caveman_4_128_train_0.dat",
    "graph_rnn_ego_small": "training_graphs/GraphRNN_RNN_citeseer_small_4_64_train_0.dat",
    "graph_rnn_community_small": "training_graphs/GraphRNN_RNN_caveman_small_4_64_train_0.dat",
    "graph_rnn_community_medium": "training_graphs/GraphRNN_RNN_community_medium_4_128_train_0.dat",
    "graph_rnn_ego_medium": "training_graphs/GraphRNN_RNN_ego_medium_4_128_train_0.dat",
}


def gaussian_noise_features(num_nodes, num_components=5, scale=1.0, rng=None):
    """add_gaussian_noise_features (graph_data.py:24-27): N(0, scale^2) float32 per node."""
    rng = np.random if rng is None else rng
    return rng.normal(scale=scale, size=(num_nodes, num_components)).astype(np.float32)


def convert_nx_repr(graph):
    """graph_data.py:33-50 without building a DiGraph: returns (n, senders, receivers) in the
    edge order networkx would iterate the converted graph -- per sender: the self-loop first,
    then its out-neighbours in the order the directed edges were added (duplicates of an
    existing (u, v) -- e.g. a pre-existing self-loop -- keep their first position)."""
    index_map = {}
    for node in graph.nodes():
        index_map[node] = len(index_map)
    n = len(index_map)
    adj = [dict.fromkeys([i]) for i in range(n)]       # insertion-ordered, self-loop first
    for u, v in graph.edges():
        adj[index_map[u]].setdefault(index_map[v])
    senders = np.fromiter((u for u in range(n) for _ in adj[u]), dtype=np.int32)
    receivers = np.fromiter((v for u in range(n) for v in adj[u]), dtype=np.int32)
    return n, senders, receivers


def preprocess_networkx_graphs(graphs):
    return [convert_nx_repr(g) for g in graphs]


def structures_from_fixture(npz):
    """Per-graph (n, senders_local, receivers_local) from a tests/golden/graphs_<family>.npz."""
    n_node, n_edge = npz["n_node"], npz["n_edge"]
    s, r = npz["senders_local"].astype(np.int32), npz["receivers_local"].astype(np.int32)
    off = np.concatenate([[0], np.cumsum(n_edge)])
    return [(int(n_node[i]), s[off[i]:off[i + 1]], r[off[i]:off[i + 1]]) for i in range(len(n_node))]


class GraphDataset:
    """graph_data.py:61-122.  `root` is the directory holding training_graphs/."""

    def __init__(self, dataset_name, node_embedding_dim, gaussian_scale=1.0, root=".", structures=None):
        self.node_embedding_dim = int(node_embedding_dim)
        self.gaussian_scale = float(gaussian_scale)
        if structures is not None:              # pre-converted train split (fixtures)
            self.graphs = None
            self.train_graphs, self.test_graphs = list(structures), []
        else:
            filename = os.path.join(root, FILENAME_MAP[dataset_name])
            with open(filename, "rb") as fh:
                self.graphs = pickle.load(fh)
            self.train_graphs, self.test_graphs = self.process_and_split_graphs(self.graphs)
        self.train_index = 0
        self.test_index = 0

    def process_and_split_graphs(self, graphs):
        n = len(graphs)                         # graph_data.py:76-85
        test = [convert_nx_repr(g.to_directed()) for g in graphs[int(0.8 * n):]]
        train = [convert_nx_repr(g.to_directed()) for g in graphs[0:int(0.8 * n)]]
        return train, test

    def full_n_nodes(self):
        return [g.number_of_nodes() for g in self.graphs]

    def train_n_nodes(self):
        return [g[0] for g in self.train_graphs]

    def test_n_nodes(self):
        return [g[0] for g in self.test_graphs]

    def _batch(self, structs) -> GraphsTuple:
        n = sum(s[0] for s in structs)
        nodes = gaussian_noise_features(n, self.node_embedding_dim, self.gaussian_scale)
        return concat_structures(structs, nodes=nodes)

    def get_next_train_batch(self, batch_size) -> GraphsTuple:
        """graph_data.py:113-122.  The reference never advances train_index (it assigns
        self.index), so every draw reshuffles and takes element 0: sampling with replacement.
        Reproduced as such; Gaussian node features are redrawn per call."""
        batch = []
        for _ in range(batch_size):
            if self.train_index == 0:
                random.shuffle(self.train_graphs)
            batch.append(self.train_graphs[self.train_index])
        return self._batch(batch)

    def get_random_test_batch(self, batch_size) -> GraphsTuple:
        return self._batch(random.choices(self.test_graphs, k=batch_size))

    def draw_batch(self, batch_size, rng) -> GraphsTuple:
        """Benchmark batches: `batch_size` draws with replacement from the train
        split with a numpy Generator, features N(0,1) from the same generator."""
        idx = rng.integers(0, len(self.train_graphs), size=batch_size)
        structs = [self.train_graphs[i] for i in idx]
        n = sum(s[0] for s in structs)
        nodes = rng.standard_normal((n, self.node_embedding_dim)).astype(np.float32) * np.float32(self.gaussian_scale)
        return concat_structures(structs, nodes=nodes)
