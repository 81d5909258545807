"""2-D synthetic flow datasets on fully connected graphs (grevnet_synthetic_data.py of the
reference), producing packed GraphsTuples for the hot path."""
from __future__ import annotations

import random
from functools import partial

import numpy as np

from .graphs import GraphsTuple, data_dicts_to_graphs_tuple

MAX_SEED = 2**32 - 1
GAUSSIAN_MEAN = [0, 0]
GAUSSIAN_COV = [[1, 0], [0, 1]]


def fully_connected_edges(num_nodes):
    """Edge list of fully_connected_nx_graph (grevnet_synthetic_data.py:17-21): complete digraph
    in permutation order, the self-loop appended LAST per sender."""
    n = int(num_nodes)
    i = np.repeat(np.arange(n, dtype=np.int32), n)
    j = np.tile(np.arange(n, dtype=np.int32), n)
    # per sender i: receivers 0..n-1 without i, then i
    keep = i != j
    recv = np.concatenate([j[keep].reshape(n, n - 1), np.arange(n, dtype=np.int32)[:, None]], axis=1) if n else j
    return i, recv.reshape(-1).astype(np.int32)


class SyntheticDataset:
    def __init__(self, graph_generator_fn):
        self.graph_generator_fn = graph_generator_fn

    def get_next_batch_data_dicts(self, batch_size):
        data_dicts = []
        for _ in range(batch_size):        # grevnet_synthetic_data.py:29-43
            (senders, receivers), node_features = self.graph_generator_fn()
            data_dicts.append({
                "n_node": node_features.shape[0], "n_edge": len(senders), "senders": senders,
                "receivers": receivers, "nodes": node_features, "globals": 0,
                "edges": np.zeros(len(senders)),
            })
        return data_dicts

    def get_next_batch(self, batch_size) -> GraphsTuple:
        return data_dicts_to_graphs_tuple(self.get_next_batch_data_dicts(batch_size))


def _make_moons(n_samples, noise, seed):
    from sklearn import datasets
    return datasets.make_moons(n_samples=n_samples, shuffle=True, noise=noise, random_state=seed)[0]


def moons_sample(n_samples, noise=0.05):
    return fully_connected_edges(n_samples), _make_moons(n_samples, noise, random.randrange(MAX_SEED)).astype(np.float32)


def mom_sample(n_samples_choices, noise=0.05):
    return moons_sample(int(np.random.choice(n_samples_choices)), noise=noise)


def mog_sample(offsets_choices, rotate=False):
    offsets = random.choice(offsets_choices)
    num_nodes = len(offsets)
    np.random.shuffle(offsets)
    features = np.random.multivariate_normal(GAUSSIAN_MEAN, GAUSSIAN_COV, num_nodes).astype(np.float32) + offsets
    if rotate:
        angle = np.random.random() * np.pi
        rot = [[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]]
        features = np.transpose(np.matmul(rot, np.transpose(features))).astype(np.float32)
    return fully_connected_edges(num_nodes), features


OFFSETS_4 = np.array([[-5, 5], [5, 5], [-5, -5], [5, -5]]).astype(np.float32)
OFFSETS_6 = np.array([[-5, 5], [5, 5], [-5, -5], [5, -5], [15, 5], [15, -5]]).astype(np.float32)
OFFSETS_9 = np.array([[-5, 5], [5, 5], [-5, -5], [5, -5], [15, 5], [15, -5], [-5, 15], [5, 15],
                      [15, 15]]).astype(np.float32)

DATASETS_MAP = {
    "moons_100": SyntheticDataset(partial(moons_sample, n_samples=100)),
    "moons_10": SyntheticDataset(partial(moons_sample, n_samples=10)),
    "moons_6": SyntheticDataset(partial(moons_sample, n_samples=6)),
    "mom_6_10": SyntheticDataset(partial(mom_sample, n_samples_choices=[6, 10])),
    "mom_6_10_20": SyntheticDataset(partial(mom_sample, n_samples_choices=[6, 10, 20])),
    "mog_4": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_4])),
    "mog_6": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_6])),
    "mog_9": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_9])),
    "mog_4_rotate": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_4], rotate=True)),
    "mog_4_6": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_4, OFFSETS_6])),
    "mog_4_9": SyntheticDataset(partial(mog_sample, offsets_choices=[OFFSETS_4, OFFSETS_9])),
}
