"""Index helpers of the reference's utils.py that sit on the hot path's input side."""
from __future__ import annotations

import numpy as np


def senders_receivers(n_node):
    """utils.py:164-183 (with cartesian_graph :133-150 and permutations :153-161): for every graph
    all n^2 ordered node pairs, self-pairs included, sender-major:
        edge k = i*n + j  ->  sender = lo + i, receiver = lo + j,   n_edge = n_node**2.
    Vectorised; int32 like the TensorArray(dtype=tf.int32) of the reference."""
    n_node = np.asarray(n_node, dtype=np.int64).reshape(-1)
    n_edge = n_node * n_node
    total = int(n_edge.sum())
    if total == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    lo = np.concatenate([[0], np.cumsum(n_node)[:-1]])
    e_lo = np.concatenate([[0], np.cumsum(n_edge)[:-1]])
    k = np.arange(total, dtype=np.int64) - np.repeat(e_lo, n_edge)
    n_rep = np.repeat(n_node, n_edge)
    lo_rep = np.repeat(lo, n_edge)
    senders = lo_rep + k // n_rep
    receivers = lo_rep + k % n_rep
    return senders.astype(np.int32), receivers.astype(np.int32)
