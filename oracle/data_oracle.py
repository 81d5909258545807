"""CPU oracle for the batch-construction contract  --  TEST
INFRASTRUCTURE, NOT PRODUCT CODE.  Plain python/networkx loops that follow the reference
line by line; the product's vectorised batching is checked bit-exact against these.

PARITY UNPINNED (no reference tests / golden vectors exist); graph_nets' utils_np
is restated from its published behaviour [upstream].
"""
from __future__ import annotations

import networkx as nx
import numpy as np

from .gnf_oracle import GraphsTuple


def convert_nx_repr(graph):
    """graph_data.py:33-50: relabel to 0..n-1 in node-iteration order, add a self-loop to
    every node FIRST, then the (directed) edges.  Features are attached by the caller."""
    new_graph = nx.DiGraph(features=0)
    index_map = {}
    new_ind = 0
    for node in graph.nodes(data=True):
        index_map[node[0]] = new_ind
        new_graph.add_node(new_ind)
        new_graph.add_edge(new_ind, new_ind, features=0)
        new_ind += 1
    for edge in graph.edges(data=True):
        new_graph.add_edge(index_map[edge[0]], index_map[edge[1]], features=0)
    return new_graph


def train_split(graphs):
    """graph_data.py:76-84: first int(0.8*len) graphs, to_directed(), convert_nx_repr."""
    n = len(graphs)
    return [convert_nx_repr(g.to_directed()) for g in graphs[0:int(0.8 * n)]]


def nx_to_arrays(g):
    """graph_nets.utils_np.networkx_to_data_dict [upstream]: nodes must be keyed 0..n-1 in
    iteration order; senders/receivers = zip(*g.edges()) in networkx iteration order."""
    for i, key in enumerate(g.nodes()):
        if key != i:
            raise ValueError("nodes must be sequentially numbered")
    n_edge = g.number_of_edges()
    if n_edge:
        s, r = zip(*g.edges())
    else:
        s, r = (), ()
    return (g.number_of_nodes(), np.array(s, dtype=np.int32), np.array(r, dtype=np.int32))


def graphs_tuple_from_arrays(per_graph, nodes):
    """graph_nets.utils_np.data_dicts_to_graphs_tuple [upstream]: concatenate and offset the
    indices by the cumulative n_node of the preceding graphs."""
    n_node = np.array([p[0] for p in per_graph], dtype=np.int32)
    n_edge = np.array([len(p[1]) for p in per_graph], dtype=np.int32)
    senders, receivers = [], []
    off = 0
    for n, s, r in per_graph:
        senders.append(s + np.int32(off))
        receivers.append(r + np.int32(off))
        off += n
    cat = lambda xs: np.concatenate(xs).astype(np.int32) if xs else np.zeros(0, np.int32)
    return GraphsTuple(nodes=nodes, edges=None, receivers=cat(receivers), senders=cat(senders),
                       globals=None, n_node=n_node, n_edge=n_edge)


def networkxs_to_graphs_tuple(graphs, nodes):
    """graph_data.py:107,110,122: gn.utils_np.networkxs_to_graphs_tuple(batch)."""
    return graphs_tuple_from_arrays([nx_to_arrays(g) for g in graphs], nodes)


def fully_connected_nx_graph(num_nodes):
    """grevnet_synthetic_data.py:17-21."""
    g = nx.complete_graph(num_nodes, create_using=nx.DiGraph)
    r = range(num_nodes)
    g.add_edges_from(zip(r, r))
    return g


def senders_receivers(n_node):
    """utils.py:164-183 (+ cartesian_graph :133-150, permutations :153-161): per graph all n^2
    ordered pairs, sender-major: edge k = i*n + j -> (lo+i, lo+j)."""
    senders, receivers = [], []
    lo = 0
    for n in n_node:
        n = int(n)
        for i in range(n):
            for j in range(n):
                senders.append(lo + i)
                receivers.append(lo + j)
        lo += n
    return np.array(senders, dtype=np.int32), np.array(receivers, dtype=np.int32)


# --------------------------------------------------------------------------- #
# f4  GrevnetDatasetFixed / GrevnetDatasetVariable  (train_grevnet_with_data.py:145-234)
# State machines restated statement by statement over in-memory "files" [(node_embeddings, n_node), ...];
# they stop (return) where the reference would index past its file list.
# --------------------------------------------------------------------------- #
def dataset_fixed_batches(files, train_batch_size):
    file_ind, prev_graph_ind, prev_node_embedding_ind = 0, 0, 0          # :147-150
    node_embeddings, n_node = files[file_ind]                            # :153-157
    n_node_cs = np.cumsum(n_node)                                        # :158
    while True:
        new_ind = prev_graph_ind + train_batch_size                      # :161
        if new_ind > len(n_node):                                        # :162
            file_ind += 1                                                # :163
            if file_ind >= len(files):
                return
            node_embeddings, n_node = files[file_ind]                    # :170-172
            prev_graph_ind, prev_node_embedding_ind = 0, 0               # :173-174
            n_node_cs = np.cumsum(n_node)                                # :175
            new_ind = prev_graph_ind + train_batch_size                  # :176
        yield (node_embeddings[prev_node_embedding_ind:n_node_cs[new_ind - 1]],   # :177-178
               n_node[prev_graph_ind:new_ind])                                     # :179
        prev_graph_ind = new_ind                                         # :180
        prev_node_embedding_ind = n_node_cs[new_ind - 1]                 # :181


def dataset_variable_batches(files, max_nodes):
    file_ind, graph_ind, prev_graph_ind, prev_node_embedding_ind = 0, 0, 0, 0    # :188-191
    node_embeddings, n_node = files[file_ind]
    n_node_cs = np.cumsum(n_node)
    while True:
        total_nodes = 0                                                  # :203
        flushed = False
        while True:                                                      # :204
            if graph_ind >= len(n_node):                                 # :205
                out = (node_embeddings[prev_node_embedding_ind:n_node_cs[graph_ind - 1]],   # :206-208
                       n_node[prev_graph_ind:graph_ind])                                       # :209
                file_ind += 1                                            # :210
                prev_graph_ind, graph_ind, prev_node_embedding_ind = 0, 0, 0   # :211-213
                yield out
                if file_ind >= len(files):
                    return
                node_embeddings, n_node = files[file_ind]                # :218-221
                n_node_cs = np.cumsum(n_node)                            # :222
                flushed = True                                           # :223 return
                break
            if total_nodes + n_node[graph_ind] < max_nodes:              # :224
                total_nodes += n_node[graph_ind]                         # :225
                graph_ind += 1                                           # :226
            else:
                break                                                    # :228
        if flushed:
            continue
        yield (node_embeddings[prev_node_embedding_ind:n_node_cs[graph_ind - 1]],    # :229-231
               n_node[prev_graph_ind:graph_ind])                                       # :232
        prev_graph_ind = graph_ind                                       # :233
        prev_node_embedding_ind = n_node_cs[graph_ind - 1]               # :234
