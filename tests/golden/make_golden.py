"""Generates the committed fixtures under tests/golden/.  Run HERE (the container that has
/root/reference); the GPU box only reads the outputs.

  graphs_<family>.npz   packed index structure of the reference's own training graphs
                        (training_graphs/GraphRNN_RNN_<family>_train_0.dat, first 80 % as
                        graph_data.py:76-78, to_directed + convert_nx_repr as :33-50,:81-84),
                        produced with oracle/data_oracle.py (python/networkx loops).
  golden_<case>.npz     seeded inputs and the numpy oracle's outputs (fp32 in reference op order,
                        and fp64) for small GRevNet configurations.

    python tests/golden/make_golden.py
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import data_oracle as DO  # noqa: E402
from oracle import gnf_oracle as O    # noqa: E402
import helpers as H                   # noqa: E402

REF = "/root/reference/training_graphs"
FAMILIES = {   # file token -> FILENAME_MAP key (graph_data.py:283-300)
    "caveman_small_4_64": "graph_rnn_community_small",
    "community_medium_4_128": "graph_rnn_community_medium",
    "grid_4_128": "graph_rnn_grid",
    "protein_4_128": "graph_rnn_protein",
    "citeseer_4_128": "graph_rnn_ego",
    "caveman_4_128": "graph_rnn_community",
}


def family_fixture(token):
    graphs = pickle.load(open(os.path.join(REF, f"GraphRNN_RNN_{token}_train_0.dat"), "rb"))
    train = DO.train_split(graphs)
    per = [DO.nx_to_arrays(g) for g in train]
    n_node = np.array([p[0] for p in per], np.int32)
    n_edge = np.array([len(p[1]) for p in per], np.int32)
    assert n_node.max() < 65536
    return dict(n_node=n_node, n_edge=n_edge,
                senders_local=np.concatenate([p[1] for p in per]).astype(np.uint16),
                receivers_local=np.concatenate([p[2] for p in per]).astype(np.uint16),
                n_graphs_total=np.int32(len(graphs)))


CASES = {
    # name: (graph source, T, D, L, K, block, agg, eps, act, weight_sharing, last_scale)
    "caveman_small_b8_t2": ("caveman_small_4_64:8", 2, 14, 256, 5, "concat", "sum", 1.0, "leaky_relu", False, 0.05),
    "rand_concat_sum_d4": ("random:6:3:12", 2, 4, 128, 4, "concat", "sum", 1.0, "leaky_relu", False, 0.1),
    "rand_concat_mean_d14": ("random:5:4:20", 1, 14, 128, 3, "concat", "mean", 1.0, "leaky_relu", False, 0.1),
    "rand_aggthen_sum_d2": ("random:4:3:9", 2, 2, 256, 5, "agg_then", "sum", 1.0, "leaky_relu", False, 0.1),
    "rand_aggthen_mean_d14_relu": ("random:5:4:20", 2, 14, 256, 2, "agg_then", "mean", 0.5, "relu", False, 0.1),
    "rand_shared_d14": ("random:5:4:20", 3, 14, 128, 3, "concat", "sum", 1.0, "leaky_relu", True, 0.05),
    "rand_isolated_mean_d6": ("random_iso:5:4:14", 1, 6, 128, 3, "concat", "mean", 1.0, "leaky_relu", False, 0.1),
    "fc_d2_small_mlp": ("fc:3:5:9", 2, 2, 48, 3, "agg_then", "mean", 1.0, "leaky_relu", False, 0.1),
}


def build_graph(src, D, seed):
    rng = np.random.default_rng(seed)
    kind = src.split(":")[0]
    if kind == "random" or kind == "random_iso":
        _, g, lo, hi = src.split(":")
        return H.random_batch(rng, int(g), int(lo), int(hi), D=D, isolated=(kind == "random_iso"))
    if kind == "fc":
        _, g, lo, hi = src.split(":")
        n_node = rng.integers(int(lo), int(hi) + 1, size=int(g))
        s, r = DO.senders_receivers(n_node)
        nodes = rng.standard_normal((int(n_node.sum()), D)).astype(np.float32)
        return O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node * n_node).astype(np.int32))
    token, b = src.split(":")
    fam = np.load(os.path.join(HERE, f"graphs_{token}.npz"))
    b = int(b)
    n_node, n_edge = fam["n_node"][:b], fam["n_edge"][:b]
    e_tot = int(n_edge.sum())
    off = np.repeat(np.concatenate([[0], np.cumsum(n_node)[:-1]]), n_edge)
    s = (fam["senders_local"][:e_tot].astype(np.int64) + off).astype(np.int32)
    r = (fam["receivers_local"][:e_tot].astype(np.int64) + off).astype(np.int32)
    nodes = rng.standard_normal((int(n_node.sum()), D)).astype(np.float32)
    return O.GraphsTuple(nodes, None, r, s, None, n_node.copy(), n_edge.copy())


def golden_case(name, spec, seed=12345):
    src, T, D, L, K, block, agg, eps, act, ws, scale = spec
    g = build_graph(src, D, seed)
    params = O.make_params(seed, T, D, L, K, agg=agg, block=block, eps=eps, act=act,
                           last_layer_scale=scale, weight_sharing=ws)
    z, ldj = O.grevnet_f(g.nodes, g.senders, g.receivers, params)
    lp = O.log_prob(z, ldj, g.n_node)
    x_back = O.grevnet_g(z, g.senders, g.receivers, params)
    p64 = O.cast_params(params, np.float64)
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, p64)
    lp64 = O.log_prob(z64, ldj64, g.n_node)
    agg0 = O.aggregate(g.nodes[:, :D // 2].copy(), g.senders, g.receivers, agg)
    flat = H.flat_from_oracle(params)
    return dict(nodes=g.nodes, senders=g.senders, receivers=g.receivers, n_node=g.n_node, n_edge=g.n_edge,
                spec=np.array([T, D, L, K, int(ws)], np.int32), block=block, agg=agg, act=act,
                eps=np.float32(eps), last_scale=np.float32(scale), seed=np.int64(seed),
                params_checksum=np.float64(flat.astype(np.float64).sum()),
                params_abs_checksum=np.float64(np.abs(flat.astype(np.float64)).sum()),
                z=z, ldj=np.float32(ldj), log_prob_zs=np.float32(lp["log_prob_zs"]),
                log_prob_xs=np.float32(lp["log_prob_xs"]), x_back=x_back, agg_first_half=agg0,
                z64=z64, ldj64=np.float64(ldj64), log_prob_xs64=np.float64(lp64["log_prob_xs"]),
                log_prob_zs64=np.float64(lp64["log_prob_zs"]))


def main():
    for token in FAMILIES:
        out = os.path.join(HERE, f"graphs_{token}.npz")
        fx = family_fixture(token)
        np.savez_compressed(out, **fx)
        print(token, len(fx["n_node"]), int(fx["n_node"].sum()), int(fx["n_edge"].sum()),
              os.path.getsize(out) // 1024, "KiB")
    for name, spec in CASES.items():
        out = os.path.join(HERE, f"golden_{name}.npz")
        np.savez_compressed(out, **golden_case(name, spec))
        print(name, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
