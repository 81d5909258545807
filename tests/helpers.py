"""Shared test helpers: random packed batches, oracle <-> product parameter mapping."""
import os
import sys
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import gnf_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_batch(rng, n_graphs, n_lo, n_hi, p_edge=0.3, D=14, self_loops=True, isolated=False):
    """Random directed graphs, sender-major edge order, optional self-loop first (the layout
    convert_nx_repr produces).  Returns an oracle GraphsTuple of numpy arrays."""
    n_node, senders, receivers = [], [], []
    off = 0
    for _ in range(n_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        adj = rng.random((n, n)) < p_edge
        np.fill_diagonal(adj, False)
        adj = adj | adj.T
        if isolated and n > 2:
            adj[:, n - 1] = False      # node n-1 receives nothing (and gets no self-loop below)
        for i in range(n):
            if self_loops and not (isolated and i == n - 1):
                senders.append(off + i)
                receivers.append(off + i)
            for j in np.nonzero(adj[i])[0]:
                senders.append(off + i)
                receivers.append(off + int(j))
        n_node.append(n)
        off += n
    senders = np.array(senders, dtype=np.int32)
    receivers = np.array(receivers, dtype=np.int32)
    n_node = np.array(n_node, dtype=np.int32)
    # per-graph edge counts
    gid = np.searchsorted(np.cumsum(n_node), senders, side="right")
    n_edge = np.bincount(gid, minlength=n_graphs).astype(np.int32)
    nodes = rng.standard_normal((off, D)).astype(np.float32)
    return O.GraphsTuple(nodes=nodes, edges=None, receivers=receivers, senders=senders, globals=None,
                         n_node=n_node, n_edge=n_edge)


def flat_from_oracle(params):
    """Oracle params -> flat float32 vector in the include/gnf_b200.h order
    (which -> half -> step; per MLP W0 b0 W1 b1 ...)."""
    chunks = []
    for which in ("s", "t"):
        for half in range(2):
            mlps = [params[which][half]] if params["weight_sharing"] else params[which][half]
            for mlp in mlps:
                gnn = mlp
                if isinstance(mlp, dict):      # dm_attn: Wq Wk Wv Wo, the MLP, then LayerNorm gamma beta (include/gnf_b200.h)
                    for k in ("wq", "wk", "wv", "wo"):
                        chunks.append(np.asarray(mlp[k], np.float32).reshape(-1))
                    mlp = mlp["mlp"]
                for (w, b) in mlp:
                    chunks.append(np.asarray(w, np.float32).reshape(-1))
                    chunks.append(np.asarray(b, np.float32).reshape(-1))
                if isinstance(gnn, dict) and "ln_gamma" in gnn:
                    chunks.append(np.asarray(gnn["ln_gamma"], np.float32).reshape(-1))
                    chunks.append(np.asarray(gnn["ln_beta"], np.float32).reshape(-1))
    return np.concatenate(chunks)


def make_grevnet(params, latent_dim, num_layers, device="cuda", math=None):
    """Product GRevNet carrying exactly the oracle's weights."""
    import torch
    import graph_normalizing_flows_b200 as G
    cfg = params["cfg"]
    D, T = params["D"], params["T"]
    mlp_fn = partial(G.make_mlp_model, latent_dim, D / 2, num_layers, cfg["act"], 0.1, 0.1)
    if cfg["block"] == "dm_attn":
        mk = lambda: G.dm_self_attn_gnn(cfg["kq_dim"], cfg["v_dim"], mlp_fn, cfg["num_heads"], cfg["out_dim"],
                                        concat=cfg["concat"], residual=cfg["residual"],
                                        layer_norm=bool(cfg.get("layer_norm", False)),
                                        kq_dim_division=cfg["kq_dim_division"])
    elif cfg["block"] == "concat":
        fac = G.sum_concat_then_mlp_gnn if cfg["agg"] == "sum" else G.avg_concat_then_mlp_gnn
        mk = lambda: fac(mlp_fn)
    else:
        fac = G.sum_then_mlp_gnn if cfg["agg"] == "sum" else G.avg_then_mlp_gnn
        mk = lambda: fac(mlp_fn, cfg["eps"])
    net = G.GRevNet(mk, T, D, use_batch_norm=False, weight_sharing=params["weight_sharing"], math=math,
                    device=device)
    flat = torch.from_numpy(flat_from_oracle(params))
    assert flat.numel() == net.params.numel(), (flat.numel(), net.params.numel())
    with torch.no_grad():
        net.params.copy_(flat.to(net.params.device))
    return net


def to_device_graph(graph, device="cuda"):
    import graph_normalizing_flows_b200 as G
    return G.GraphsTuple(*graph).to(device)


def rel_err(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-30)
