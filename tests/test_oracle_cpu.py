"""CPU suite: the oracle against the committed golden vectors and the known-answer tests
K1-K5.  The reference ships no tests, so these are the pins ("parity unpinned")."""
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import data_oracle as DO
from oracle import gnf_oracle as O

CASES = sorted(glob.glob(os.path.join(H.GOLDEN, "golden_*.npz")))


def load_case(path):
    g = np.load(path, allow_pickle=False)
    T, D, L, K, ws = [int(v) for v in g["spec"]]
    params = O.make_params(int(g["seed"]), T, D, L, K, agg=str(g["agg"]), block=str(g["block"]),
                           eps=float(g["eps"]), act=str(g["act"]), last_layer_scale=float(g["last_scale"]),
                           weight_sharing=bool(ws))
    return g, params, (T, D, L, K)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_oracle_reproduces_golden(path):
    g, params, _ = load_case(path)
    flat = H.flat_from_oracle(params)
    # RNG stream drift would silently change every weight: pin it first
    assert np.isclose(flat.astype(np.float64).sum(), float(g["params_checksum"]), rtol=0, atol=1e-9)
    z, ldj = O.grevnet_f(g["nodes"], g["senders"], g["receivers"], params)
    assert np.allclose(z, g["z"], rtol=1e-5, atol=1e-6)
    assert abs(float(ldj) - float(g["ldj"])) <= 1e-5 * max(1.0, abs(float(g["ldj"])))
    lp = O.log_prob(z, ldj, g["n_node"])
    assert H.rel_err(lp["log_prob_xs"], g["log_prob_xs"]) < 1e-6
    # segment op is integer-indexed fp32 adds in a fixed order: bit-exact
    agg = O.aggregate(g["nodes"][:, :g["nodes"].shape[1] // 2].copy(), g["senders"], g["receivers"], str(g["agg"]))
    assert np.array_equal(agg, g["agg_first_half"])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_fp32_oracle_close_to_fp64(path):
    g, _, _ = load_case(path)
    assert H.rel_err(g["log_prob_xs"], g["log_prob_xs64"]) < 1e-6
    assert np.abs(g["z"] - g["z64"]).max() < 1e-4


def test_k1_round_trip():
    rng = np.random.default_rng(0)
    g = H.random_batch(rng, 6, 4, 15, D=14)
    for block, agg in (("concat", "sum"), ("agg_then", "mean")):
        p = O.make_params(5, 3, 14, 64, 4, agg=agg, block=block, last_layer_scale=0.1)
        z, _ = O.grevnet_f(g.nodes, g.senders, g.receivers, p)
        x = O.grevnet_g(z, g.senders, g.receivers, p)
        assert np.abs(x - g.nodes).max() < 2e-5


def test_k2_zero_last_layer_is_identity():
    rng = np.random.default_rng(1)
    g = H.random_batch(rng, 4, 4, 10, D=6)
    p = O.make_params(5, 2, 6, 32, 3, last_layer_scale=0.0)
    z, ldj = O.grevnet_f(g.nodes, g.senders, g.receivers, p)
    assert np.array_equal(z, g.nodes) and float(ldj) == 0.0
    lp = O.log_prob(z, ldj, g.n_node)
    n, d = g.nodes.shape
    want = -0.5 * float((g.nodes.astype(np.float64) ** 2).sum()) - 0.5 * n * d * np.log(2 * np.pi)
    assert H.rel_err(lp["log_prob_xs"], want) < 1e-6


def test_k3_logdet_is_the_jacobian_logdet():
    """sum(s) == log|det d vec(z)/d vec(x)| (triangular coupling across nodes AND features):
    central finite differences of the fp64 oracle on a tiny graph."""
    rng = np.random.default_rng(2)
    g = H.random_batch(rng, 2, 2, 3, p_edge=0.8, D=4)
    p = O.cast_params(O.make_params(9, 2, 4, 16, 3, last_layer_scale=0.5), np.float64)
    x = g.nodes.astype(np.float64)
    z0, ldj = O.grevnet_f(x, g.senders, g.receivers, p)
    n = x.size
    J = np.zeros((n, n))
    eps = 1e-6
    for i in range(n):
        dx = np.zeros(n)
        dx[i] = eps
        zp, _ = O.grevnet_f(x + dx.reshape(x.shape), g.senders, g.receivers, p)
        zm, _ = O.grevnet_f(x - dx.reshape(x.shape), g.senders, g.receivers, p)
        J[:, i] = ((zp - zm) / (2 * eps)).reshape(-1)
    sign, logdet = np.linalg.slogdet(J)
    assert sign > 0
    assert abs(logdet - float(ldj)) < 1e-6 * max(1.0, abs(float(ldj)))


def test_k4_graph_independence():
    rng = np.random.default_rng(3)
    g = H.random_batch(rng, 5, 4, 12, D=14)
    p = O.make_params(5, 2, 14, 64, 3, last_layer_scale=0.1)
    z, ldj = O.grevnet_f(g.nodes, g.senders, g.receivers, p)
    node_off = np.concatenate([[0], np.cumsum(g.n_node)])
    edge_off = np.concatenate([[0], np.cumsum(g.n_edge)])
    total = 0.0
    for k in range(len(g.n_node)):
        s = g.senders[edge_off[k]:edge_off[k + 1]] - node_off[k]
        r = g.receivers[edge_off[k]:edge_off[k + 1]] - node_off[k]
        zk, lk = O.grevnet_f(g.nodes[node_off[k]:node_off[k + 1]], s, r, p)
        assert np.allclose(zk, z[node_off[k]:node_off[k + 1]], rtol=1e-5, atol=1e-6)
        total += float(lk)
    assert abs(total - float(ldj)) < 1e-4 * max(1.0, abs(float(ldj)))


def test_k5_index_formulas():
    from graph_normalizing_flows_b200 import graph_data as GD, grevnet_synthetic_data as SD, utils as U
    from graph_normalizing_flows_b200.graphs import networkxs_to_graphs_tuple
    import networkx as nx
    rng = np.random.default_rng(4)
    graphs = []
    for _ in range(6):
        g = nx.gnp_random_graph(int(rng.integers(3, 12)), 0.4, seed=int(rng.integers(1 << 30)))
        g = nx.relabel_nodes(g, {i: (i, i + 1) for i in g.nodes()})   # tuple ids like the grid family
        if g.number_of_nodes() > 2:
            first = list(g.nodes())[0]
            g.add_edge(first, first)                                   # pre-existing self-loop (citeseer)
        graphs.append(g)
    want = DO.networkxs_to_graphs_tuple([DO.convert_nx_repr(g.to_directed()) for g in graphs], None)
    from graph_normalizing_flows_b200.graphs import concat_structures
    got = concat_structures([GD.convert_nx_repr(g.to_directed()) for g in graphs])
    for f in ("senders", "receivers", "n_node", "n_edge"):
        a, b = getattr(want, f), getattr(got, f)
        assert a.dtype == b.dtype == np.int32 and np.array_equal(a, b), f
    got2 = networkxs_to_graphs_tuple([DO.convert_nx_repr(g.to_directed()) for g in graphs])
    assert np.array_equal(got2.senders, want.senders) and np.array_equal(got2.receivers, want.receivers)
    for n_node in ([3, 1, 4, 0, 2], [1], [], [7, 7]):
        a, b = DO.senders_receivers(n_node), U.senders_receivers(n_node)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and b[0].dtype == np.int32
    for n in (1, 2, 5, 9):
        a = DO.nx_to_arrays(DO.fully_connected_nx_graph(n))
        s, r = SD.fully_connected_edges(n)
        assert np.array_equal(a[1], s) and np.array_equal(a[2], r)


def test_family_fixtures_are_consistent():
    for path in sorted(glob.glob(os.path.join(H.GOLDEN, "graphs_*.npz"))):
        f = np.load(path)
        assert int(f["n_edge"].sum()) == len(f["senders_local"]) == len(f["receivers_local"])
        rep = np.repeat(f["n_node"], f["n_edge"])
        assert (f["senders_local"] < rep).all() and (f["receivers_local"] < rep).all()
        # convert_nx_repr: every node has its self-loop, and it comes first for its sender
        off = np.concatenate([[0], np.cumsum(f["n_edge"])])
        for k in range(0, len(f["n_node"]), max(1, len(f["n_node"]) // 10)):
            s = f["senders_local"][off[k]:off[k + 1]].astype(int)
            r = f["receivers_local"][off[k]:off[k + 1]].astype(int)
            assert (np.diff(s) >= 0).all()
            first = np.r_[True, np.diff(s) > 0]
            assert (r[first] == s[first]).all() and first.sum() == f["n_node"][k]


def test_torch_baseline_matches_numpy_oracle():
    import torch
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(6)
    g = H.random_batch(rng, 6, 4, 15, D=14)
    for block, agg in (("concat", "sum"), ("agg_then", "mean")):
        p = O.make_params(5, 2, 14, 64, 4, agg=agg, block=block, last_layer_scale=0.1)
        z, ldj = O.grevnet_f(g.nodes, g.senders, g.receivers, p)
        zt, lt = OT.grevnet_f(torch.from_numpy(g.nodes), torch.from_numpy(g.senders).long(),
                              torch.from_numpy(g.receivers).long(), OT.params_to_torch(p))
        assert np.allclose(zt.numpy(), z, rtol=1e-5, atol=1e-5)
        assert abs(float(lt) - float(ldj)) < 1e-4 * max(1.0, abs(float(ldj)))
        lp = O.log_prob(z, ldj, g.n_node)["log_prob_xs"]
        assert H.rel_err(OT.log_prob_xs(zt, lt), lp) < 1e-5


def test_torch_attention_restatement_matches_numpy_oracle():
    """f1: the differentiable torch restatement of DMSelfAttentionMLP (reference gradients for the GPU
    backward test) agrees with the numpy oracle it mirrors; its autograd runs (isolated receivers too)."""
    import torch
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(7)
    g = H.random_batch(rng, 5, 4, 12, D=6, isolated=True)
    g = g._replace(nodes=(g.nodes * 0.1).astype(np.float32))       # residual adds x to s: keep exp(s) tame
    for attn in (dict(num_heads=3, kq_dim=5, v_dim=4, out_dim=9, concat=True, residual=False, kq_dim_division=True),
                 dict(num_heads=2, kq_dim=3, v_dim=6, out_dim=5, concat=False, residual=True, kq_dim_division=False),
                 dict(num_heads=2, kq_dim=3, v_dim=6, out_dim=5, concat=True, residual=True, kq_dim_division=False,
                      layer_norm=True)):
        p = O.make_params(5, 2, 6, 32, 3, block="dm_attn", act="relu", attn=attn, last_layer_scale=0.1)
        z, ldj = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(p, np.float64))
        zt, lt = OT.grevnet_f(torch.from_numpy(g.nodes).double(), torch.from_numpy(g.senders).long(),
                              torch.from_numpy(g.receivers).long(), OT.params_to_torch(p, torch.float64))
        assert np.allclose(zt.numpy(), z, rtol=1e-10, atol=1e-10)
        assert abs(float(lt) - float(ldj)) < 1e-9 * max(1.0, abs(float(ldj)))
        loss, grads = OT.loss_and_grads(g.nodes, g.senders, g.receivers, p, 1.0 / g.nodes.shape[0])
        assert np.isfinite(grads).all() and grads.shape == (H.flat_from_oracle(p).shape[0],)


def test_torch_batch_norm_restatement_matches_numpy_oracle():
    """a9: the differentiable torch restatement of the TFP batch-norm bijector (reference gradients for the GPU
    test of the BN backward) reproduces oracle/gnf_oracle.py::grevnet_f_bn; a finite difference on gamma checks the
    autograd path through the batch statistics and the N-tiled log-det term."""
    import torch
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(8)
    g = H.random_batch(rng, 6, 4, 15, D=6)
    T = 2
    p = O.make_params(5, T, 6, 32, 3, last_layer_scale=0.1)
    gm = (1.0 + 0.2 * rng.standard_normal((2, T, 3))).clip(0.5, 1.5)
    bt = 0.1 * rng.standard_normal((2, T, 3))
    bns = O.make_bn_state(T, 3, np.float64)
    for half in range(2):
        for i in range(T):
            bns[half][i]["gamma"], bns[half][i]["beta"] = gm[half, i].copy(), bt[half, i].copy()
    z, ldj = O.grevnet_f_bn(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(p, np.float64), bns)
    bn_t = (torch.from_numpy(gm), torch.from_numpy(bt))
    zt, lt = OT.grevnet_f_autograd(torch.from_numpy(g.nodes).double(), torch.from_numpy(g.senders).long(),
                                   torch.from_numpy(g.receivers).long(), OT.params_to_torch(p, torch.float64), bn=bn_t)
    assert np.allclose(zt.numpy(), z, rtol=1e-10, atol=1e-10)
    assert abs(float(lt) - float(ldj)) < 1e-9 * max(1.0, abs(float(ldj)))
    loss, flat, gg, gb = OT.loss_and_grads(g.nodes, g.senders, g.receivers, p, 1.0, bn=(gm, bt))
    eps = 1e-6
    gp, gmn = gm.copy(), gm.copy()
    gp[1, 0, 2] += eps
    gmn[1, 0, 2] -= eps
    lp = OT.loss_and_grads(g.nodes, g.senders, g.receivers, p, 1.0, bn=(gp, bt))[0]
    lm = OT.loss_and_grads(g.nodes, g.senders, g.receivers, p, 1.0, bn=(gmn, bt))[0]
    assert abs((lp - lm) / (2 * eps) - gg[1, 0, 2]) < 1e-5 * max(1.0, abs(gg[1, 0, 2]))


def test_batch_norm_backward_formulas_match_autograd():
    """The closed forms the CUDA path implements for the bijector's backward (include/gnf_b200.h:
    gnf_bn_backward_sums / _coef / _apply) against torch autograd of the restatement, on the CPU:
        S1 = sum G_y, S2 = sum G_y xhat;  dL/dbeta = S1;  dL/dgamma = S2 - c N / gamma
        dL/dx = gamma/s * G_y + (c - gamma S2 / N)/s * xhat - gamma S1 / (N s),   c = loss_scale, s = sqrt(var + eps)
    for L = sum(G_y_fixed * y) - c * ildj(x)  (a downstream gradient G_y plus the bijector's own log-det term)."""
    import torch
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(9)
    n, h, c = 37, 5, 0.3
    x = torch.from_numpy(rng.standard_normal((n, h)) * 1.5 + 0.4).requires_grad_(True)
    gamma = torch.from_numpy(1.0 + 0.2 * rng.standard_normal(h)).requires_grad_(True)
    beta = torch.from_numpy(0.1 * rng.standard_normal(h)).requires_grad_(True)
    gy = torch.from_numpy(rng.standard_normal((n, h)))
    y, ildj = OT._bn_inverse(x, gamma, beta)
    loss = (gy * y).sum() - c * ildj
    gx_ref, gg_ref, gb_ref = torch.autograd.grad(loss, [x, gamma, beta])
    with torch.no_grad():
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        s = torch.sqrt(var + OT.BN_EPS)
        xhat = (y - beta) / gamma                                  # what the kernels recover from y
        s1, s2 = gy.sum(0), (gy * xhat).sum(0)
        gb, gg = s1, s2 - c * n / gamma
        gx = gamma / s * gy + (c - gamma * s2 / n) / s * xhat - gamma * s1 / (n * s)
        x_back = xhat * s + mean
    assert torch.allclose(gx, gx_ref, rtol=1e-10, atol=1e-12)
    assert torch.allclose(gg, gg_ref, rtol=1e-10, atol=1e-12) and torch.allclose(gb, gb_ref, rtol=1e-10, atol=1e-12)
    assert torch.allclose(x_back, x.detach(), rtol=1e-12, atol=1e-12)


def test_attention_backward_formulas_match_autograd():
    """The closed forms of k_attn_bwd_recv / k_attn_bwd_send (backward.cu) against torch autograd of the segment-softmax
    attention core (DMSelfAttention._build, gnn.py:419-477), on the CPU: with w_e the softmax weight of edge e at its
    receiver r and head h, g_w_e = <G[r,h,:], v[s_e,:]>, dot[r,h] = sum_e w_e g_w_e, g_l_e = w_e (g_w_e - dot) * inv_scale:
        dL/dqueries[r,h,:] = sum_{e into r} g_l_e keys[s_e,h,:]     dL/dkeys[s,h,:] = sum_{e out of s} g_l_e queries[r_e,h,:]
        dL/dv[s,:] = sum_h sum_{e out of s} w_e G[r_e,h,:]          (the value projection is shared by the heads)"""
    import torch
    rng = np.random.default_rng(10)
    g = H.random_batch(rng, 4, 4, 9, D=4, isolated=True)
    n, heads, kq, vd, inv_scale = g.nodes.shape[0], 3, 4, 5, 1.0 / np.sqrt(4.0)
    snd, rcv = torch.from_numpy(g.senders).long(), torch.from_numpy(g.receivers).long()
    keys = torch.from_numpy(rng.standard_normal((n, heads, kq))).requires_grad_(True)
    queries = torch.from_numpy(rng.standard_normal((n, heads, kq))).requires_grad_(True)
    vals = torch.from_numpy(rng.standard_normal((n, vd))).requires_grad_(True)
    G = torch.from_numpy(rng.standard_normal((n, heads, vd)))
    logits = (keys[snd] * queries[rcv]).sum(-1) * inv_scale
    idx = rcv[:, None].expand(-1, heads)
    seg_max = torch.full((n, heads), -float("inf"), dtype=torch.float64).scatter_reduce(0, idx, logits.detach(), "amax")
    ex = torch.exp(logits - seg_max[rcv])
    seg_sum = torch.zeros((n, heads), dtype=torch.float64).index_add_(0, rcv, ex)
    w = ex / seg_sum[rcv]
    att = torch.zeros((n, heads, vd), dtype=torch.float64).index_add_(0, rcv, vals[snd][:, None, :] * w[..., None])
    gk_ref, gq_ref, gv_ref = torch.autograd.grad((att * G).sum(), [keys, queries, vals])
    with torch.no_grad():
        gw = (G[rcv] * vals[snd][:, None, :]).sum(-1)                                   # [E, heads]
        dot = torch.zeros((n, heads), dtype=torch.float64).index_add_(0, rcv, w * gw)
        gl = w * (gw - dot[rcv]) * inv_scale
        g_queries = torch.zeros_like(queries).index_add_(0, rcv, gl[..., None] * keys[snd])
        g_keys = torch.zeros_like(keys).index_add_(0, snd, gl[..., None] * queries[rcv])
        g_v = torch.zeros_like(vals).index_add_(0, snd, (w[..., None] * G[rcv]).sum(1))
    assert torch.allclose(g_queries, gq_ref, rtol=1e-10, atol=1e-12)
    assert torch.allclose(g_keys, gk_ref, rtol=1e-10, atol=1e-12)
    assert torch.allclose(g_v, gv_ref, rtol=1e-10, atol=1e-12)


def test_f4_embedding_pickle_readers(tmp_path):
    """GrevnetDatasetFixed / Variable (train_grevnet_with_data.py:145-234) + transform_example (:237-271)."""
    import pickle
    from graph_normalizing_flows_b200 import train_grevnet_with_data as TD
    rng = np.random.default_rng(0)
    files = []
    for k in range(2):
        n_node = rng.integers(3, 9, size=10)
        emb = rng.standard_normal((int(n_node.sum()), 6)).astype(np.float32)
        pickle.dump((emb, n_node), open(tmp_path / f"part{k}.p", "wb"))
        files.append((emb, n_node))
    ds = TD.GrevnetDatasetFixed(str(tmp_path), 4)
    e, n = ds.train_batch()
    assert np.array_equal(n, files[0][1][:4]) and np.array_equal(e, files[0][0][:n.sum()])
    e, n = ds.train_batch()
    assert np.array_equal(n, files[0][1][4:8])
    e, n = ds.train_batch()                      # 2 graphs left < batch: next file (reference behaviour)
    assert np.array_equal(n, files[1][1][:4]) and np.array_equal(e, files[1][0][:n.sum()])
    dv = TD.GrevnetDatasetVariable(str(tmp_path), 20)
    seen = []
    for _ in range(3):
        e, n = dv.train_batch()
        assert e.shape[0] == n.sum() and n.sum() < 20
        seen.append(n)
    assert np.array_equal(np.concatenate(seen), files[0][1][:sum(len(s) for s in seen)])
    g = TD.transform_example(e, n)
    s, r = DO.senders_receivers(n)
    assert np.array_equal(g.senders, s) and np.array_equal(g.receivers, r)
    assert np.array_equal(g.n_edge, (n ** 2).astype(np.int32)) and g.nodes.dtype == np.float32


def test_f4_table_readers_reproduce_the_reference_state_machines(tmp_path):
    """The index-table readers (every file cut once into graph ranges) against statement-by-statement restatements
    of the reference's reader state machines (oracle/data_oracle.py), batch for batch until the files run out."""
    import pickle
    from graph_normalizing_flows_b200 import train_grevnet_with_data as TD
    rng = np.random.default_rng(3)
    files = []
    for k in range(3):
        n_node = rng.integers(3, 12, size=int(rng.integers(9, 17)))
        emb = rng.standard_normal((int(n_node.sum()), 4)).astype(np.float32)
        pickle.dump((emb, n_node), open(tmp_path / f"part{k}.p", "wb"))
        files.append((emb, n_node.astype(np.int64)))
    for batch in (3, 4, 5):
        ds = TD.GrevnetDatasetFixed(str(tmp_path), batch)
        count = 0
        for e_ref, n_ref in DO.dataset_fixed_batches(files, batch):
            e, n = ds.train_batch()
            assert np.array_equal(n, n_ref) and np.array_equal(e, e_ref)
            count += 1
        assert count == sum(len(f[1]) // batch for f in files)
        with pytest.raises(IndexError):
            ds.train_batch()                                   # past the last file, as the reference
    for max_nodes in (15, 24, 40):
        dv = TD.GrevnetDatasetVariable(str(tmp_path), max_nodes)
        count = 0
        for e_ref, n_ref in DO.dataset_variable_batches(files, max_nodes):
            e, n = dv.train_batch()
            assert np.array_equal(n, n_ref) and np.array_equal(e, e_ref), (max_nodes, count)
            count += 1
        assert count > len(files)


def test_f3_oracle_pred_adj_properties():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 4))
    p = O.pred_adj(x, [4, 5])
    assert not np.diag(p).any() and not p[:4, 4:].any() and np.allclose(p, p.T)
    d01 = ((x[0] - x[1]) ** 2).sum() / 2.0
    assert np.isclose(p[0, 1], 1 / (1 + np.exp(-10 * (1 - d01))))
