"""GPU tests added in round 2 (-m gpu): fp16 range guard, the embedding-flow default shape of
train_grevnet_with_data.py, NCCL world_size 2 through GraphShardedGRevNet, deferred index validation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import sharding

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOGPROB_RTOL = 1e-5


def dev_graph(g):
    return H.to_device_graph(g, DEV)


# ------------------------------------------------------------------------------- fp16 range guard ---
def _big_hidden_params():
    """Weights that drive a hidden PRE-activation past 65504 with O(1) inputs: first layer scaled up, later layers
    scaled down so that s, t (and exp(s)) stay ordinary numbers in fp32 arithmetic."""
    params = O.make_params(3, 2, 14, 128, 4, last_layer_scale=0.05)
    for which in ("s", "t"):
        for half in range(2):
            for i in range(2):
                layers = params[which][half][i]
                w0, b0 = layers[0]
                layers[0] = ((w0 * 3e5).astype(np.float32), b0)
                w1, b1 = layers[1]
                layers[1] = ((w1 / 3e5).astype(np.float32), b1)
    return params


@pytest.mark.parametrize("case", ["inputs_x1e3", "hidden_preactivation"])
def test_fp16_range_guard_raises_and_auto_mode_falls_back(case):
    """The reference computes in fp32 (unclamped exp, gnn.py:323): inputs or hidden activations beyond the fp16
    range are legal.  tc3x (fp16 hi/lo split) must not return inf/NaN silently: the device flag raises
    FloatingPointError at check_numerics(); math=None then re-runs in tc3x_bf16 and meets the fp64 oracle."""
    rng = np.random.default_rng(31)
    g = H.random_batch(rng, 10, 5, 30, D=14)
    if case == "inputs_x1e3":
        # node features x 1e3 and a sum aggregation over ~10 neighbours; the first layer is scaled down to keep s small
        g = g._replace(nodes=(g.nodes * 3e4).astype(np.float32))
        params = O.make_params(3, 2, 14, 128, 4, last_layer_scale=0.05)
        for which in ("s", "t"):
            for half in range(2):
                for i in range(2):
                    w0, b0 = params[which][half][i][0]
                    params[which][half][i][0] = ((w0 / 3e4).astype(np.float32), b0)
    else:
        params = _big_hidden_params()
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
    assert np.isfinite(z64).all() and np.isfinite(want)
    dg = dev_graph(g)
    # explicit tc3x: loud failure
    net = H.make_grevnet(params, 128, 4, device=DEV, math="tc3x")
    G.loss.log_prob(net, dg)
    with pytest.raises(FloatingPointError, match="fp16 range"):
        net.check_numerics()
    net.check_numerics()                         # the flag was consumed and reset
    # auto mode: first call trips the guard, later calls run in tc3x_bf16 and are right
    auto = H.make_grevnet(params, 128, 4, device=DEV, math=None)
    assert auto.math == "tc3x"
    G.loss.log_prob(auto, dg)
    with pytest.raises(FloatingPointError):
        auto.check_numerics()
    assert auto.math == "tc3x_bf16"
    out = G.loss.log_prob(auto, dg, return_z=True)
    auto.check_numerics()
    assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL
    scale = max(1.0, float(np.abs(z64).max()))
    assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4 * scale
    # the fp32 kernels agree as well (independent arithmetic)
    o32 = G.loss.log_prob(H.make_grevnet(params, 128, 4, device=DEV, math="fp32"), dg)
    assert H.rel_err(o32["log_prob_xs"], want) < LOGPROB_RTOL


def test_fp16_range_guard_is_quiet_on_ordinary_batches():
    rng = np.random.default_rng(32)
    g = H.random_batch(rng, 10, 5, 30, D=14)
    params = O.make_params(3, 2, 14, 256, 5, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device=DEV, math=None)
    for _ in range(3):
        out = G.loss.log_prob(net, dev_graph(g), return_z=True)
        net(out["z"], inverse=False)
    net.check_numerics()
    assert net.math == "tc3x"


# ------------------------------------------------- train_grevnet_with_data.py default shape (f1/f4) ---
def test_embedding_flow_default_shape_runs_and_matches_oracle():
    """train_grevnet_with_data.py:41-47,104-117 defaults: node_embedding_dim=200, gnn_latent_dim=2048,
    gnn_num_layers=3, dm_self_attn with attn_kq_dim=64, attn_v_dim=64, 1 head, concat_heads_output_dim=64, on fully
    connected graphs (round 1 rejected v_dim > 32).  Two coupling steps instead of ten keep the oracle in seconds."""
    from graph_normalizing_flows_b200 import utils as U
    rng = np.random.default_rng(33)
    n_node = np.array([5, 9, 7], np.int32)
    senders, receivers = U.senders_receivers(n_node)
    n = int(n_node.sum())
    nodes = (0.5 * rng.standard_normal((n, 200))).astype(np.float32)
    g = O.GraphsTuple(nodes, None, np.asarray(receivers, np.int32), np.asarray(senders, np.int32), None, n_node,
                      (n_node.astype(np.int64) ** 2).astype(np.int32))
    params = O.make_params(12345, 2, 200, 2048, 3, block="dm_attn", act="relu", last_layer_scale=0.02,
                           attn=dict(num_heads=1, kq_dim=64, v_dim=64, out_dim=64, concat=True, residual=False,
                                     kq_dim_division=True, layer_norm=False))
    z64, ldj64 = O.grevnet_f(nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, n_node)["log_prob_xs"]
    net = H.make_grevnet(params, 2048, 3, device=DEV)
    dg = dev_graph(g)
    # round 2: the 2048-wide MLP layers run on the tensor cores (k_gemm_tc, layer by layer), the attention projections
    # in k_linear_tc; math=None picks that path
    assert net.math == "tc3x"
    out = G.loss.log_prob(net, dg, return_z=True)
    assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL
    assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 2e-4
    net.check_numerics()
    net32 = H.make_grevnet(params, 2048, 3, device=DEV, math="fp32")          # the FFMA kernels, same flow
    out32 = G.loss.log_prob(net32, dg, return_z=True)
    assert H.rel_err(out32["log_prob_xs"], want) < LOGPROB_RTOL
    assert float((out32["z"].nodes - out["z"].nodes).abs().max()) < 2e-4
    x_back = net(out["z"], inverse=False).nodes
    assert float((x_back - dg.nodes).abs().max()) < 1e-3
    # and it trains: reversible backward vs autograd of the fp64 restatement
    from oracle import gnf_oracle_torch as OT
    loss_ref, grad_ref = OT.loss_and_grads(nodes, g.senders, g.receivers, params, 1.0 / n)
    scal, grads = net.loss_and_grad(dg, per_node=True)
    got = grads.double().cpu().numpy()
    assert abs(float(scal["loss_per_node"]) - loss_ref) <= 1e-5 * abs(loss_ref)
    assert np.abs(got - grad_ref).max() <= 5e-4 * np.abs(grad_ref).max()


# ------------------------------------------------------------------------- deferred validation safety ---
def test_deferred_validation_never_writes_out_of_bounds():
    """BatchPrefetcher builds the CSR before the bad-id count is read: out-of-range receivers must be skipped by the
    CSR kernels (no illegal address, no corruption of neighbouring allocations) and reported as ValueError."""
    rng = np.random.default_rng(34)
    g = H.random_batch(rng, 20, 5, 30, D=14)
    canary = torch.full((1 << 16,), 7, dtype=torch.int32, device=DEV)
    bad = g._replace(receivers=g.receivers.copy(), senders=g.senders.copy())
    n = g.nodes.shape[0]
    bad.receivers[::7] = n + 100000
    bad.receivers[3] = -5
    bad.senders[11] = 2 ** 30
    host = G.GraphsTuple(*[torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if v is not None else None for v in bad])
    pf = G.graphs.BatchPrefetcher(DEV)
    with pytest.raises(ValueError, match="outside"):
        pf.wait(pf.submit(host))
    with pytest.raises(ValueError, match="outside"):
        G.graphs.structure_of(G.GraphsTuple(*bad).to(DEV))
    torch.cuda.synchronize()                      # a sticky illegal-address error would surface here
    assert bool((canary == 7).all())
    ok = G.loss.log_prob(H.make_grevnet(O.make_params(1, 1, 14, 128, 3, last_layer_scale=0.05), 128, 3, device=DEV),
                         dev_graph(g))
    assert np.isfinite(float(ok["log_prob_xs"]))


# --------------------------------------------------------------------------------- NCCL, world size 2 ---
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices (gpurun --gpus 2)")
def test_nccl_two_ranks_sharded_wrapper_matches_unsharded():
    """GraphShardedGRevNet on real NCCL (one process per GPU): local_shard + log_prob / log_prob_async (4-vector
    all-reduce), loss_and_grad (gradient all-reduce), use_batch_norm (the [2H+1] moment and [2H] gradient-sum
    all-reduces inside every half step), broadcast_parameters -- against the unsharded single-GPU result."""
    import dist_worker as W
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=W.run, args=(r, 2, port, "nccl", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g, params, gm, bt = W.make_case()
    dg = dev_graph(g)
    n = g.nodes.shape[0]
    print("peer-memory all-reduce used:", res[0]["plain"]["peer_used"])
    for key, use_bn in (("plain", False), ("bn", True)):
        net = W.build_net(params, DEV, use_bn, gm, bt)
        ref = G.loss.log_prob(net, dg)
        G.loss.log_prob(net, dg)
        scal, grads = net.loss_and_grad(dg, per_node=True)
        want = [float(ref[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")]
        gref = grads.double().cpu().numpy()
        assert res[0][key]["n_local"] + res[1][key]["n_local"] == n
        for r in range(2):
            got = res[r][key]
            for a, b in zip(got["vec"], want):
                assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (key, r, a, b)
            assert got["vec"] == res[0][key]["vec"]                       # every rank holds the same global scalars
            # NVLink peer-memory all-reduce (one kernel with the log-prob assembly) vs NCCL: same sums up to the
            # order of two fp64 additions
            for a, b in zip(got["vec"], got["vec_nccl"]):
                assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (key, r, a, b)
            for a, b in zip(got["vec_async"], want):
                assert abs(a - b) <= 1e-6 * max(1.0, abs(b))
            assert abs(got["loss_per_node"] - float(scal["loss_per_node"])) <= 1e-6 * abs(float(scal["loss_per_node"]))
            assert np.abs(got["grads"] - gref).max() <= 2e-4 * np.abs(gref).max(), (key, r)
            assert np.array_equal(got["grads"], res[0][key]["grads"])      # all-reduced: identical on both ranks
            if use_bn:
                gg, gb = net.bn_gamma.grad.double().cpu().numpy(), net.bn_beta.grad.double().cpu().numpy()
                assert np.abs(got["g_gamma"] - gg).max() <= 2e-4 * np.abs(gg).max()
                assert np.abs(got["g_beta"] - gb).max() <= 2e-4 * max(np.abs(gb).max(), np.abs(gg).max())
                assert np.allclose(got["moving_mean"], net.bn_moving_mean.double().cpu().numpy(), atol=1e-6)


# ----------------------------------------------------- whole flow in one cooperative launch (persistent) ---
def _with_persist(value, fn):
    old = os.environ.get("GNF_PERSIST")
    os.environ["GNF_PERSIST"] = value
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("GNF_PERSIST", None)
        else:
            os.environ["GNF_PERSIST"] = old


@pytest.mark.parametrize("math", ["tc3x", "bf16", "tc3x_bf16"])
@pytest.mark.parametrize("shape", ["small", "many_tiles", "agg_then_mean_l128"])
def test_persistent_launch_matches_per_half_step_launches(shape, math):
    """gnf_grevnet_forward / _inverse run the whole flow either as 2T launches (one fused kernel per half step) or
    as ONE cooperative launch with a grid barrier between half steps (GNF_PERSIST=1).  Same tiles, same arithmetic:
    z and x must be bit-identical, the log-det equal up to the order of its fp64 partial sums; and the persistent
    path meets the oracle on its own."""
    rng = np.random.default_rng(51)
    if shape == "small":
        g = H.random_batch(rng, 6, 5, 40, D=14)                      # 1-2 tiles: fewer CTAs than SMs
        params, L, K = O.make_params(2, 3, 14, 256, 5, last_layer_scale=0.05), 256, 5
    elif shape == "many_tiles":
        g = H.random_batch(rng, 900, 20, 60, p_edge=0.15, D=14)      # ~36 k nodes, ~280 tiles: 2 tiles on most CTAs
        params, L, K = O.make_params(2, 2, 14, 256, 5, last_layer_scale=0.05), 256, 5
    else:
        g = H.random_batch(rng, 40, 5, 40, D=6, isolated=True)
        params, L, K = O.make_params(2, 4, 6, 128, 3, agg="mean", block="agg_then", eps=0.7, act="relu",
                                     last_layer_scale=0.1), 128, 3
    dg = dev_graph(g)
    net = H.make_grevnet(params, L, K, device=DEV, math=math)

    def run():
        out = G.loss.log_prob(net, dg, return_z=True)
        x = net(out["z"], inverse=False).nodes
        torch.cuda.synchronize()
        return out["z"].nodes.clone(), float(out["log_det_jacobian"]), float(out["log_prob_xs"]), x.clone()

    z0, ldj0, lp0, x0 = _with_persist("0", run)
    z1, ldj1, lp1, x1 = _with_persist("1", run)
    z1b, ldj1b, _, _ = _with_persist("1", run)
    assert torch.equal(z0, z1) and torch.equal(x0, x1)
    assert torch.equal(z1, z1b) and ldj1 == ldj1b                      # deterministic run to run
    assert abs(ldj0 - ldj1) <= 1e-12 * max(1.0, abs(ldj0))
    if shape != "many_tiles" and math != "bf16":
        z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
        want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
        assert H.rel_err(lp1, want) < LOGPROB_RTOL
        assert float((x1 - dg.nodes).abs().max()) < 1e-3
    net.check_numerics()


# ------------------------------------------------- wide MLP inputs: layer 0 in fp32, the rest on tensor cores ---
@pytest.mark.parametrize("case", ["mp_d24_concat", "mp_d32_mean", "attn_many_tiles"])
def test_inject_mode_wide_mlp_inputs(case):
    """MLP inputs wider than the fused kernel's 16-column layer-0 tile (message passing with D > 16; dm_self_attn):
    input assembly + layer 0 in the fp32 kernels, layers 1..K-1 + coupling update in the tcgen05 kernel (k_coupling_tc
    MODE inject).  Against the fp64 oracle, the all-fp32 kernels, and the round trip through the inverse."""
    rng = np.random.default_rng(61)
    if case == "mp_d24_concat":
        D, T, L, K = 24, 2, 256, 4
        g = H.random_batch(rng, 12, 5, 40, D=D)
        params = O.make_params(5, T, D, L, K, last_layer_scale=0.05)
    elif case == "mp_d32_mean":
        D, T, L, K = 32, 2, 128, 3
        g = H.random_batch(rng, 12, 5, 40, D=D, isolated=True)
        params = O.make_params(5, T, D, L, K, agg="mean", block="concat", act="relu", last_layer_scale=0.05)
    else:
        D, T, L, K = 2, 2, 256, 5
        n_node = rng.integers(10, 40, size=40)                       # ~1000 nodes: several tiles per launch
        s, r = G.utils.senders_receivers(n_node)
        nodes = rng.standard_normal((int(n_node.sum()), D)).astype(np.float32)
        g = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
        params = O.make_params(13, T, D, L, K, block="dm_attn", act="relu", last_layer_scale=0.1,
                               attn=dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False,
                                         kq_dim_division=False))
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
    dg = dev_graph(g)
    net = H.make_grevnet(params, L, K, device=DEV)
    assert net.math == "tc3x"
    res = {}
    for math in ("tc3x", "tc3x_bf16", "tc2x", "fp32"):
        net.math = math
        out = G.loss.log_prob(net, dg, return_z=True)
        res[math] = out["z"].nodes.clone()
        assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL, math
        tol = 1e-3 if math == "tc2x" else 1e-4
        assert np.abs(res[math].cpu().numpy() - z64).max() < tol * max(1.0, np.abs(z64).max()), math
        x_back = net(out["z"], inverse=False).nodes
        assert float((x_back - dg.nodes).abs().max()) < 10 * tol, math
    net.check_numerics()
    # training: the reversible backward of these shapes runs its MLPs on the tensor cores too (k_linear_tc for layer 0 and
    # g_h = delta_0 W_0^T, k_bwd_chain<BINJ> for the chains, k_dw_tc for every weight gradient incl. layer 0's); the
    # attention front end / aggregation below layer 0 stays on the fp32 kernels.  Tolerances as test_f2_* (BWD_TOL).
    net.math = None
    assert net._backward_math(None) == "tc3x"
    from oracle import gnf_oracle_torch as OT
    n = g.nodes.shape[0]
    for per_node in (True, False):
        scale = 1.0 / n if per_node else 1.0
        loss_ref, grad_ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, scale)
        for bmath, tol, ctol in ((None, 5e-4, 1e-6), ("tc3x_bf16", 1e-2, 1e-4), ("bf16", 1e-2, 1e-4), ("fp32", 2e-4, 1e-6)):
            scal, grads = net.loss_and_grad(dg, per_node=per_node, backward_math=bmath)
            got = grads.double().cpu().numpy()
            loss = float(scal["loss_per_node"] if per_node else scal["total_loss"])
            assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), bmath
            assert np.isfinite(got).all(), bmath
            assert np.abs(got - grad_ref).max() <= tol * np.abs(grad_ref).max(), (bmath, per_node)
            assert float(got @ grad_ref) / (np.linalg.norm(got) * np.linalg.norm(grad_ref)) > 1 - ctol, (bmath, per_node)
    # the backward reconstructs the input on its way
    z = G.loss.log_prob(net, dg, return_z=True)["z"]
    _, x_rec = net.backward_from_z(dg, z.nodes, 1.0, return_x=True)
    assert float((x_rec - dg.nodes).abs().max()) < 1e-4


# ------------------------------------------- attention on graphs wider than the staged kernels' 128-row window ---
@pytest.mark.parametrize("bmath", ["tc3x", "fp32"])
def test_attention_kernels_hand_wide_graphs_back(bmath):
    """The block-staged attention kernels (forward k_dm_attn_block, backward k_attn_bwd_block<0/1>) stage the compact
    row range a group of 32 nodes touches; a graph of 300 nodes with random edges does not fit the 128-row window, so
    its groups go back to the thread-per-head kernels while the small graphs of the same batch stay staged.  Density
    pass vs the fp64 oracle, reversible backward vs autograd, both backward arithmetics."""
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(71)
    D, T, L, K = 4, 1, 128, 3
    small = H.random_batch(rng, 6, 5, 30, D=D)
    wide = H.random_batch(rng, 1, 300, 300, p_edge=0.02, D=D)
    off = small.nodes.shape[0]
    g = O.GraphsTuple(nodes=np.concatenate([small.nodes, wide.nodes]).astype(np.float32), edges=None,
                      receivers=np.concatenate([small.receivers, wide.receivers + off]).astype(np.int32),
                      senders=np.concatenate([small.senders, wide.senders + off]).astype(np.int32), globals=None,
                      n_node=np.concatenate([small.n_node, wide.n_node]).astype(np.int32),
                      n_edge=np.concatenate([small.n_edge, wide.n_edge]).astype(np.int32))
    attn = dict(num_heads=4, kq_dim=6, v_dim=5, out_dim=12, concat=True, residual=False, kq_dim_division=True)
    params = O.make_params(23, T, D, L, K, block="dm_attn", act="leaky_relu", attn=attn, last_layer_scale=0.1)
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
    net = H.make_grevnet(params, L, K, device=DEV)
    dg = dev_graph(g)
    out = G.loss.log_prob(net, dg, return_z=True)
    assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4 * max(1.0, np.abs(z64).max())
    assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL
    n = g.nodes.shape[0]
    loss_ref, grad_ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, 1.0 / n)
    scal, grads = net.loss_and_grad(dg, per_node=True, backward_math=bmath)
    got = grads.double().cpu().numpy()
    assert abs(float(scal["loss_per_node"]) - loss_ref) <= 1e-5 * abs(loss_ref)
    assert np.abs(got - grad_ref).max() <= (5e-4 if bmath == "tc3x" else 2e-4) * np.abs(grad_ref).max()
    assert float(got @ grad_ref) / (np.linalg.norm(got) * np.linalg.norm(grad_ref)) > 1 - 1e-6


# ------------------------------------------------- layered tensor-core Linear (gemm_tc.cu) on its own ---
@pytest.mark.parametrize("shape", [(300, 164, 2048), (300, 2048, 2048), (517, 2048, 100), (129, 16, 16), (1, 40, 300),
                                   (700, 100, 104), (256, 32, 512), (40000, 176, 512)])
@pytest.mark.parametrize("math,tol", [("tc3x", 2e-6), ("tc3x_bf16", 4e-5), ("bf16", 2e-2)])
def test_layered_tensor_core_linear(shape, math, tol):
    """k_gemm_tc (gnf_debug_linear_tc): act(a @ w + b) for the widths the fused kernels do not take -- the 2048-wide
    layers of train_grevnet_with_data.py:41-47, ragged K / N (not multiples of 16 or of the 256-column block), row
    counts below one tile and far above one wave -- against a float64 matmul; the error bound scales with
    (|a| @ |w|), the sum a rounding error is relative to."""
    from graph_normalizing_flows_b200 import _lib
    lib = _lib.load()
    m, k, n = shape
    if m > 10000 and math != "tc3x":
        pytest.skip("the many-items case runs once")
    gen = torch.Generator(device="cpu").manual_seed(m + k + n)
    a = torch.randn(m, k, generator=gen)
    w = torch.randn(k, n, generator=gen) / np.sqrt(k)
    b = torch.randn(n, generator=gen)
    dev = torch.device(DEV)
    ad, wd, bd = a.to(dev), w.to(dev), b.to(dev)
    pre = a.double() @ w.double() + b.double()
    scale = (a.double().abs() @ w.double().abs()).max().item() + b.abs().max().item()
    wsb = lib.gnf_debug_linear_tc_workspace(k, n)
    for act, fn in ((2, lambda x: x), (0, lambda x: torch.nn.functional.leaky_relu(x, 0.2)), (1, torch.relu)):
        c = torch.full((m, n), float("nan"), device=dev)
        ws = _lib.workspace(wsb, dev)
        _lib.check(lib.gnf_debug_linear_tc(_lib.ptr(ad), _lib.ptr(wd), _lib.ptr(bd), m, k, n, act, _lib.MATH[math],
                                           _lib.ptr(c), _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), "gnf_debug_linear_tc")
        got = c.double().cpu()
        assert torch.isfinite(got).all(), (shape, math, act)
        assert float((got - fn(pre)).abs().max()) <= tol * scale, (shape, math, act)


# ---------------------------------------------- warp-per-(receiver, head) attention (k_dm_attn_warp) ---
@pytest.mark.parametrize("shape", [(3, 33, 40), (1, 64, 64), (2, 32, 16)])
def test_attention_warp_kernel_shapes(shape):
    """Wide keys / values (beyond the block-staged kernel's 10 value columns) on fully connected graphs with more than
    32 in-edges per node (several chunks, running max moves) next to tiny graphs: kq not a multiple of 4 (scalar key
    loads), v_dim above 32 (second value column per lane), several heads.  Density pass vs the fp64 oracle under every
    arithmetic; reversible backward vs autograd."""
    from graph_normalizing_flows_b200 import utils as U
    from oracle import gnf_oracle_torch as OT
    heads, kq, vd = shape
    rng = np.random.default_rng(heads * 100 + kq)
    D, T, L, K = 8, 1, 64, 3
    n_node = np.array([70, 3, 45, 1], np.int32)
    senders, receivers = U.senders_receivers(n_node)
    n = int(n_node.sum())
    nodes = rng.standard_normal((n, D)).astype(np.float32)
    g = O.GraphsTuple(nodes, None, np.asarray(receivers, np.int32), np.asarray(senders, np.int32), None, n_node,
                      (n_node.astype(np.int64) ** 2).astype(np.int32))
    attn = dict(num_heads=heads, kq_dim=kq, v_dim=vd, out_dim=24, concat=True, residual=False, kq_dim_division=True)
    params = O.make_params(5, T, D, L, K, block="dm_attn", act="leaky_relu", attn=attn, last_layer_scale=0.1)
    z64, ldj64 = O.grevnet_f(nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, n_node)["log_prob_xs"]
    dg = dev_graph(g)
    for math in ("tc3x", "fp32"):
        net = H.make_grevnet(params, L, K, device=DEV, math=math)
        out = G.loss.log_prob(net, dg, return_z=True)
        assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4 * max(1.0, np.abs(z64).max()), math
        assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL, math
    loss_ref, grad_ref = OT.loss_and_grads(nodes, g.senders, g.receivers, params, 1.0 / n)
    scal, grads = net.loss_and_grad(dg, per_node=True)
    got = grads.double().cpu().numpy()
    assert abs(float(scal["loss_per_node"]) - loss_ref) <= 1e-5 * abs(loss_ref)
    assert np.abs(got - grad_ref).max() <= 2e-4 * np.abs(grad_ref).max()


# ------------------------------------------- batch-norm flow: one library call vs the per-half-step chain ---
@pytest.mark.parametrize("math", ["tc3x", "fp32"])
def test_batch_norm_chained_entry_equals_per_half_step_calls(math):
    """gnf_grevnet_forward_bn / gnf_grevnet_inverse_bn run the same kernels in the same order as the host mirror's
    half-step-by-half-step loop (the path sharded runs keep): z, log-det, saved batch statistics, moving statistics and
    the sampling direction agree bit for bit; the training step works from the statistics the chained call saved."""
    rng = np.random.default_rng(3)
    g = H.random_batch(rng, 20, 4, 25, D=14)
    g = g._replace(nodes=(g.nodes * 1.3 - 0.2).astype(np.float32))
    params = O.make_params(8, 2, 14, 128, 3, last_layer_scale=0.05)
    dg = dev_graph(g)
    res = {}
    for chain in (True, False):
        net = H.make_grevnet(params, 128, 3, device=DEV, math=math)
        net.use_batch_norm = True
        net.bn_chain = chain
        with torch.no_grad():
            net.bn_gamma.copy_(1.0 + 0.1 * torch.arange(net.bn_gamma.numel(), device=DEV).reshape(net.bn_gamma.shape) / 20)
            net.bn_beta.copy_(0.05 * torch.ones_like(net.bn_beta))
        out = G.loss.log_prob(net, dg, return_z=True)
        out2 = G.loss.log_prob(net, dg, return_z=True)            # second pass: moving statistics moved once already
        x_back = net(out2["z"], inverse=False).nodes
        saved = torch.stack([net._bn_saved[(h, i)].clone() for h in range(2) for i in range(2)])
        scal, grads = net.loss_and_grad(dg, per_node=True)
        res[chain] = (out["z"].nodes, out["log_det_jacobian"], out2["z"].nodes, x_back, saved,
                      net.bn_moving_mean.clone(), net.bn_moving_var.clone(), grads, net.bn_gamma.grad.clone())
    for a, b in zip(res[True], res[False]):
        assert torch.equal(a, b)
