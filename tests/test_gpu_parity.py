"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs, against the committed golden fixtures, and -- at BASELINE's full size --
through size-independent properties.

Tolerances (north star): indexing and segment ops BIT-EXACT; floating-point log-prob within
1e-5 relative (fp32 and tc3x modes).  The single-pass bf16 mode has its own, looser, stated
tolerance (1e-4 relative on log-prob, 2e-2 absolute on z).
"""
import glob
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib, sharding

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOGPROB_RTOL = 1e-5
CASES = sorted(glob.glob(os.path.join(H.GOLDEN, "golden_*.npz")))


def dev_graph(g):
    return H.to_device_graph(g, DEV)


def load_case(path):
    g = np.load(path, allow_pickle=False)
    T, D, L, K, ws = [int(v) for v in g["spec"]]
    params = O.make_params(int(g["seed"]), T, D, L, K, agg=str(g["agg"]), block=str(g["block"]),
                           eps=float(g["eps"]), act=str(g["act"]), last_layer_scale=float(g["last_scale"]),
                           weight_sharing=bool(ws))
    graph = O.GraphsTuple(g["nodes"], None, g["receivers"], g["senders"], None, g["n_node"], g["n_edge"])
    return g, graph, params, (T, D, L, K)


# ------------------------------------------------------------------------------- a1: CSR -------
@pytest.mark.parametrize("kind", ["random", "fc_big_degree", "isolated", "single_node"])
def test_csr_is_stable_and_exact(kind):
    rng = np.random.default_rng(0)
    if kind == "random":
        g = H.random_batch(rng, 40, 3, 60, D=2)
    elif kind == "fc_big_degree":            # in-degree 70 > 32 exercises the long-segment path
        s, r = G.utils.senders_receivers([70, 3, 45])
        perm = rng.permutation(len(s))       # and an arbitrary (not sender-sorted) edge order
        g = O.GraphsTuple(np.zeros((118, 2), np.float32), None, r[perm], s[perm], None,
                          np.array([70, 3, 45], np.int32), np.array([4900, 9, 2025], np.int32))
    elif kind == "isolated":
        g = H.random_batch(rng, 10, 4, 20, D=2, isolated=True)
    else:
        g = O.GraphsTuple(np.zeros((1, 2), np.float32), None, np.zeros(1, np.int32), np.zeros(1, np.int32), None,
                          np.ones(1, np.int32), np.ones(1, np.int32))
    dg = dev_graph(g)
    st = G.graphs.structure_of(dg)
    n = g.nodes.shape[0]
    order = np.argsort(g.receivers, kind="stable")
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(g.receivers, minlength=n))]).astype(np.int32)
    assert np.array_equal(st.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(st.perm.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(st.csr_senders.cpu().numpy(), g.senders[order])
    assert G.graphs.structure_of(dg.replace(nodes=dg.nodes * 2)) is st        # cached across .replace


def test_index_validation_raises_like_tf():
    g = H.random_batch(np.random.default_rng(1), 3, 3, 6, D=4)
    bad = g._replace(senders=g.senders.copy())
    bad.senders[2] = g.nodes.shape[0]                      # out of range -> TF InvalidArgumentError
    with pytest.raises(ValueError, match="outside"):
        G.graphs.structure_of(dev_graph(bad))
    neg = g._replace(receivers=g.receivers.copy())
    neg.receivers[0] = -1
    with pytest.raises(ValueError, match="outside"):
        G.graphs.structure_of(dev_graph(neg))
    wrong = g._replace(n_node=g.n_node + 1)
    with pytest.raises(ValueError, match="n_node"):
        G.graphs.structure_of(dev_graph(wrong))
    with pytest.raises(ValueError):
        G.graphs.BatchStructure(dev_graph(g).senders.long(), dev_graph(g).receivers, g.nodes.shape[0])


# ---------------------------------------------------------------------- a3 / a4: segment ops ----
@pytest.mark.parametrize("h", [1, 2, 7, 100])
@pytest.mark.parametrize("agg", ["sum", "mean"])
def test_segment_ops_bit_exact(h, agg):
    """K6 + K7: serial in-edge-order fp32 accumulation, bit for bit; empty segments -> 0."""
    rng = np.random.default_rng(h)
    g = H.random_batch(rng, 30, 3, 50, D=2, isolated=True)
    x = (rng.standard_normal((g.nodes.shape[0], h)) * 10).astype(np.float32)
    want = O.aggregate(x, g.senders, g.receivers, agg)
    dg = dev_graph(g._replace(nodes=x))
    st = G.graphs.structure_of(dg)
    lib = _lib.load()
    fused = G.gnn.gather_segment_reduce(dg.nodes, st, agg)
    assert np.array_equal(fused.cpu().numpy(), want)
    # the two-step form of the reference: gather (a3) then aggregate (a4)
    edges = torch.empty(len(g.senders), h, device=DEV)
    _lib.check(lib.gnf_gather_rows(_lib.ptr(dg.nodes), h, _lib.ptr(dg.senders), len(g.senders), _lib.ptr(edges),
                                   _lib.stream_ptr()))
    assert np.array_equal(edges.cpu().numpy(), x[g.senders])
    fn = G.gnn.unsorted_segment_sum if agg == "sum" else G.gnn.unsorted_segment_mean
    two = fn(edges, dg.receivers, g.nodes.shape[0])
    assert np.array_equal(two.cpu().numpy(), want)
    iso = np.setdiff1d(np.arange(g.nodes.shape[0]), g.receivers)
    assert len(iso) > 0 and not fused.cpu().numpy()[iso].any()


# --------------------------------------------------------------------- a2-a8: golden parity -----
@pytest.mark.parametrize("math", ["fp32", "tc3x", "tc3x_bf16"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_golden_parity(path, math):
    gold, graph, params, (T, D, L, K) = load_case(path)
    net = H.make_grevnet(params, L, K, device=DEV, math=math)
    # (latent widths other than 128 / 256 run layer by layer in k_gemm_tc under the tensor-core modes)
    dg = dev_graph(graph)
    out = G.loss.log_prob(net, dg, return_z=True)
    z = out["z"].nodes.cpu().numpy()
    assert np.abs(z - gold["z64"]).max() < 5e-5
    assert np.abs(z - gold["z"]).max() < 5e-5
    assert H.rel_err(out["log_prob_xs"], gold["log_prob_xs64"]) < LOGPROB_RTOL
    assert H.rel_err(out["log_prob_xs"], gold["log_prob_xs"]) < LOGPROB_RTOL
    assert H.rel_err(out["log_prob_zs"], gold["log_prob_zs64"]) < LOGPROB_RTOL
    assert abs(float(out["log_det_jacobian"]) - float(gold["ldj64"])) < LOGPROB_RTOL * abs(float(gold["log_prob_xs64"]))
    assert float(out["num_nodes"]) == graph.nodes.shape[0]
    # reference API: (GraphsTuple, scalar) / GraphsTuple, inverse=True is the density direction
    z2, ldj = net(dg, inverse=True)
    assert isinstance(z2, G.GraphsTuple) and ldj.dim() == 0 and z2.senders is dg.senders
    x_back = net(z2, inverse=False).nodes.cpu().numpy()
    assert np.abs(x_back - graph.nodes).max() < 5e-5                     # K1
    assert np.abs(x_back - gold["x_back"]).max() < 5e-5
    assert H.rel_err(net.log_prob(dg), gold["log_prob_xs64"]) < LOGPROB_RTOL


@pytest.mark.parametrize("path", [p for p in CASES if "caveman_small" in p or "concat_mean_d14" in p])
def test_bf16_single_pass_has_its_own_tolerance(path):
    gold, graph, params, (T, D, L, K) = load_case(path)
    net = H.make_grevnet(params, L, K, device=DEV, math="bf16")
    out = G.loss.log_prob(net, dev_graph(graph), return_z=True)
    assert H.rel_err(out["log_prob_xs"], gold["log_prob_xs64"]) < 1e-4
    assert np.abs(out["z"].nodes.cpu().numpy() - gold["z64"]).max() < 2e-2


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_tc2x_meets_the_log_prob_tolerance(path):
    """tc2x = fp16 activations (one unbiased rounding) x fp16 hi/lo weights, 2 MMAs per product.
    The north-star tolerance (1e-5 relative on log-prob) holds; per-element z is only ~1e-3-close,
    which is why tc3x stays the default."""
    gold, graph, params, (T, D, L, K) = load_case(path)
    if L not in (128, 256):
        pytest.skip("shape not served by the fused kernel")
    net = H.make_grevnet(params, L, K, device=DEV, math="tc2x")
    out = G.loss.log_prob(net, dev_graph(graph), return_z=True)
    assert H.rel_err(out["log_prob_xs"], gold["log_prob_xs64"]) < LOGPROB_RTOL
    assert np.abs(out["z"].nodes.cpu().numpy() - gold["z64"]).max() < 5e-3


@pytest.mark.parametrize("math", ["fp32", "tc3x"])
def test_k2_zero_last_layer_is_bit_exact_identity(math):
    rng = np.random.default_rng(2)
    g = H.random_batch(rng, 20, 5, 40, D=14)
    params = O.make_params(1, 3, 14, 256, 4, last_layer_scale=0.0)
    net = H.make_grevnet(params, 256, 4, device=DEV, math=math)
    dg = dev_graph(g)
    out = G.loss.log_prob(net, dg, return_z=True)
    assert torch.equal(out["z"].nodes, dg.nodes)
    assert float(out["log_det_jacobian"]) == 0.0
    n, d = g.nodes.shape
    want = -0.5 * float((g.nodes.astype(np.float64) ** 2).sum()) - 0.5 * n * d * np.log(2 * np.pi)
    assert H.rel_err(out["log_prob_xs"], want) < 1e-12
    assert torch.equal(net(out["z"], inverse=False).nodes, dg.nodes)


@pytest.mark.parametrize("math", ["fp32", "tc3x"])
def test_k4_graph_independence_and_shard_additivity(math):
    rng = np.random.default_rng(3)
    g = H.random_batch(rng, 37, 5, 40, D=14)
    params = O.make_params(4, 2, 14, 256, 5, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device=DEV, math=math)
    full = G.loss.log_prob(net, dev_graph(g), return_z=True)
    zfull = full["z"].nodes.cpu().numpy()
    node_off = np.concatenate([[0], np.cumsum(g.n_node)])
    tot = np.zeros(4)
    for world in (3,):
        parts = sharding.partition_graphs(g.n_node, g.n_edge, world)
        for ids in parts:
            sh = sharding.shard_graphs_tuple(G.GraphsTuple(*g), ids)
            out = G.loss.log_prob(net, sh.to(DEV), return_z=True)
            zs = out["z"].nodes.cpu().numpy()
            rows = np.concatenate([np.arange(node_off[k], node_off[k + 1]) for k in ids])
            assert np.abs(zs - zfull[rows]).max() < 1e-6          # other graphs in the batch do not matter
            tot += np.array([float(out[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")])
    want = np.array([float(full[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")])
    assert np.all(np.abs(tot - want) <= 1e-7 * np.maximum(1.0, np.abs(want)))


def test_sharded_wrapper_world_size_one():
    rng = np.random.default_rng(5)
    g = H.random_batch(rng, 10, 5, 20, D=14)
    params = O.make_params(4, 2, 14, 128, 3, last_layer_scale=0.05)
    net = H.make_grevnet(params, 128, 3, device=DEV)
    assert net.math == "tc3x"
    sh = sharding.GraphShardedGRevNet(net)
    local = sh.local_shard(G.GraphsTuple(*g)).to(DEV)
    out = sh.log_prob(local)
    z, ldj = O.grevnet_f(g.nodes, g.senders, g.receivers, params)
    assert H.rel_err(out["log_prob_xs"], O.log_prob(z, ldj, g.n_node)["log_prob_xs"]) < LOGPROB_RTOL


@pytest.mark.parametrize("block,agg", [("concat", "sum"), ("concat", "mean"), ("agg_then", "sum"), ("agg_then", "mean")])
def test_node_block_gnn_standalone(block, agg):
    """NodeBlockGNN as a callable GraphsTuple -> GraphsTuple (gnn.py:155-156)."""
    from functools import partial
    rng = np.random.default_rng(6)
    g = H.random_batch(rng, 8, 4, 20, D=7)                       # nodes are [N, H=7] here
    mlp_fn = partial(G.make_mlp_model, 64, 7, 3, G.gnn.leaky_relu, 0.1, 0.1)
    fac = {("concat", "sum"): G.sum_concat_then_mlp_gnn, ("concat", "mean"): G.avg_concat_then_mlp_gnn,
           ("agg_then", "sum"): partial(G.sum_then_mlp_gnn, epsilon=0.7),
           ("agg_then", "mean"): partial(G.avg_then_mlp_gnn, epsilon=0.7)}[(block, agg)]
    gnn = fac(mlp_fn)
    dg = dev_graph(g)
    out = gnn(dg)
    layers = [(w.cpu().numpy(), b.cpu().numpy()) for w, b in zip(gnn.mlp.weights, gnn.mlp.biases)]
    want = O.node_block_gnn(g.nodes, g.senders, g.receivers, layers,
                            {"agg": agg, "block": block, "eps": 0.7, "act": "leaky_relu"})
    assert isinstance(out, G.GraphsTuple) and out.nodes.shape == dg.nodes.shape
    assert np.abs(out.nodes.cpu().numpy() - want).max() < 1e-4


@pytest.mark.parametrize("math", ["fp32", "tc3x"])
def test_a9_batch_norm_bijector(math):
    """use_batch_norm=True (gnn.py:310-313,325-328,356-358,369-371) against the oracle's restatement of
    tfb.BatchNormalization(training=True): batch statistics + N-tiled ildj in the density direction,
    moving statistics in the sampling direction, moving averages updated per density pass."""
    rng = np.random.default_rng(12)
    g = H.random_batch(rng, 14, 5, 30, D=14)
    g = g._replace(nodes=(g.nodes * 1.7 + 0.3).astype(np.float32))
    params = O.make_params(6, 3, 14, 256, 4, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 4, device=DEV, math=math)
    net.use_batch_norm = True
    bns = O.make_bn_state(3, 7)
    for half in range(2):                      # non-trivial gamma / beta
        for i in range(3):
            gm = (1.0 + 0.2 * rng.standard_normal(7)).astype(np.float32).clip(0.5, 1.5)
            bt = (0.1 * rng.standard_normal(7)).astype(np.float32)
            bns[half][i]["gamma"], bns[half][i]["beta"] = gm, bt
            net.bn_gamma.data[half, i] = torch.from_numpy(gm).to(DEV)
            net.bn_beta.data[half, i] = torch.from_numpy(bt).to(DEV)
    dg = dev_graph(g)
    for rep in range(2):                       # second pass starts from updated moving statistics
        z_ref, ldj_ref = O.grevnet_f_bn(g.nodes.astype(np.float64), g.senders, g.receivers,
                                        O.cast_params(params, np.float64),
                                        bns)
        out = G.loss.log_prob(net, dg, return_z=True)
        want = O.log_prob(z_ref, ldj_ref, g.n_node)
        assert np.abs(out["z"].nodes.cpu().numpy() - z_ref).max() < 1e-4
        assert H.rel_err(out["log_prob_xs"], want["log_prob_xs"]) < LOGPROB_RTOL
        assert abs(float(out["log_det_jacobian"]) - float(ldj_ref)) < LOGPROB_RTOL * abs(float(want["log_prob_xs"]))
    for half in range(2):
        for i in range(3):
            assert np.allclose(net.bn_moving_mean[half, i].cpu().numpy(), bns[half][i]["moving_mean"], atol=1e-5)
            assert np.allclose(net.bn_moving_var[half, i].cpu().numpy(), bns[half][i]["moving_var"], atol=1e-5)
    x_ref = O.grevnet_g_bn(z_ref, g.senders, g.receivers, O.cast_params(params, np.float64), bns)
    x = net(out["z"], inverse=False).nodes.cpu().numpy()
    assert np.abs(x - x_ref).max() < 1e-3 * max(1.0, np.abs(x_ref).max())


# backward arithmetic -> (max-norm tolerance, 1 - cosine tolerance); see include/gnf_b200.h (f2)
BWD_TOL = {"fp32": (2e-4, 1e-6), "tc3x": (5e-4, 1e-6), "bf16": (1e-2, 1e-4)}


@pytest.mark.parametrize("bmath", ["fp32", "tc3x", "bf16"])
@pytest.mark.parametrize("block,agg,ws,act", [("concat", "sum", False, "leaky_relu"), ("concat", "mean", False, "relu"),
                                              ("agg_then", "sum", True, "leaky_relu"), ("agg_then", "mean", False, "leaky_relu")])
def test_f2_reversible_backward_matches_autograd(block, agg, ws, act, bmath):
    """Row f2: analytic reversible backward (gnf_grevnet_backward) vs torch autograd of the fp64
    torch restatement of the reference.  Stated tolerances (of the gradient's max-norm over the
    whole parameter vector): fp32 FFMA kernels 2e-4; tensor-core backward "tc3x" (fp16 hi/lo
    recomputed forward chains, bf16 hi/lo gradient operands) 5e-4; "bf16" (bf16 hi/lo forward chains:
    pre-activations good to ~4e-5, so a few act' masks flip at the leaky-relu kink -- each flip is a
    1/N effect, visible at these tiny N -- and a single-bf16 weight-gradient GEMM) 1e-2."""
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(21)
    D, T, L, K = (14, 2, 128, 4) if block == "concat" else (6, 3, 128, 3)
    g = H.random_batch(rng, 9, 4, 25, D=D, isolated=(agg == "mean"))
    params = O.make_params(8, T, D, L, K, agg=agg, block=block, eps=0.8, act=act, last_layer_scale=0.2,
                           weight_sharing=ws)
    n = g.nodes.shape[0]
    tol, ctol = BWD_TOL[bmath]
    for per_node in (True, False):
        scale = 1.0 / n if per_node else 1.0
        loss_ref, grad_ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, scale)
        net = H.make_grevnet(params, L, K, device=DEV)
        out, grads = net.loss_and_grad(dev_graph(g), per_node=per_node, backward_math=bmath)
        got = grads.cpu().numpy().astype(np.float64)
        loss = float(out["loss_per_node"] if per_node else out["total_loss"])
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
        assert got.shape == grad_ref.shape and np.isfinite(got).all()
        assert np.abs(got - grad_ref).max() <= tol * np.abs(grad_ref).max()
        # cosine of the whole gradient vector
        assert float(got @ grad_ref) / (np.linalg.norm(got) * np.linalg.norm(grad_ref)) > 1 - ctol
        assert net.params.grad is grads
    # the backward reconstructs the input on its way (reversibility)
    z = out["z"]
    _, x_rec = net.backward_from_z(dev_graph(g), z.nodes, 1.0, return_x=True, math=bmath)
    assert np.abs(x_rec.cpu().numpy() - g.nodes).max() < 5e-5


@pytest.mark.parametrize("bmath", ["fp32", "tc3x"])
def test_f2_f4_backward_through_batch_norm(bmath):
    """use_batch_norm=True training step: reversed half steps (gnf_coupling_half_backward) interleaved with the
    backward of the batch-norm bijector in training mode (gnf_bn_backward_sums / _apply: batch statistics depend
    on x, the -N/2 log(var+eps) log-det term too) vs torch autograd of the fp64 restatement
    (oracle/gnf_oracle_torch.py::_bn_inverse, TFP semantics [upstream, unverifiable]).  Flow parameters, gamma and
    beta gradients; the input is reconstructed with the saved batch statistics."""
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(27)
    D, T, L, K = 14, 2, 128, 4
    g = H.random_batch(rng, 10, 5, 25, D=D)
    g = g._replace(nodes=(g.nodes * 1.7 + 0.3).astype(np.float32))
    params = O.make_params(6, T, D, L, K, last_layer_scale=0.1)
    gm = (1.0 + 0.2 * rng.standard_normal((2, T, D // 2))).astype(np.float32).clip(0.5, 1.5)
    bt = (0.1 * rng.standard_normal((2, T, D // 2))).astype(np.float32)
    n = g.nodes.shape[0]
    tol, ctol = BWD_TOL[bmath]
    for per_node in (True, False):
        scale = 1.0 / n if per_node else 1.0
        loss_ref, grad_ref, gg_ref, gb_ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, scale, bn=(gm, bt))
        net = H.make_grevnet(params, L, K, device=DEV)
        net.use_batch_norm = True
        net.bn_gamma.data.copy_(torch.from_numpy(gm))
        net.bn_beta.data.copy_(torch.from_numpy(bt))
        out, grads = net.loss_and_grad(dev_graph(g), per_node=per_node, backward_math=bmath)
        loss = float(out["loss_per_node"] if per_node else out["total_loss"])
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
        got = grads.double().cpu().numpy()
        assert np.isfinite(got).all()
        assert np.abs(got - grad_ref).max() <= tol * np.abs(grad_ref).max()
        assert float(got @ grad_ref) / (np.linalg.norm(got) * np.linalg.norm(grad_ref)) > 1 - ctol
        gg, gb = net.bn_gamma.grad.double().cpu().numpy(), net.bn_beta.grad.double().cpu().numpy()
        assert np.abs(gg - gg_ref).max() <= max(tol, 1e-4) * np.abs(gg_ref).max()
        assert np.abs(gb - gb_ref).max() <= max(tol, 1e-4) * max(np.abs(gb_ref).max(), np.abs(gg_ref).max())
    _, x_rec = net.backward_from_z(dev_graph(g), out["z"].nodes, 1.0, return_x=True, math=bmath)
    assert np.abs(x_rec.cpu().numpy() - g.nodes).max() < 5e-5


def test_f2_tensor_core_backward_deep_mlp():
    """K = 7 layers: more hidden activations than the shared-memory act' mask store holds (4), so the backward
    chains re-read the sign bits from the activation images; tensor-core vs fp32 FFMA backward."""
    rng = np.random.default_rng(29)
    g = H.random_batch(rng, 30, 5, 25, D=14)
    params = O.make_params(4, 1, 14, 128, 7, last_layer_scale=0.1)
    net = H.make_grevnet(params, 128, 7, device=DEV)
    dg = dev_graph(g)
    n = g.nodes.shape[0]
    z, _ = net.f64(dg)
    ref = net.backward_from_z(dg, z.nodes, 1.0 / n, math="fp32").double().cpu().numpy()
    got = net.backward_from_z(dg, z.nodes, 1.0 / n, math="tc3x").double().cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= BWD_TOL["tc3x"][0] * np.abs(ref).max()
    assert float(got @ ref) / (np.linalg.norm(got) * np.linalg.norm(ref)) > 1 - 1e-6


@pytest.mark.parametrize("bmath", ["tc3x", "bf16"])
def test_f2_tensor_core_backward_many_tiles(bmath):
    """More tiles than SMs (every CTA walks several 128-node tiles; the weight-gradient GEMM splits
    over node ranges) at the bench shape L=256, K=5: tensor-core backward vs the fp32 FFMA backward
    on the same device, and determinism (fixed-order reductions -> bit-identical reruns)."""
    rng = np.random.default_rng(23)
    g = H.random_batch(rng, 1200, 10, 30, p_edge=0.2, D=14)
    params = O.make_params(5, 2, 14, 256, 5, last_layer_scale=0.1)
    net = H.make_grevnet(params, 256, 5, device=DEV)
    dg = dev_graph(g)
    n = g.nodes.shape[0]
    assert n > 148 * 128
    z, _ = net.f64(dg)
    ref, x_ref = net.backward_from_z(dg, z.nodes, 1.0 / n, return_x=True, math="fp32")
    got, x_got = net.backward_from_z(dg, z.nodes, 1.0 / n, return_x=True, math=bmath)
    again = net.backward_from_z(dg, z.nodes, 1.0 / n, math=bmath)
    ref, got = ref.double().cpu().numpy(), got.double().cpu().numpy()
    tol, ctol = BWD_TOL[bmath]
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= tol * np.abs(ref).max()
    assert float(got @ ref) / (np.linalg.norm(got) * np.linalg.norm(ref)) > 1 - ctol
    assert torch.equal(again.cpu(), torch.from_numpy(got).float())
    assert (x_got - x_ref).abs().max().item() < 5e-5
    assert np.abs(x_got.cpu().numpy() - g.nodes).max() < 5e-5


def test_f2_weight_gradient_gemm():
    """k_dw_tc on its own (gnf_debug_dw_gemm): a^T b over the node dimension from the MN-major bf16
    hi/lo tile images; split parts=2 is fp32-class (2^-17 operands), parts=1 is one bf16 rounding."""
    from graph_normalizing_flows_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator(device="cpu").manual_seed(3)
    for (fa, fb, n, parts, splits, tol) in [(256, 256, 128, 2, 1, 3e-5), (256, 256, 128, 1, 1, 8e-3),
                                            (256, 16, 300, 2, 2, 3e-5), (128, 128, 1000, 2, 3, 3e-5),
                                            (128, 16, 77, 1, 1, 8e-3), (256, 256, 5000, 2, 7, 3e-5)]:
        a = torch.randn(n, fa, generator=gen).to(DEV)
        b = (torch.randn(n, fb, generator=gen) * 1e-3).to(DEV)
        out = torch.empty(fa, fb, device=DEV)
        wsb = 4 * ((n + 127) // 128) * 128 * (fa + fb) + 4 * splits * fa * fb + 4096
        ws = _lib.workspace(wsb, torch.device(DEV))
        _lib.check(lib.gnf_debug_dw_gemm(_lib.ptr(a), _lib.ptr(b), n, fa, fb, parts, splits, _lib.ptr(out),
                                         _lib.ptr(ws), wsb, _lib.stream_ptr(torch.device(DEV))), "gnf_debug_dw_gemm")
        ref = a.double().T @ b.double()
        assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol, (fa, fb, n, parts)


def test_f2_training_steps_reduce_the_loss():
    """A few Adam steps on one batch (optimiser flags of run_grevnet.py:345-377: lr 1e-4 default;
    a larger lr here) lower loss_per_node; parameters are re-packed after every step."""
    rng = np.random.default_rng(22)
    g = H.random_batch(rng, 16, 6, 20, D=4)
    g = g._replace(nodes=(g.nodes * 2.0 + 1.0).astype(np.float32))
    params = O.make_params(9, 2, 4, 128, 3, last_layer_scale=0.1)
    net = H.make_grevnet(params, 128, 3, device=DEV)
    dg = dev_graph(g)
    opt = torch.optim.Adam([net.params], lr=3e-3, betas=(0.9, 0.9), eps=1e-8)
    losses = []
    for _ in range(12):
        out, _ = net.loss_and_grad(dg, per_node=True)
        losses.append(float(out["loss_per_node"]))
        opt.step()
    assert losses[-1] < losses[0] - 0.05, losses


def test_run_grevnet_port_trains_with_reference_defaults(tmp_path, caplog):
    """The trainer port with the reference's default GNN and bijector (dm_self_attn, use_batch_norm=True,
    run_grevnet.py:56,83): f1 forward/backward + batch-norm backward + Adam + gamma constraint run end to end and
    write the checkpoint."""
    import logging
    from absl import flags
    from graph_normalizing_flows_b200 import run_grevnet as RG
    argv = ["run_grevnet", "--num_train_iters=12", "--train_batch_size=8", "--num_coupling_layers=2",
            "--gnn_latent_dim=64", "--gnn_num_layers=3", "--log_every_n_steps=4", "--lr=1e-3",
            f"--logdir={tmp_path}", "--dataset=mog_4", "--last_layer_init_scale=0.1"]
    flags.FLAGS.unparse_flags()
    flags.FLAGS(argv)
    assert flags.FLAGS.make_gnn_fn == "dm_self_attn" and flags.FLAGS.use_batch_norm
    with caplog.at_level(logging.INFO):
        assert RG.main([]) == 0
    assert os.path.exists(os.path.join(str(tmp_path), "grevnet_12.pt"))
    losses = [float(r.getMessage().split("loss_per_node ")[1].split()[0]) for r in caplog.records
              if "loss_per_node" in r.getMessage()]
    assert len(losses) >= 3 and all(np.isfinite(losses)), losses


def test_f3_decode_tail_pred_adj():
    """Row f3: pred_adj(graph, scaled_hacky_sigmoid_l2) per graph block vs the dense reference
    formulation (loss.py:154-159,45-53); threshold 0.5 -> networkx graphs as
    train_grevnet_with_data.py:529-548."""
    from graph_normalizing_flows_b200 import train_grevnet_with_data as TD
    rng = np.random.default_rng(31)
    n_node = np.array([5, 1, 12, 7, 30], np.int32)
    for d in (2, 14, 200):
        nodes = (rng.standard_normal((int(n_node.sum()), d)) * 0.6).astype(np.float32)
        g = TD.transform_example(nodes, n_node).to(DEV)
        blocks, adj_off = G.loss.pred_adj(g)
        dense = G.loss.dense_pred_adj(blocks, adj_off, n_node).cpu().numpy()
        want = O.pred_adj(nodes.astype(np.float64), n_node)
        assert dense.shape == want.shape and np.abs(dense - want).max() < 2e-5
        assert not np.diag(dense).any()
    graphs = G.loss.sampled_graphs(blocks, adj_off, n_node)
    assert [h.number_of_nodes() for h in graphs] == n_node.tolist()
    adj = (want > 0.5)
    lo = 0
    for h, n in zip(graphs, n_node):
        import networkx as nx
        assert np.array_equal(nx.to_numpy_array(h) > 0, adj[lo:lo + n, lo:lo + n])
        lo += n


def test_f3_f4_sampling_pipeline_on_fc_graphs():
    """z ~ N(0,I) on fully connected graphs (utils.senders_receivers) -> g -> pred_adj -> graphs."""
    from graph_normalizing_flows_b200 import train_grevnet_with_data as TD
    params = O.make_params(4, 2, 14, 128, 3, agg="mean", block="concat", last_layer_scale=0.05)
    net = H.make_grevnet(params, 128, 3, device=DEV)
    gen = torch.Generator().manual_seed(3)
    graphs, lp = TD.sample_graphs(net, [6, 11, 20], device=DEV, generator=gen)
    assert [h.number_of_nodes() for h in graphs] == [6, 11, 20] and len(lp) == 3 and all(np.isfinite(lp))
    # same latent through the oracle
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(37, 14, generator=gen).numpy()
    s, r = G.utils.senders_receivers([6, 11, 20])
    x = O.grevnet_g(z, s, r, params)
    want = O.pred_adj(x.astype(np.float64), [6, 11, 20]) > 0.5
    import networkx as nx
    lo = 0
    agree = total = 0
    for h, n in zip(graphs, [6, 11, 20]):
        a = nx.to_numpy_array(h) > 0
        agree += int((a == want[lo:lo + n, lo:lo + n]).sum())
        total += n * n
        lo += n
    assert agree >= total - 2          # thresholding: allow a borderline entry


def test_batch_prefetcher_stages_and_validates():
    rng = np.random.default_rng(41)
    g = H.random_batch(rng, 12, 5, 30, D=14)
    params = O.make_params(4, 2, 14, 128, 3, last_layer_scale=0.05)
    net = H.make_grevnet(params, 128, 3, device=DEV)
    direct = G.loss.log_prob(net, dev_graph(g))
    host = G.GraphsTuple(*[torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if v is not None else None for v in g])
    pf = G.graphs.BatchPrefetcher(DEV)
    t1 = pf.submit(host)
    t2 = pf.submit(host)                       # two batches in flight
    for t in (t1, t2):
        dg = pf.wait(t)
        assert getattr(dg.senders, "_gnf_structure", None) is not None      # CSR already staged
        out = G.loss.log_prob(net, dg)
        assert float(out["log_prob_xs"]) == float(direct["log_prob_xs"])
    bad = g._replace(senders=g.senders.copy())
    bad.senders[0] = g.nodes.shape[0] + 5
    with pytest.raises(ValueError, match="outside"):
        pf.wait(pf.submit(G.GraphsTuple(*bad)))
    with pytest.raises(ValueError, match="n_node"):
        pf.submit(G.GraphsTuple(*g._replace(n_node=g.n_node + 1)))


def _tame_layer_norm(params):
    """LayerNorm makes s unit-variance whatever the last-layer scale: shrink gamma so exp(s) stays well conditioned
    for the fp32 round trip (same purpose as last_layer_scale)."""
    for which in ("s", "t"):
        for half in params[which]:
            for gnn in (half if isinstance(half, list) else [half]):
                if isinstance(gnn, dict) and "ln_gamma" in gnn:
                    gnn["ln_gamma"] = (gnn["ln_gamma"] * 0.2).astype(gnn["ln_gamma"].dtype)


@pytest.mark.parametrize("variant", ["default_d2_fc", "d14_sparse_noconcat_residual", "kq_division_shared", "layer_norm"])
def test_f1_dm_self_attn_gnn(variant):
    """Row f1: DMSelfAttentionMLP (gnn.py:385-573), the default GNN of both scripts, inside the flow.
    fp32 kernels vs the oracle restatement (graph_nets segment softmax [upstream])."""
    rng = np.random.default_rng(51)
    if variant == "default_d2_fc":          # run_grevnet.py defaults: D=2, kq=v=10, 8 heads, 80, concat, FC graphs
        D, T, L, K, ws = 2, 3, 256, 5, False
        attn = dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False, kq_dim_division=False)
        n_node = rng.integers(4, 40, size=6)
        s, r = G.utils.senders_receivers(n_node)
        nodes = rng.standard_normal((int(n_node.sum()), D)).astype(np.float32)
        g = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
    elif variant == "d14_sparse_noconcat_residual":
        D, T, L, K, ws = 14, 1, 64, 3, False
        attn = dict(num_heads=3, kq_dim=5, v_dim=7, out_dim=12, concat=False, residual=True, kq_dim_division=False)
        g = H.random_batch(rng, 9, 4, 30, D=D, isolated=True)      # isolated receivers: attention output 0
        g = g._replace(nodes=(g.nodes * 0.1).astype(np.float32))  # residual adds x to s: keep exp(s) tame
    elif variant == "layer_norm":            # snt.LayerNorm on the GNN output (gnn.py:554-556), with the residual
        D, T, L, K, ws = 10, 2, 64, 3, False
        attn = dict(num_heads=2, kq_dim=6, v_dim=5, out_dim=9, concat=True, residual=True, kq_dim_division=False,
                    layer_norm=True)
        g = H.random_batch(rng, 8, 4, 25, D=D)
        g = g._replace(nodes=(g.nodes * 0.3).astype(np.float32))
    else:
        D, T, L, K, ws = 6, 2, 128, 3, True
        attn = dict(num_heads=4, kq_dim=16, v_dim=8, out_dim=20, concat=True, residual=False, kq_dim_division=True)
        g = H.random_batch(rng, 7, 4, 25, D=D)
    params = O.make_params(13, T, D, L, K, block="dm_attn", act="relu", attn=attn, last_layer_scale=0.1,
                           weight_sharing=ws)
    _tame_layer_norm(params)
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
    net = H.make_grevnet(params, L, K, device=DEV)
    dg = dev_graph(g)
    assert np.isfinite(z64).all()
    # round 2: without residual / layer_norm and with L in {128, 256} the attention GNN runs its layers 1..K-1 and the
    # coupling update in the fused tcgen05 kernel (layer 0 and the projections in k_linear_tc); the residual / layer_norm
    # variants run their MLP layers one by one in k_gemm_tc and the projections in k_linear_tc (layered path)
    assert net.math == "tc3x"
    for math in ["tc3x", "tc3x_bf16", "fp32"]:
        net.math = math
        out = G.loss.log_prob(net, dg, return_z=True)
        assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4 * max(1.0, np.abs(z64).max()), math
        assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL, math
        x_back = net(out["z"], inverse=False).nodes.cpu().numpy()
        assert np.abs(x_back - g.nodes).max() < 1e-4, math
    net.check_numerics()
    net.math = None
    # one attention GNN on its own: callable GraphsTuple -> GraphsTuple
    gnn1 = net.s[0] if ws else net.s[0][0]
    half = dg.replace(nodes=dg.nodes[:, :D // 2].contiguous())
    got = gnn1(half).nodes.cpu().numpy()
    ref = O.node_block_gnn(g.nodes[:, :D // 2].astype(np.float64), g.senders, g.receivers,
                           O.cast_params(params, np.float64)["s"][0] if ws else O.cast_params(params, np.float64)["s"][0][0],
                           params["cfg"])
    assert np.abs(got - ref).max() < 1e-4


@pytest.mark.parametrize("variant", ["concat_kqdiv", "noconcat_residual_shared", "layer_norm"])
def test_f1_dm_self_attn_backward_matches_autograd(variant):
    """Row f1 + f2: reversible backward THROUGH the attention GNN (fp32 kernels: k_attn_bwd_recv / k_attn_bwd_send,
    projection dW/dX GEMMs) vs torch autograd of the fp64 restatement.  Tolerance as the fp32 MLP backward (2e-4
    of the max-norm), cosine > 1 - 1e-6; the input is reconstructed on the way."""
    from oracle import gnf_oracle_torch as OT
    rng = np.random.default_rng(53)
    if variant == "concat_kqdiv":
        D, T, L, K, ws = 6, 2, 64, 3, False
        attn = dict(num_heads=4, kq_dim=16, v_dim=8, out_dim=20, concat=True, residual=False, kq_dim_division=True)
        g = H.random_batch(rng, 8, 4, 25, D=D, isolated=True)
    elif variant == "layer_norm":
        D, T, L, K, ws = 10, 2, 64, 3, False
        attn = dict(num_heads=2, kq_dim=6, v_dim=5, out_dim=9, concat=True, residual=True, kq_dim_division=False,
                    layer_norm=True)
        g = H.random_batch(rng, 8, 4, 25, D=D)
        g = g._replace(nodes=(g.nodes * 0.3).astype(np.float32))
    else:
        D, T, L, K, ws = 14, 2, 32, 4, True
        attn = dict(num_heads=3, kq_dim=5, v_dim=7, out_dim=12, concat=False, residual=True, kq_dim_division=False)
        g = H.random_batch(rng, 9, 4, 30, D=D)
        g = g._replace(nodes=(g.nodes * 0.1).astype(np.float32))
    params = O.make_params(17, T, D, L, K, block="dm_attn", act="leaky_relu", attn=attn, last_layer_scale=0.1,
                           weight_sharing=ws)
    _tame_layer_norm(params)
    n = g.nodes.shape[0]
    net = H.make_grevnet(params, L, K, device=DEV)
    dg = dev_graph(g)
    for per_node in (True, False):
        scale = 1.0 / n if per_node else 1.0
        loss_ref, grad_ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, scale)
        out, grads = net.loss_and_grad(dg, per_node=per_node)
        got = grads.double().cpu().numpy()
        loss = float(out["loss_per_node"] if per_node else out["total_loss"])
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
        assert got.shape == grad_ref.shape and np.isfinite(got).all()
        assert np.abs(got - grad_ref).max() <= 2e-4 * np.abs(grad_ref).max()
        assert float(got @ grad_ref) / (np.linalg.norm(got) * np.linalg.norm(grad_ref)) > 1 - 1e-6
    _, x_rec = net.backward_from_z(dg, out["z"].nodes, 1.0, return_x=True)
    assert np.abs(x_rec.cpu().numpy() - g.nodes).max() < 5e-5


def test_cuda_graph_replay_matches_eager():
    rng = np.random.default_rng(61)
    g = H.random_batch(rng, 10, 5, 30, D=14)
    params = O.make_params(4, 3, 14, 256, 5, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device=DEV)
    dg = dev_graph(g)
    eager = G.loss.log_prob(net, dg, return_z=True)
    runner = G.loss.GraphedLogProb(net, dg)
    vec = runner()
    assert float(vec[2]) == float(eager["log_prob_xs"]) and torch.equal(runner.z, eager["z"].nodes)
    new_nodes = dg.nodes * 0.5 + 0.1                          # new features, same structure
    vec2 = runner(new_nodes).clone()
    eager2 = G.loss.log_prob(net, dg.replace(nodes=new_nodes))
    assert float(vec2[2]) == float(eager2["log_prob_xs"])


def test_empty_and_tiny_batches():
    params = O.make_params(4, 2, 14, 128, 3, last_layer_scale=0.05)
    net = H.make_grevnet(params, 128, 3, device=DEV)
    empty = G.GraphsTuple(torch.zeros(0, 14, device=DEV), None, torch.zeros(0, dtype=torch.int32, device=DEV),
                          torch.zeros(0, dtype=torch.int32, device=DEV), None,
                          torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV))
    out = G.loss.log_prob(net, empty)
    assert float(out["log_prob_xs"]) == 0.0 and float(out["num_nodes"]) == 0.0
    one = O.GraphsTuple(np.ones((1, 14), np.float32), None, np.zeros(1, np.int32), np.zeros(1, np.int32), None,
                        np.ones(1, np.int32), np.ones(1, np.int32))
    z, ldj = O.grevnet_f(one.nodes, one.senders, one.receivers, params)
    got = G.loss.log_prob(net, dev_graph(one))
    assert H.rel_err(got["log_prob_xs"], O.log_prob(z, ldj, one.n_node)["log_prob_xs"]) < LOGPROB_RTOL
    # no edges at all (segment ids never hit): agg = 0 everywhere
    noedge = O.GraphsTuple(np.ones((5, 14), np.float32), None, np.zeros(0, np.int32), np.zeros(0, np.int32), None,
                           np.full(1, 5, np.int32), np.zeros(1, np.int32))
    z, ldj = O.grevnet_f(noedge.nodes, noedge.senders, noedge.receivers, params)
    got = G.loss.log_prob(net, dev_graph(noedge))
    assert H.rel_err(got["log_prob_xs"], O.log_prob(z, ldj, noedge.n_node)["log_prob_xs"]) < LOGPROB_RTOL
    # the training-step entry point on the same edge cases: empty batch -> zero gradient, one node / no edges ->
    # the tensor-core backward (a single, mostly padded tile) agrees with the fp32 FFMA backward
    _, g0 = net.loss_and_grad(empty)
    assert float(g0.abs().max()) == 0.0
    for tiny in (one, noedge):
        dg = dev_graph(tiny)
        zt, _ = net.f64(dg)
        n = tiny.nodes.shape[0]
        ref = net.backward_from_z(dg, zt.nodes, 1.0 / n, math="fp32").double().cpu().numpy()
        got = net.backward_from_z(dg, zt.nodes, 1.0 / n, math="tc3x").double().cpu().numpy()
        assert np.isfinite(got).all() and np.abs(got - ref).max() <= BWD_TOL["tc3x"][0] * np.abs(ref).max()


def test_property_random_batches():
    """K8: random small batches, D in {2,4,14}, T in {1,2}, every factory, both kernel families."""
    rng = np.random.default_rng(8)
    for trial in range(12):
        D = int(rng.choice([2, 4, 14]))
        T = int(rng.choice([1, 2]))
        L = int(rng.choice([128, 256]))
        K = int(rng.choice([2, 3, 5]))
        block = str(rng.choice(["concat", "agg_then"]))
        agg = str(rng.choice(["sum", "mean"]))
        g = H.random_batch(rng, int(rng.integers(1, 12)), 1, 35, p_edge=float(rng.uniform(0.05, 0.6)), D=D,
                           isolated=bool(rng.integers(0, 2)))
        params = O.make_params(int(rng.integers(1 << 30)), T, D, L, K, agg=agg, block=block, eps=0.9,
                               last_layer_scale=0.1)
        p64 = O.cast_params(params, np.float64)
        z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, p64)
        want = O.log_prob(z64, ldj64, g.n_node)["log_prob_xs"]
        for math in ("fp32", "tc3x"):
            net = H.make_grevnet(params, L, K, device=DEV, math=math)
            out = G.loss.log_prob(net, dev_graph(g), return_z=True)
            assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL, (trial, math)
            assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4, (trial, math)


# ----------------------------------------------------------- BASELINE full size: properties -----
def _community_batch(n_graphs, seed):
    from graph_normalizing_flows_b200 import graph_data as GD
    npz = np.load(os.path.join(H.GOLDEN, "graphs_community_medium_4_128.npz"))
    ds = GD.GraphDataset(None, 14, structures=GD.structures_from_fixture(npz))
    return ds.draw_batch(n_graphs, np.random.default_rng(seed))


def test_full_size_community_medium_properties():
    """BASELINE configs[1] at full size (B=4096, T=6, D=14, L=256, K=5): round trip, shard
    additivity, permutation invariance of the batch scalars, fp32-vs-tc3x agreement."""
    host = _community_batch(4096, 12345)
    params = O.make_params(12345, 6, 14, 256, 5, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device=DEV, math="tc3x")
    dg = host.to(DEV)
    full = G.loss.log_prob(net, dg, return_z=True)
    assert torch.isfinite(full["z"].nodes).all()
    x_back = net(full["z"], inverse=False).nodes
    assert float((x_back - dg.nodes).abs().max()) < 1e-4
    # two contiguous shards: scalars add up (what the NCCL all-reduce relies on)
    half = 2048
    ids = np.arange(4096)
    tot = np.zeros(3)
    for part in (ids[:half], ids[half:]):
        sh = sharding.shard_graphs_tuple(host, part).to(DEV)
        o = G.loss.log_prob(net, sh)
        tot += np.array([float(o["log_prob_zs"]), float(o["log_det_jacobian"]), float(o["log_prob_xs"])])
    want = np.array([float(full["log_prob_zs"]), float(full["log_det_jacobian"]), float(full["log_prob_xs"])])
    assert np.all(np.abs(tot - want) <= 1e-8 * np.abs(want) + 1e-6)
    # independent arithmetic (fp32 FFMA kernels) agrees with the tensor-core path
    net32 = H.make_grevnet(params, 256, 5, device=DEV, math="fp32")
    o32 = G.loss.log_prob(net32, dg, return_z=True)
    assert H.rel_err(full["log_prob_xs"], o32["log_prob_xs"]) < LOGPROB_RTOL
    assert float((o32["z"].nodes - full["z"].nodes).abs().max()) < 1e-4


def _family_batch(fam, b, seed):
    from graph_normalizing_flows_b200 import graph_data as GD
    npz = np.load(os.path.join(H.GOLDEN, f"graphs_{fam}.npz"))
    ds = GD.GraphDataset(None, 14, structures=GD.structures_from_fixture(npz))
    return ds.draw_batch(b, np.random.default_rng(seed))


def test_config2_grid_12_step_bf16():
    """BASELINE configs[2]: grid graphs, 12-step GRevNet, bf16 single-pass fused kernel.  Stated
    tolerance for this mode: 1e-3 relative on log-prob, 5e-2 absolute on z (vs the fp64 oracle);
    the tc3x mode on the same inputs keeps 1e-5."""
    host = _family_batch("grid_4_128", 4, 3)
    params = O.make_params(12345, 12, 14, 256, 5, last_layer_scale=0.02)
    z64, ldj64 = O.grevnet_f(host.nodes.astype(np.float64), host.senders, host.receivers,
                             O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, host.n_node)["log_prob_xs"]
    dg = host.to(DEV)
    bf = G.loss.log_prob(H.make_grevnet(params, 256, 5, device=DEV, math="bf16"), dg, return_z=True)
    assert H.rel_err(bf["log_prob_xs"], want) < 1e-3
    assert np.abs(bf["z"].nodes.cpu().numpy() - z64).max() < 5e-2
    tc = G.loss.log_prob(H.make_grevnet(params, 256, 5, device=DEV, math="tc3x"), dg, return_z=True)
    assert H.rel_err(tc["log_prob_xs"], want) < LOGPROB_RTOL
    x_back = H.make_grevnet(params, 256, 5, device=DEV, math="bf16")(bf["z"], inverse=False).nodes
    assert float((x_back - dg.nodes).abs().max()) < 5e-2


def test_config3_protein_sharded_over_four_ranks():
    """BASELINE configs[3]: protein batch, cost-balanced over 4 (virtual) ranks; the reduced scalars
    equal the unsharded ones and the oracle's."""
    host = _family_batch("protein_4_128", 64, 4)
    params = O.make_params(12345, 2, 14, 256, 5, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device=DEV, math="tc3x")
    full = G.loss.log_prob(net, host.to(DEV))
    parts = sharding.partition_graphs(host.n_node, host.n_edge, 4)
    cost = sharding.graph_costs(host.n_node, host.n_edge)
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.25
    vec = torch.zeros(4, dtype=torch.float64, device=DEV)
    for ids in parts:
        sh = sharding.shard_graphs_tuple(host, ids).to(DEV)
        z, ldj = net.f64(sh)
        vec += G.loss.mvn_log_prob_sum(z.nodes, ldj)          # what the NCCL all-reduce(SUM) adds up
    assert H.rel_err(vec[2], full["log_prob_xs"]) < 1e-9
    z64, ldj64 = O.grevnet_f(host.nodes.astype(np.float64), host.senders, host.receivers,
                             O.cast_params(params, np.float64))
    assert H.rel_err(vec[2], O.log_prob(z64, ldj64, host.n_node)["log_prob_xs"]) < LOGPROB_RTOL


def test_config4_mixed_family_batch():
    """BASELINE configs[4]: citeseer + a mixed batch with equal draws from all five 4_128 families."""
    from graph_normalizing_flows_b200.graphs import concat_structures
    from graph_normalizing_flows_b200 import graph_data as GD
    rng = np.random.default_rng(5)
    structs = []
    for fam in ("community_medium_4_128", "grid_4_128", "protein_4_128", "citeseer_4_128", "caveman_4_128"):
        npz = np.load(os.path.join(H.GOLDEN, f"graphs_{fam}.npz"))
        all_s = GD.structures_from_fixture(npz)
        structs += [all_s[i] for i in rng.integers(0, len(all_s), size=3)]
    order = rng.permutation(len(structs))
    structs = [structs[i] for i in order]
    n = sum(s[0] for s in structs)
    host = concat_structures(structs, nodes=rng.standard_normal((n, 14)).astype(np.float32))
    params = O.make_params(12345, 2, 14, 256, 5, last_layer_scale=0.05)
    z64, ldj64 = O.grevnet_f(host.nodes.astype(np.float64), host.senders, host.receivers,
                             O.cast_params(params, np.float64))
    want = O.log_prob(z64, ldj64, host.n_node)["log_prob_xs"]
    out = G.loss.log_prob(H.make_grevnet(params, 256, 5, device=DEV, math="tc3x"), host.to(DEV), return_z=True)
    assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL
    assert np.abs(out["z"].nodes.cpu().numpy() - z64).max() < 1e-4


def test_oracle_parity_on_real_family_batches():
    """Real graph families at sizes the oracle finishes in seconds."""
    for fam, b, T in (("community_medium_4_128", 48, 6), ("grid_4_128", 6, 3), ("protein_4_128", 24, 2),
                      ("citeseer_4_128", 6, 2), ("caveman_4_128", 8, 2)):
        from graph_normalizing_flows_b200 import graph_data as GD
        npz = np.load(os.path.join(H.GOLDEN, f"graphs_{fam}.npz"))
        ds = GD.GraphDataset(None, 14, structures=GD.structures_from_fixture(npz))
        host = ds.draw_batch(b, np.random.default_rng(7))
        params = O.make_params(12345, T, 14, 256, 5, last_layer_scale=0.05)
        p64 = O.cast_params(params, np.float64)
        z64, ldj64 = O.grevnet_f(host.nodes.astype(np.float64), host.senders, host.receivers, p64)
        want = O.log_prob(z64, ldj64, host.n_node)["log_prob_xs"]
        agg = O.aggregate(np.ascontiguousarray(host.nodes[:, :7]), host.senders, host.receivers, "sum")
        dg = host.to(DEV)
        st = G.graphs.structure_of(dg)
        got_agg = G.gnn.gather_segment_reduce(dg.nodes[:, :7].contiguous(), st, "sum")
        assert np.array_equal(got_agg.cpu().numpy(), agg), fam
        for math in ("tc3x", "fp32"):
            net = H.make_grevnet(params, 256, 5, device=DEV, math=math)
            out = G.loss.log_prob(net, dg)
            assert H.rel_err(out["log_prob_xs"], want) < LOGPROB_RTOL, (fam, math)
