"""CPU suite: the C-ABI library loads and exports every symbol include/gnf_b200.h declares
(no compute calls without a GPU), and the host-side mirror of the reference interface behaves
(construction, parameter layout, error contract, no CPU fallback)."""
import ctypes as C
import os
import re
from functools import partial

import numpy as np
import pytest
import torch

import helpers as H
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib


def header_symbols():
    src = open(os.path.join(H.ROOT, "include", "gnf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/gnf_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.gnf_abi_version() == 5


def test_param_count_matches_reference_formula():
    lib = _lib.load()
    for (T, D, L, K, block, ws) in [(6, 14, 256, 5, "concat", 0), (12, 2, 256, 5, "agg_then", 0), (3, 14, 128, 4, "concat", 1)]:
        d = _lib.FlowDesc(num_timesteps=T, node_embedding_dim=D, latent_dim=L, num_layers=K, agg=0,
                          block=_lib.BLOCK[block], act=0, weight_sharing=ws, eps=1.0)
        Hh = D // 2
        inn = D if block == "concat" else Hh
        per = inn * L + (K - 2) * L * L + L * Hh + (K - 1) * L + Hh
        assert lib.gnf_flow_param_count(C.byref(d)) == per * 4 * (1 if ws else T)
    bad = _lib.FlowDesc(num_timesteps=1, node_embedding_dim=3, latent_dim=8, num_layers=2)
    assert lib.gnf_flow_param_count(C.byref(bad)) == -1
    assert b"even" in lib.gnf_last_error()
    assert lib.gnf_padded_half(7) == 8 and lib.gnf_padded_half(1) == 4 and lib.gnf_padded_half(100) == 100


def mk(L=256, D=14, K=5, act=G.gnn.leaky_relu, fac=G.sum_concat_then_mlp_gnn, **kw):
    return lambda: fac(partial(G.make_mlp_model, L, D / 2, K, act, 0.1, 0.1), **kw)


def test_grevnet_mirrors_reference_constructor():
    net = G.GRevNet(mk(), 6, 14, use_batch_norm=False, weight_sharing=False, seed=12345, device="cpu")
    assert net.params.numel() == 203015 * 24
    assert len(net.s) == 2 and len(net.s[0]) == 6 and len(net.t[1]) == 6       # s[2][T], t[2][T] gnn.py:292-299
    m = net.mlp_of("s", 0, 0)
    assert [tuple(w.shape) for w in m.weights] == [(14, 256), (256, 256), (256, 256), (256, 256), (256, 7)]
    # glorot trunc-normal weights (|w| <= 2 std), trunc-normal(0.1) biases (gnn.py:171-174)
    std = np.sqrt(2.0 / (256 + 256)) / 0.87962566103423978
    assert float(m.weights[1].abs().max()) <= 2 * std + 1e-6
    assert float(m.biases[1].abs().max()) <= 0.2 + 1e-6
    assert abs(float(m.weights[1].std()) - std * 0.8796) < 0.01
    # views alias the flat buffer handed to the C ABI, flat order which -> half -> step
    assert m.weights[0].data_ptr() == net.params.data_ptr()
    t10 = net.mlp_of("t", 1, 0)
    assert t10.weights[0].data_ptr() == net.params.data_ptr() + 4 * 203015 * (3 * 6)
    shared = G.GRevNet(mk(), 6, 14, weight_sharing=True, device="cpu")
    assert shared.params.numel() == 203015 * 4 and len(shared.s) == 2


def test_scale_last_layers():
    net = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, seed=1, device="cpu")
    w = net.mlp_of("t", 1, 1).weights[-1].clone()
    net.scale_last_layers_(0.05)
    assert torch.allclose(net.mlp_of("t", 1, 1).weights[-1], w * 0.05)


def test_parameter_edits_invalidate_the_packed_weights():
    """The device-side packed weights are refreshed whenever the flat parameter's in-place version
    changes: optimiser steps, copy_, and edits through the per-MLP views all bump it."""
    net = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, seed=1, device="cpu")
    v0 = net.params._version
    net.scale_last_layers_(0.5)
    v1 = net.params._version
    assert v1 > v0
    net.params.grad = torch.ones_like(net.params)
    torch.optim.SGD([net.params], lr=0.1).step()
    assert net.params._version > v1 and net.params.detach()._version == net.params._version


def test_error_contract_on_host():
    with pytest.raises(ValueError, match="even"):
        G.GRevNet(mk(D=3), 2, 3, device="cpu")                         # tf.split failure at gnn.py:306
    bn = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, use_batch_norm=True, device="cpu")   # bns[2][T], gnn.py:301-302
    assert bn.bn_gamma.shape == (2, 2, 2) and float(bn.bn_moving_var.min()) == 1.0
    with pytest.raises(ValueError):
        G.make_mlp_model(8, 2, 3, activation="tanh")
    with pytest.raises(ValueError):
        G.GRevNet(mk(D=6), 2, 14, device="cpu")                        # MLP output_dim != D/2
    net = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, device="cpu")
    g = H.random_batch(np.random.default_rng(0), 2, 3, 5, D=4)
    cpu_graph = G.GraphsTuple(*[torch.from_numpy(v) if v is not None else None for v in g])
    with pytest.raises(RuntimeError, match="CUDA"):                       # no CPU fallback, loudly
        net(cpu_graph, inverse=True)
    with pytest.raises(TypeError):
        net(G.GraphsTuple(*g), inverse=True)                               # numpy arrays: not device tensors


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Round-1 additions (training-step entry points): bad arguments are rejected on the host, before any CUDA
    call, with the documented codes and a message in gnf_last_error."""
    lib = _lib.load()
    null = C.c_void_p(0)
    assert lib.gnf_grevnet_backward(null, null, 0, 0, null, null, null, null, 1.0, null, null, 1, null, 0, null) == _lib.GNF_EINVAL
    assert b"null flow" in lib.gnf_last_error()
    assert lib.gnf_coupling_half_backward(null, 0, 0, null, null, null, null, 0, 0, null, null, null, null, 1.0, null, 1,
                                          null, 0, null) == _lib.GNF_EINVAL
    assert lib.gnf_grevnet_backward_workspace(null, 10, 1) == 0
    assert lib.gnf_bn_backward_apply(null, null, -1, 4, null, null) == _lib.GNF_EINVAL
    assert lib.gnf_bn_finalize(null, 4, null, null, 1e-3, 1.0, null, null, null, null, null, -1.0, null) == _lib.GNF_EINVAL
    assert lib.gnf_bn_backward_coef(null, null, 4, null, null, 1e-3, 1.0, null, null, null, null, null) == _lib.GNF_EINVAL
    assert lib.gnf_debug_dw_gemm(null, null, 0, 256, 256, 2, 1, null, null, 0, null) == _lib.GNF_EINVAL
    assert lib.gnf_debug_bwd_layout(null, 10, null) == _lib.GNF_EINVAL
    assert lib.gnf_debug_kernel_timing(0) == _lib.GNF_OK        # switching the (unused) event brackets off is a no-op
    # round-2 additions: chained batch-norm entries, the layered tensor-core Linear on its own
    assert lib.gnf_grevnet_bn_workspace(null) == 0
    assert lib.gnf_grevnet_forward_bn(null, null, 0, 0, null, null, null, null, null, null, 1e-3, 0.99, null, null, null, 1,
                                      null, 0, null, 0, null) == _lib.GNF_EINVAL
    assert b"null flow" in lib.gnf_last_error()
    assert lib.gnf_grevnet_inverse_bn(null, null, 0, 0, null, null, null, null, null, null, 1e-3, null, 1, null, 0, null, 0,
                                      null) == _lib.GNF_EINVAL
    # image = K padded to 16 x N padded to column blocks of <= 256, hi + lo of 2 bytes, in two element types, + job scratch
    assert lib.gnf_debug_linear_tc_workspace(2048, 2048) == 2 * 2048 * 2048 * 4 + 256
    assert lib.gnf_debug_linear_tc_workspace(164, 300) == 2 * 176 * 512 * 4 + 256
    assert lib.gnf_debug_linear_tc_workspace(40, 100) == 2 * 48 * 112 * 4 + 256
    assert lib.gnf_debug_linear_tc_workspace(0, 4) == 0
    assert lib.gnf_debug_linear_tc(null, null, null, 4, 8, 8, 2, 1, null, null, 0, null) == _lib.GNF_EINVAL
    assert lib.gnf_debug_linear_tc(C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), 4, 6, 8, 2, 1, C.c_void_p(256), null, 0,
                                   null) == _lib.GNF_EINVAL            # k not a multiple of 4


def test_attention_gnn_parameter_layout_and_layer_norm():
    """dm_self_attn_gnn: Wq Wk Wv Wo, the MLP, then (layer_norm=True) the LayerNorm gamma = 1, beta = 0
    (snt.LayerNorm variables are created after the MLP's, gnn.py:547-556); the C ABI counts the same."""
    lib = _lib.load()
    D, Hh, L, K = 6, 3, 16, 3
    mlp_fn = partial(G.make_mlp_model, L, D / 2, K, G.gnn.relu, 0.1, 0.1)
    for ln in (False, True):
        mk_attn = lambda: G.dm_self_attn_gnn(5, 4, mlp_fn, 2, 7, concat=True, residual=True, layer_norm=ln)
        net = G.GRevNet(mk_attn, 2, D, seed=3, device="cpu")
        attn = 2 * Hh * 10 + Hh * 4 + 8 * 7
        mlp = (Hh + 7) * L + L + L * L + L + L * Hh + Hh
        per = attn + mlp + (2 * Hh if ln else 0)
        assert net.params.numel() == per * 8
        flags = _lib.ATTN_CONCAT | _lib.ATTN_RESIDUAL | (_lib.ATTN_LAYER_NORM if ln else 0)
        d = _lib.FlowDesc(num_timesteps=2, node_embedding_dim=D, latent_dim=L, num_layers=K, agg=0, block=_lib.BLOCK["dm_attn"],
                          act=1, weight_sharing=0, eps=1.0, attn_num_heads=2, attn_kq_dim=5, attn_v_dim=4, attn_out_dim=7,
                          attn_flags=flags)
        assert lib.gnf_flow_param_count(C.byref(d)) == per * 8
        g0 = net.s[0][0]
        if ln:
            assert g0.ln_gamma.data_ptr() == net.params.data_ptr() + 4 * (attn + mlp)
            assert torch.equal(g0.ln_gamma, torch.ones(Hh)) and torch.equal(g0.ln_beta, torch.zeros(Hh))
        else:
            assert g0.ln_gamma is None


def test_batch_norm_parameters_are_trainable_only_with_the_bijector():
    on = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, use_batch_norm=True, device="cpu")
    off = G.GRevNet(mk(L=16, K=3, D=4), 2, 4, use_batch_norm=False, device="cpu")
    assert on.bn_gamma.requires_grad and on.bn_beta.requires_grad
    assert not off.bn_gamma.requires_grad and not off.bn_beta.requires_grad
    with pytest.raises(RuntimeError, match="CUDA"):
        on.backward_from_z(None, torch.zeros(3, 4), 1.0)           # no CPU fallback on the BN training path either


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgnf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(H.ROOT, "graph_normalizing_flows_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn


def test_graphs_tuple_replace_and_batching():
    g = H.random_batch(np.random.default_rng(0), 3, 3, 6, D=4)
    gt = G.GraphsTuple(*g)
    g2 = gt.replace(nodes=g.nodes * 2)
    assert g2.senders is gt.senders and not np.array_equal(g2.nodes, gt.nodes)
    dd = [{"n_node": 2, "senders": [0, 1], "receivers": [1, 0], "nodes": np.zeros((2, 4))},
          {"n_node": 3, "senders": [0, 2], "receivers": [2, 1], "nodes": np.ones((3, 4))}]
    b = G.graphs.data_dicts_to_graphs_tuple(dd)
    assert b.senders.tolist() == [0, 1, 2, 4] and b.receivers.tolist() == [1, 0, 4, 3]
    assert b.n_node.tolist() == [2, 3] and b.n_edge.tolist() == [2, 2] and b.nodes.dtype == np.float32
