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
    assert lib.gnf_abi_version() == 3


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
