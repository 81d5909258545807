"""Independent pins for the oracle's restated [upstream] semantics ("next" item 8).

The reference delegates its arithmetic to graph_nets / Sonnet / TF / TFP, none of which can be imported here, so
oracle/gnf_oracle.py restates them from their published behaviour.  These tests check every restated primitive
against THIRD-PARTY code that is in the image (scipy, torch) or against a formulation that shares no code with
the oracle (dense masked softmax, autograd Jacobians, python loops).  What remains restated from memory only is
listed in DESIGN.md section 2.
"""
import math

import numpy as np
import pytest
import scipy.stats
import torch
import torch.nn.functional as F

import helpers as H
from oracle import gnf_oracle as O
from oracle import gnf_oracle_torch as OT


def test_pin_a8_mvn_log_prob_against_scipy_and_torch_distributions():
    """tfd.MultivariateNormalDiag(0, I).log_prob summed over nodes (run_grevnet.py:292-294)."""
    rng = np.random.default_rng(0)
    z = rng.standard_normal((57, 14)) * 1.7
    ours = O.log_prob(z, 0.0, np.array([57]))["log_prob_zs"]
    ref = scipy.stats.multivariate_normal(mean=np.zeros(14), cov=np.eye(14)).logpdf(z).sum()
    assert abs(ours - ref) <= 1e-12 * abs(ref)
    td = torch.distributions.Independent(torch.distributions.Normal(torch.zeros(14, dtype=torch.float64), 1.0), 1)
    assert abs(ours - float(td.log_prob(torch.from_numpy(z)).sum())) <= 1e-12 * abs(ref)
    # the torch-CPU baseline's assembly too
    got = float(OT.log_prob_xs(torch.from_numpy(z.astype(np.float32)), torch.tensor(3.5)))
    assert abs(got - (ref + 3.5)) <= 2e-6 * abs(ref)


def test_pin_a6_activations_bit_exact_against_torch():
    """tf.nn.leaky_relu(alpha=0.2) / tf.nn.relu: max(x, 0.2 x) must equal torch's select form bit for bit."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 50, np.array([0.0, -0.0, 1e-38, -1e-38, 3e38, -3e38], np.float32)])
    t = torch.from_numpy(x)
    assert np.array_equal(O.activation(x, "leaky_relu"), F.leaky_relu(t, 0.2).numpy())
    assert np.array_equal(O.activation(x, "relu"), F.relu(t).numpy())


def test_pin_a6_mlp_against_torch_nn_linear():
    """snt.nets.MLP: y = x @ W + b per layer, activation between layers, none after the last (gnn.py:167-180)."""
    rng = np.random.default_rng(2)
    layers = O.init_mlp(rng, 14, 32, 7, 4, dtype=np.float64)
    x = rng.standard_normal((19, 14))
    mods = []
    for i, (w, b) in enumerate(layers):
        lin = torch.nn.Linear(w.shape[0], w.shape[1]).double()
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(w.T))
            lin.bias.copy_(torch.from_numpy(b))
        mods.append(lin)
        if i != len(layers) - 1:
            mods.append(torch.nn.LeakyReLU(0.2))
    ref = torch.nn.Sequential(*mods)(torch.from_numpy(x)).detach().numpy()
    assert np.allclose(O.mlp_forward(x, layers, "leaky_relu"), ref, rtol=1e-13, atol=1e-13)


def test_pin_a6_glorot_truncated_normal_constants():
    """tf.initializers.glorot_normal = VarianceScaling(1, fan_avg, truncated_normal): the 0.8796... constant is the
    std of a unit normal truncated at 2 sigma (scipy), so the samples have variance 2 / (fan_in + fan_out)."""
    assert abs(scipy.stats.truncnorm(-2, 2).std() - 0.87962566103423978) < 1e-12
    w = O.glorot_truncated_normal(np.random.default_rng(3), 300, 500, np.float64)
    target = math.sqrt(2.0 / 800)
    assert abs(w.std() - target) < 0.01 * target
    assert np.abs(w).max() <= 2.0 * target / 0.87962566103423978 + 1e-12
    b = O.truncated_normal(np.random.default_rng(4), (20000,), 0.1, np.float64)
    assert np.abs(b).max() <= 0.2 + 1e-12 and abs(b.std() - 0.1 * 0.87962566103423978) < 2e-3


def test_pin_a4_segment_sum_order_against_index_add_and_python_loop():
    """tf.unsorted_segment_sum on CPU adds rows in ascending edge index: np.add.at (oracle), torch index_add_ on
    CPU (the timed baseline) and a plain python loop must agree BIT FOR BIT in fp32."""
    rng = np.random.default_rng(5)
    e, n, h = 5000, 37, 7
    edges = (rng.standard_normal((e, h)) * 10 ** rng.uniform(-3, 3, size=(e, 1))).astype(np.float32)
    recv = rng.integers(0, n, size=e).astype(np.int32)
    ours = O.segment_sum_serial(edges, recv, n)
    ref = torch.zeros(n, h).index_add_(0, torch.from_numpy(recv).long(), torch.from_numpy(edges)).numpy()
    assert np.array_equal(ours, ref)
    loop = np.zeros((n, h), np.float32)
    for k in range(e):
        loop[recv[k]] = loop[recv[k]] + edges[k]
    assert np.array_equal(ours, loop)
    # mean: sum / max(count, 1), empty segments give 0
    recv2 = recv.copy()
    recv2[recv2 == 3] = 4
    m = O.segment_mean_serial(edges, recv2, n)
    cnt = np.bincount(recv2, minlength=n)
    assert cnt[3] == 0 and np.all(m[3] == 0)
    assert np.array_equal(m[5], O.segment_sum_serial(edges, recv2, n)[5] / np.float32(cnt[5]))


def test_pin_a9_batch_norm_bijector_against_torch_batch_norm_and_autograd():
    """tfb.BatchNormalization(training=True).inverse == tf.layers.BatchNormalization(training=True): checked
    against torch.nn.functional.batch_norm(training=True); the inverse-log-det against the autograd Jacobian of the
    normalisation with the batch statistics held fixed (what the TFP bijector differentiates), tiled over nodes."""
    rng = np.random.default_rng(6)
    n, h = 23, 5
    x = rng.standard_normal((n, h)) * 3 + 1
    bn = O.make_bn_state(1, h, np.float64)[0][0]
    bn["gamma"] = rng.uniform(0.5, 2.0, h)
    bn["beta"] = rng.standard_normal(h)
    y, ildj = O.bn_inverse(x.copy(), bn, update=True)
    rm, rv = torch.zeros(h, dtype=torch.float64), torch.ones(h, dtype=torch.float64)
    ref = F.batch_norm(torch.from_numpy(x), rm, rv, torch.from_numpy(bn["gamma"]), torch.from_numpy(bn["beta"]),
                       training=True, momentum=1 - O.BN_MOMENTUM, eps=O.BN_EPS)
    assert np.allclose(y, ref.numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(bn["moving_mean"], rm.numpy(), rtol=1e-12, atol=1e-15)
    # torch updates running_var with the UNBIASED batch variance, Keras with the biased one (tf.nn.moments)
    biased = x.var(axis=0)
    assert np.allclose(bn["moving_var"], O.BN_MOMENTUM * 1.0 + (1 - O.BN_MOMENTUM) * biased, rtol=1e-12)
    assert np.allclose(rv.numpy(), O.BN_MOMENTUM * 1.0 + (1 - O.BN_MOMENTUM) * biased * n / (n - 1), rtol=1e-12)
    # log-det: Jacobian of vec(x) -> vec(y) with mean/var fixed is diagonal; slogdet over all N*H coordinates
    mean, var = torch.from_numpy(x.mean(axis=0)), torch.from_numpy(biased)
    g, b = torch.from_numpy(bn["gamma"]), torch.from_numpy(bn["beta"])
    fn = lambda v: (g * (v.view(n, h) - mean) / torch.sqrt(var + O.BN_EPS) + b).reshape(-1)
    jac = torch.autograd.functional.jacobian(fn, torch.from_numpy(x).reshape(-1))
    sign, logdet = torch.linalg.slogdet(jac)
    assert float(sign) == 1.0 and abs(float(logdet) - float(ildj)) <= 1e-10 * abs(float(ildj))
    # bijector forward (moving statistics) inverts an inverse taken with the same statistics
    bn2 = dict(bn, moving_mean=x.mean(axis=0), moving_var=biased)
    assert np.allclose(O.bn_forward(y, bn2), x, rtol=1e-10, atol=1e-10)


def _attn_case(layer_norm):
    rng = np.random.default_rng(7)
    g = H.random_batch(rng, 3, 4, 7, p_edge=0.5, D=8)
    params = O.make_params(11, 1, 8, 16, 3, block="dm_attn", act="relu",
                           attn=dict(num_heads=2, kq_dim=3, v_dim=5, out_dim=6, concat=True, residual=False,
                                     kq_dim_division=True, layer_norm=layer_norm))
    return g, params


def test_pin_f1_layer_norm_against_torch_layer_norm():
    """snt.LayerNorm (gnn.py:554-556): moments over the feature axis, eps 1e-5, gamma * xhat + beta."""
    g, params = _attn_case(True)
    p64 = O.cast_params(params, np.float64)
    gnn = dict(p64["s"][0][0])
    gnn["ln_gamma"] = np.linspace(0.5, 1.5, 4)
    gnn["ln_beta"] = np.linspace(-0.3, 0.3, 4)
    x = g.nodes[:, :4].astype(np.float64)
    with_ln = O.dm_self_attention_mlp(x, g.senders, g.receivers, gnn, p64["cfg"])
    no_ln = O.dm_self_attention_mlp(x, g.senders, g.receivers, gnn, dict(p64["cfg"], layer_norm=False))
    ref = F.layer_norm(torch.from_numpy(no_ln), (4,), torch.from_numpy(gnn["ln_gamma"]), torch.from_numpy(gnn["ln_beta"]),
                       eps=1e-5).numpy()
    assert np.allclose(with_ln, ref, rtol=1e-12, atol=1e-12)


def test_pin_f1_segment_softmax_attention_against_dense_masked_softmax():
    """gn.modules._unsorted_segment_softmax + weighted segment sum (gnn.py:445-475) == dense attention with a
    -inf mask where there is no edge (torch.softmax), an independent formulation; duplicate edges excluded."""
    g, params = _attn_case(False)
    p64 = O.cast_params(params, np.float64)
    gnn, cfg = p64["s"][0][0], p64["cfg"]
    x = g.nodes[:, :4].astype(np.float64)
    n, heads, kq, vd = x.shape[0], cfg["num_heads"], cfg["kq_dim"], cfg["v_dim"]
    assert len(set(zip(g.senders.tolist(), g.receivers.tolist()))) == len(g.senders)
    xt = torch.from_numpy(x)
    keys = (xt @ torch.from_numpy(gnn["wq"])).view(n, heads, kq)       # at the sender  (gnn.py:531-532)
    queries = (xt @ torch.from_numpy(gnn["wk"])).view(n, heads, kq)    # at the receiver
    values = xt @ torch.from_numpy(gnn["wv"])                           # shared by the heads (gnn.py:528)
    logits = torch.einsum("rhd,shd->hrs", queries, keys) / math.sqrt(kq)
    mask = torch.full((n, n), float("-inf"), dtype=torch.float64)
    mask[torch.from_numpy(g.receivers).long(), torch.from_numpy(g.senders).long()] = 0.0
    w = torch.softmax(logits + mask, dim=-1)                             # [heads, receiver, sender]
    att = torch.einsum("hrs,sc->rhc", w, values).reshape(n, heads * vd)
    h = torch.cat([xt, att @ torch.from_numpy(gnn["wo"])], dim=1).numpy()
    ref = O.mlp_forward(h, gnn["mlp"], cfg["act"])
    ours = O.dm_self_attention_mlp(x, g.senders, g.receivers, gnn, cfg)
    assert np.allclose(ours, ref, rtol=1e-11, atol=1e-12)


def test_pin_a7_log_det_is_the_autograd_jacobian_log_det():
    """K3 with an exact Jacobian: slogdet(d vec(z) / d vec(x)) from torch.autograd of the torch restatement equals
    the sum of s the flow returns (gnn.py:322,337), for a flow whose coupling couples nodes AND features."""
    rng = np.random.default_rng(8)
    g = H.random_batch(rng, 2, 3, 4, p_edge=0.6, D=4)
    params = O.make_params(5, 2, 4, 8, 3, last_layer_scale=0.5)
    p = OT.params_to_torch(params, torch.float64)
    s, r = torch.from_numpy(g.senders).long(), torch.from_numpy(g.receivers).long()
    x = torch.from_numpy(g.nodes.astype(np.float64))
    n, d = x.shape
    fn = lambda v: OT.grevnet_f_autograd(v.view(n, d), s, r, p)[0].reshape(-1)
    jac = torch.autograd.functional.jacobian(fn, x.reshape(-1))
    sign, logdet = torch.linalg.slogdet(jac)
    _, ldj = OT.grevnet_f_autograd(x, s, r, p)
    assert float(sign) == 1.0 and abs(float(logdet) - float(ldj)) <= 1e-10 * max(1.0, abs(float(ldj)))
    # and the numpy oracle returns the same log-det
    _, ldj_np = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, O.cast_params(params, np.float64))
    assert abs(float(ldj_np) - float(ldj)) <= 1e-10 * max(1.0, abs(float(ldj)))
