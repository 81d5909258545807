"""torch-CPU restatement of the reference path, used ONLY as the timed CPU baseline
(bench.py cpu_baseline / --impl reference) and cross-checked against oracle/gnf_oracle.py in
tests/.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see gnf_oracle.py).

Same op order as the reference executes on a CPU TensorFlow build (SURVEY §8d "CPU baseline"):
tf.gather -> index_select, tf.unsorted_segment_sum -> index_add_ (serial over edges on CPU),
snt.Linear -> addmm, s and t GNNs evaluated separately (gnn.py:320-321), x*exp(s)+t, scalar sum.
All host threads torch can use (MKL / OpenMP), like TF's intra-op pool.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def params_to_torch(params, dtype=torch.float32):
    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)

    def c(m):
        if isinstance(m, dict):                  # dm_self_attn GNN: projections + MLP (oracle/gnf_oracle.py)
            out = {k: t(m[k]) for k in ("wq", "wk", "wv", "wo", "ln_gamma", "ln_beta") if k in m}
            out["mlp"] = [(t(w), t(b)) for (w, b) in m["mlp"]]
            return out
        return [(t(w), t(b)) for (w, b) in m]
    out = dict(params)
    for k in ("s", "t"):
        out[k] = [c(m) for m in params[k]] if params["weight_sharing"] else [[c(m) for m in half] for half in params[k]]
    return out


def _act(h, kind):
    return torch.maximum(h, 0.2 * h) if kind == "leaky_relu" else torch.relu(h)   # tf.nn.leaky_relu alpha=0.2


def _mlp(h, layers, act):                       # gnn.py:167-180
    last = len(layers) - 1
    for i, (w, b) in enumerate(layers):
        h = torch.addmm(b, h, w)
        if i != last:
            h = _act(h, act)
    return h


def _dm_attn(nodes, senders, receivers, gnn, cfg):
    """DMSelfAttentionMLP._build (gnn.py:501-558), same restatement as oracle/gnf_oracle.py::dm_self_attention_mlp,
    differentiable."""
    n, heads, kq, vd = nodes.shape[0], cfg["num_heads"], cfg["kq_dim"], cfg["v_dim"]
    project_q = (nodes @ gnn["wq"]).reshape(n, heads, kq)
    project_k = (nodes @ gnn["wk"]).reshape(n, heads, kq)
    project_v = (nodes @ gnn["wv"])[:, None, :].expand(n, heads, vd)
    logits = (project_q[senders] * project_k[receivers]).sum(-1)          # keys = project_q, queries = project_k
    if cfg["kq_dim_division"]:
        logits = logits / math.sqrt(kq)
    idx = receivers[:, None].expand(-1, heads)
    seg_max = torch.full((n, heads), -float("inf"), dtype=nodes.dtype).scatter_reduce(0, idx, logits.detach(), "amax")
    ex = torch.exp(logits - seg_max[receivers])
    seg_sum = torch.zeros((n, heads), dtype=nodes.dtype).index_add_(0, receivers, ex)
    w = ex / seg_sum[receivers]
    agg = torch.zeros((n, heads, vd), dtype=nodes.dtype).index_add_(0, receivers, project_v[senders] * w[..., None])
    new_nodes = agg.reshape(n, heads * vd) @ gnn["wo"]
    if cfg["concat"]:
        new_nodes = torch.cat([nodes, new_nodes], 1)
    new_nodes = _mlp(new_nodes, gnn["mlp"], cfg["act"])
    if cfg["residual"]:
        new_nodes = new_nodes + nodes
    if cfg.get("layer_norm"):
        mean = new_nodes.mean(1, keepdim=True)
        var = ((new_nodes - mean) ** 2).mean(1, keepdim=True)
        new_nodes = (new_nodes - mean) / torch.sqrt(var + 1e-5) * gnn["ln_gamma"] + gnn["ln_beta"]
    return new_nodes


def _gnn(nodes, senders, receivers, layers, cfg):   # gnn.py:155-156 + 107-111 / 122-126
    if cfg["block"] == "dm_attn":
        return _dm_attn(nodes, senders, receivers, layers, cfg)
    edges = nodes.index_select(0, senders)                       # gnn.py:151 (use_sender_nodes)
    agg = torch.zeros_like(nodes).index_add_(0, receivers, edges)  # gnn.py:103 unsorted_segment_sum
    if cfg["agg"] == "mean":
        cnt = torch.bincount(receivers, minlength=nodes.shape[0]).to(nodes.dtype).clamp_(min=1)
        agg = agg / cnt[:, None]
    h = torch.cat([nodes, agg], 1) if cfg["block"] == "concat" else cfg["eps"] * nodes + agg
    return _mlp(h, layers, cfg["act"])


def _st(p, which, half, i):
    return p[which][half] if p["weight_sharing"] else p[which][half][i]


BN_EPS = 1e-3      # tf.layers.BatchNormalization default epsilon


def _bn_inverse(x, gamma, beta):
    """tfb.BatchNormalization(training=True).inverse + inverse_log_det_jacobian(x, 2) (gnn.py:310-313), as
    oracle/gnf_oracle.py::bn_inverse restates it: batch moments (biased variance), scalar ildj tiled over the
    node axis.  Differentiable through the batch statistics."""
    mean = x.mean(0)
    var = ((x - mean) ** 2).mean(0)
    y = gamma * (x - mean) / torch.sqrt(var + BN_EPS) + beta
    ildj = x.shape[0] * (torch.log(gamma) - 0.5 * torch.log(var + BN_EPS)).sum()
    return y, ildj


def grevnet_f_autograd(nodes, senders, receivers, p, bn=None, ldj64_out=None):
    """GRevNet.f (gnn.py:304-341), differentiable (tests: reference gradients by autograd).
    bn = (gamma[2][T][H], beta[2][T][H]) switches use_batch_norm on.
    ldj64_out: optional list; receives the same log-det with every reduce_sum(s) accumulated in float64 (separates the
    reference's fp32 summation error from the error of s itself when a parity report compares log-dets)."""
    cfg = p["cfg"]
    h = nodes.shape[1] // 2
    x0, x1 = nodes[:, :h].contiguous(), nodes[:, h:].contiguous()
    ldj = torch.zeros((), dtype=nodes.dtype)
    for i in range(p["T"]):
        if bn is not None:
            x0, l = _bn_inverse(x0, bn[0][0][i], bn[1][0][i])
            ldj = ldj + l
        s = _gnn(x0, senders, receivers, _st(p, "s", 0, i), cfg)
        t = _gnn(x0, senders, receivers, _st(p, "t", 0, i), cfg)
        ldj = ldj + s.sum()
        if ldj64_out is not None:
            ldj64_out.append(float(s.detach().double().sum()))
        x1 = x1 * torch.exp(s) + t
        if bn is not None:
            x1, l = _bn_inverse(x1, bn[0][1][i], bn[1][1][i])
            ldj = ldj + l
        s = _gnn(x1, senders, receivers, _st(p, "s", 1, i), cfg)
        t = _gnn(x1, senders, receivers, _st(p, "t", 1, i), cfg)
        ldj = ldj + s.sum()
        if ldj64_out is not None:
            ldj64_out.append(float(s.detach().double().sum()))
        x0 = x0 * torch.exp(s) + t
    return torch.cat([x0, x1], 1), ldj


@torch.no_grad()
def grevnet_f(nodes, senders, receivers, p, ldj64_out=None):
    return grevnet_f_autograd(nodes, senders, receivers, p, ldj64_out=ldj64_out)


def log_prob_xs_autograd(z, ldj):              # run_grevnet.py:292-295
    d = z.shape[1]
    return (-0.5 * (z * z).sum(1) - 0.5 * d * math.log(2 * math.pi)).sum() + ldj


@torch.no_grad()
def log_prob_xs(z, ldj):
    return log_prob_xs_autograd(z, ldj)


def loss_and_grads(nodes, senders, receivers, params, scale=1.0, dtype=torch.float64, bn=None):
    """loss = -scale * log_prob_xs and its gradient w.r.t. every (W, b), by autograd, flattened in the
    include/gnf_b200.h parameter order (which -> half -> step; W0 b0 W1 b1 ...).
    bn = (gamma, beta) numpy [2, T, H]: use_batch_norm=True; returns (loss, grads, g_gamma, g_beta)."""
    p = params_to_torch(params, dtype)
    bn_t = None
    if bn is not None:
        bn_t = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dtype).requires_grad_(True) for a in bn)
    leaves = []
    for which in ("s", "t"):
        for half in range(2):
            mlps = [p[which][half]] if p["weight_sharing"] else p[which][half]
            for mlp in mlps:
                mlp_all = mlp
                if isinstance(mlp, dict):
                    for k in ("wq", "wk", "wv", "wo"):
                        mlp[k].requires_grad_(True)
                        leaves.append(mlp[k])
                    mlp = mlp["mlp"]
                ln = [mlp_all[k] for k in ("ln_gamma", "ln_beta") if isinstance(mlp_all, dict) and k in mlp_all]
                for i, (w, b) in enumerate(mlp):
                    w.requires_grad_(True)
                    b.requires_grad_(True)
                    leaves += [w, b]
                for v in ln:                     # LayerNorm variables come after the MLP's in the flat layout
                    v.requires_grad_(True)
                    leaves.append(v)
    z, ldj = grevnet_f_autograd(torch.as_tensor(nodes).to(dtype), torch.as_tensor(senders).long(),
                                torch.as_tensor(receivers).long(), p, bn=bn_t)
    loss = -scale * log_prob_xs_autograd(z, ldj)
    if bn_t is not None:
        grads = torch.autograd.grad(loss, leaves + list(bn_t))
        flat = torch.cat([g.reshape(-1) for g in grads[:-2]]).numpy()
        return float(loss.detach()), flat, grads[-2].numpy(), grads[-1].numpy()
    grads = torch.autograd.grad(loss, leaves)
    return float(loss.detach()), torch.cat([g.reshape(-1) for g in grads]).numpy()
