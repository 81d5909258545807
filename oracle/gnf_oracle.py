"""CPU oracle for the GRevNet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does: it fails loudly when
the CUDA library is missing.

PARITY UNPINNED: the reference (jliu/graph-normalizing-flows @ d8b9256) ships no
tests and no golden vectors, and it cannot be imported here (needs TF 1.14,
graph_nets, dm-sonnet 1.34, tensorflow-probability 0.7.0 -- none installed, no
network).  This file restates the reference's algorithm op for op in numpy; what
pins it are the known-answer tests K1..K8 in tests/ (round trip, zero weights,
log-det vs. autograd Jacobian, graph independence, index formulas).

Third-party semantics restated from their published behaviour (not in the tree):
  graph_nets (unpinned, 1.0.x)   broadcast_sender_nodes_to_edges = gather(nodes, senders)
                                 ReceivedEdgesToNodesAggregator(r) = r(edges, receivers, sum(n_node))
  TF 1.14                        unsorted_segment_sum: CPU kernel accumulates serially in
                                 ascending edge index; unsorted_segment_mean = sum / max(count, 1);
                                 leaky_relu alpha = 0.2
  dm-sonnet 1.34                 Linear: x @ W + b;  nets.MLP: act after every layer but the last
  tensorflow-probability 0.7.0   MultivariateNormalDiag(0, I).log_prob(z) = -1/2 |z|^2 - D/2 ln(2 pi)

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import collections
import math

import numpy as np

# graph_nets.graphs.GraphsTuple field order [upstream]; used at gnn.py:307-308
# through .replace(nodes=...).
_FIELDS = ("nodes", "edges", "receivers", "senders", "globals", "n_node", "n_edge")


class GraphsTuple(collections.namedtuple("GraphsTuple", _FIELDS)):
    def replace(self, **kwargs):
        return self._replace(**kwargs)


LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- #
# a6  make_mlp_model  (gnn.py:159-180)
# --------------------------------------------------------------------------- #
def activation(x, kind):
    """tf.nn.leaky_relu (alpha=0.2) for the *_then_mlp factories (run_grevnet.py:158,166,179),
    tf.nn.relu is make_mlp_model's default (gnn.py:162)."""
    if kind == "leaky_relu":
        return np.maximum(x, x.dtype.type(0.2) * x)
    if kind == "relu":
        return np.maximum(x, x.dtype.type(0))
    raise ValueError(kind)


def mlp_forward(h, layers, act):
    """snt.nets.MLP(activate_final=False): gnn.py:167-180.  layers = [(W[in,out], b[out]), ...]."""
    last = len(layers) - 1
    for i, (w, b) in enumerate(layers):
        h = h @ w + b
        if i != last:
            h = activation(h, act)
    return h


def glorot_truncated_normal(rng, fan_in, fan_out, dtype=np.float32):
    """tf.initializers.glorot_normal (gnn.py:172): VarianceScaling(1, fan_avg, truncated_normal):
    stddev = sqrt(2/(fan_in+fan_out)) / 0.8796..., resampled outside 2 stddev."""
    std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    return truncated_normal(rng, (fan_in, fan_out), std, dtype)


def truncated_normal(rng, shape, std, dtype=np.float32):
    """tf.initializers.truncated_normal (gnn.py:173): N(0, std) re-drawn beyond 2 std."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(dtype)


def init_mlp(rng, in_dim, latent_dim, out_dim, num_layers, bias_init_stddev=0.1,
             last_layer_scale=1.0, dtype=np.float32):
    """Layer sizes [latent]*(K-1)+[out] (gnn.py:165-166)."""
    sizes = [latent_dim] * (num_layers - 1) + [out_dim]
    layers = []
    d = in_dim
    for i, o in enumerate(sizes):
        w = glorot_truncated_normal(rng, d, o, dtype)
        b = truncated_normal(rng, (o,), bias_init_stddev, dtype)
        if i == len(sizes) - 1:
            w = (w * last_layer_scale).astype(dtype)
            b = (b * last_layer_scale).astype(dtype)
        layers.append((w, b))
        d = o
    return layers


# --------------------------------------------------------------------------- #
# a3 / a4  sender gather and received-edges aggregator
# --------------------------------------------------------------------------- #
def gather_senders(nodes, senders):
    """EdgeBlock(IdentityModule, use_sender_nodes only), gnn.py:135-140,151-152:
    edges[e,:] = nodes[senders[e],:]."""
    return nodes[senders]


def segment_sum_serial(edges, receivers, num_segments):
    """tf.unsorted_segment_sum on CPU (gnn.py:239-257 pass it as the reducer): serial
    accumulation in ascending edge index.  np.add.at is unbuffered and in index order,
    so per (segment, feature) the floating-point add order is exactly that."""
    out = np.zeros((num_segments, edges.shape[1]), dtype=edges.dtype)
    np.add.at(out, receivers, edges)
    return out


def segment_mean_serial(edges, receivers, num_segments):
    """tf.unsorted_segment_mean [upstream]: segment_sum / max(count, 1)."""
    s = segment_sum_serial(edges, receivers, num_segments)
    cnt = np.bincount(receivers, minlength=num_segments).astype(edges.dtype)
    return s / np.maximum(cnt, edges.dtype.type(1))[:, None]


def aggregate(nodes, senders, receivers, agg):
    edges = gather_senders(nodes, senders)
    if agg == "sum":
        return segment_sum_serial(edges, receivers, nodes.shape[0])
    if agg == "mean":
        return segment_mean_serial(edges, receivers, nodes.shape[0])
    raise ValueError(agg)


# --------------------------------------------------------------------------- #
# a5  node blocks  +  NodeBlockGNN
# --------------------------------------------------------------------------- #
def dm_self_attention_mlp(nodes, senders, receivers, gnn, cfg):
    """DMSelfAttentionMLP._build (gnn.py:501-558) around DMSelfAttention._build (gnn.py:419-477).
    gnn = {"wq","wk","wv","wo","mlp"};  cfg carries num_heads, kq_dim, v_dim, concat, residual,
    kq_dim_division.  [graph_nets _unsorted_segment_softmax: subtract segment max, exp, / segment sum]"""
    dt = nodes.dtype
    n, heads, kq, vd = nodes.shape[0], cfg["num_heads"], cfg["kq_dim"], cfg["v_dim"]
    project_q = (nodes @ gnn["wq"]).reshape(n, heads, kq)           # gnn.py:509-520
    project_k = (nodes @ gnn["wk"]).reshape(n, heads, kq)
    project_v = np.repeat((nodes @ gnn["wv"])[:, None, :], heads, axis=1)    # keras.backend.repeat, gnn.py:528
    # attn_module(project_v, project_q, project_k, graph): keys = project_q, queries = project_k (gnn.py:531-532)
    sender_keys = project_q[senders]                                 # gnn.py:445-446
    sender_values = project_v[senders]
    receiver_queries = project_k[receivers]                          # gnn.py:453-454
    logits = np.sum(sender_keys * receiver_queries, axis=-1, dtype=dt)            # [E, heads]
    if cfg["kq_dim_division"]:
        logits = logits / np.sqrt(dt.type(kq))
    seg_max = np.full((n, heads), -np.inf, dt)
    np.maximum.at(seg_max, receivers, logits)
    ex = np.exp(logits - seg_max[receivers])
    seg_sum = np.zeros((n, heads), dt)
    np.add.at(seg_sum, receivers, ex)
    w = ex / seg_sum[receivers]
    attended = sender_values * w[..., None]                          # gnn.py:467
    agg = np.zeros((n, heads, vd), dt)
    np.add.at(agg, receivers, attended)                              # unsorted_segment_sum, gnn.py:471-475
    new_nodes = agg.reshape(n, heads * vd) @ gnn["wo"]               # gnn.py:541-545
    if cfg["concat"]:
        new_nodes = np.concatenate([nodes, new_nodes], axis=1)       # gnn.py:547-548
    new_nodes = mlp_forward(new_nodes, gnn["mlp"], cfg["act"])
    if cfg["residual"]:
        new_nodes = new_nodes + nodes                                # gnn.py:551-552
    if cfg.get("layer_norm"):                                        # snt.LayerNorm, gnn.py:554-556 [upstream: eps 1e-5,
        mean = new_nodes.mean(axis=1, keepdims=True)                 #  moments over the last axis, gamma * xhat + beta]
        var = ((new_nodes - mean) ** 2).mean(axis=1, keepdims=True)
        new_nodes = (new_nodes - mean) / np.sqrt(var + dt.type(1e-5)) * gnn["ln_gamma"] + gnn["ln_beta"]
    return new_nodes


def node_block_gnn(nodes, senders, receivers, layers, cfg):
    """NodeBlockGNN._build (gnn.py:155-156) with ConcatThenMLPBlock (gnn.py:107-111)
    or AggThenMLPBlock (gnn.py:122-126)."""
    if cfg["block"] == "dm_attn":
        return dm_self_attention_mlp(nodes, senders, receivers, layers, cfg)
    agg = aggregate(nodes, senders, receivers, cfg["agg"])
    if cfg["block"] == "concat":
        h = np.concatenate([nodes, agg], axis=1)
    elif cfg["block"] == "agg_then":
        h = nodes.dtype.type(cfg["eps"]) * nodes + agg
    else:
        raise ValueError(cfg["block"])
    return mlp_forward(h, layers, cfg["act"])


# --------------------------------------------------------------------------- #
# a7  GRevNet.f / GRevNet.g   (gnn.py:304-373)
# --------------------------------------------------------------------------- #
def _st(params, which, half, i):
    if params["weight_sharing"]:          # gnn.py:284-286,316-319
        return params[which][half]
    return params[which][half][i]         # gnn.py:320-321


def grevnet_f(nodes, senders, receivers, params):
    """GRevNet.f (gnn.py:304-341), use_batch_norm=False.  Returns (z, log_det_jacobian);
    ldj accumulates in python order exactly as gnn.py:322,337 do."""
    cfg = params["cfg"]
    dt = nodes.dtype
    h = nodes.shape[1] // 2
    x0, x1 = nodes[:, :h].copy(), nodes[:, h:].copy()       # tf.split, gnn.py:306
    ldj = dt.type(0)
    for i in range(params["T"]):
        s = node_block_gnn(x0, senders, receivers, _st(params, "s", 0, i), cfg)
        t = node_block_gnn(x0, senders, receivers, _st(params, "t", 0, i), cfg)
        ldj = ldj + np.sum(s, dtype=dt)                       # gnn.py:322
        x1 = x1 * np.exp(s) + t                               # gnn.py:323
        s = node_block_gnn(x1, senders, receivers, _st(params, "s", 1, i), cfg)
        t = node_block_gnn(x1, senders, receivers, _st(params, "t", 1, i), cfg)
        ldj = ldj + np.sum(s, dtype=dt)                       # gnn.py:337
        x0 = x0 * np.exp(s) + t                               # gnn.py:338
    return np.concatenate([x0, x1], axis=1), ldj              # gnn.py:340-341


def grevnet_g(nodes, senders, receivers, params):
    """GRevNet.g (gnn.py:343-373), use_batch_norm=False."""
    cfg = params["cfg"]
    h = nodes.shape[1] // 2
    z0, z1 = nodes[:, :h].copy(), nodes[:, h:].copy()
    for i in reversed(range(params["T"])):                   # gnn.py:347
        s = node_block_gnn(z1, senders, receivers, _st(params, "s", 1, i), cfg)
        t = node_block_gnn(z1, senders, receivers, _st(params, "t", 1, i), cfg)
        z0 = (z0 - t) * np.exp(-s)                            # gnn.py:359
        s = node_block_gnn(z0, senders, receivers, _st(params, "s", 0, i), cfg)
        t = node_block_gnn(z0, senders, receivers, _st(params, "t", 0, i), cfg)
        z1 = (z1 - t) * np.exp(-s)                            # gnn.py:372
    return np.concatenate([z0, z1], axis=1)


# --------------------------------------------------------------------------- #
# a9  TFP batch-norm bijector inside the coupling (gnn.py:260-263,310-313,325-328,356-358,369-371)
# --------------------------------------------------------------------------- #
BN_EPS = 1e-3         # tf.layers.BatchNormalization default epsilon
BN_MOMENTUM = 0.99    # default momentum


def make_bn_state(T, H, dtype=np.float32):
    """bns[2][T] of make_batch_norm (gnn.py:260-263,301-302): gamma=1, beta=0, moving_mean=0,
    moving_variance=1 (Keras defaults)."""
    mk = lambda: {"gamma": np.ones(H, dtype), "beta": np.zeros(H, dtype),
                  "moving_mean": np.zeros(H, dtype), "moving_var": np.ones(H, dtype)}
    return [[mk() for _ in range(T)], [mk() for _ in range(T)]]


def bn_inverse(x, bn, update=True):
    """tfb.BatchNormalization(training=True): returns (inverse(x), inverse_log_det_jacobian(x, 2)).
    [upstream, tensorflow-probability 0.7.0, restated; unverifiable here]
      inverse  = batchnorm_layer(x, training=True): gamma*(x-mu_B)/sqrt(var_B+eps)+beta with the
                 BIASED batch variance over the N nodes (tf.nn.moments); moving stats updated with
                 momentum 0.99 (the scripts run UPDATE_OPS, run_grevnet.py:362).
      ildj     = sum_f(log gamma_f - 1/2 log(var_B,f + eps)), a scalar, tiled over the extra event
                 dimension (event_ndims=2, forward_min_event_ndims=1) -> multiplied by N."""
    dt = x.dtype
    n = x.shape[0]
    mean = np.mean(x, axis=0, dtype=dt)
    var = np.mean((x - mean) ** 2, axis=0, dtype=dt)
    ildj = dt.type(n) * np.sum(np.log(bn["gamma"].astype(dt)) - dt.type(0.5) * np.log(var + dt.type(BN_EPS)), dtype=dt)
    y = bn["gamma"].astype(dt) * (x - mean) / np.sqrt(var + dt.type(BN_EPS)) + bn["beta"].astype(dt)
    if update:
        bn["moving_mean"] = (bn["moving_mean"] * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)).astype(bn["moving_mean"].dtype)
        bn["moving_var"] = (bn["moving_var"] * BN_MOMENTUM + var * (1 - BN_MOMENTUM)).astype(bn["moving_var"].dtype)
    return y, ildj


def bn_forward(z, bn):
    """bijector forward = de-normalise with the MOVING statistics:
    sqrt(moving_var+eps)/gamma * (z - beta) + moving_mean."""
    dt = z.dtype
    return (np.sqrt(bn["moving_var"].astype(dt) + dt.type(BN_EPS)) / bn["gamma"].astype(dt)
            * (z - bn["beta"].astype(dt)) + bn["moving_mean"].astype(dt))


def grevnet_f_bn(nodes, senders, receivers, params, bns, update=True):
    """GRevNet.f with use_batch_norm=True (gnn.py:304-341)."""
    cfg = params["cfg"]
    dt = nodes.dtype
    h = nodes.shape[1] // 2
    x0, x1 = nodes[:, :h].copy(), nodes[:, h:].copy()
    ldj = dt.type(0)
    for i in range(params["T"]):
        y, ild = bn_inverse(x0, bns[0][i], update)            # gnn.py:310-313 (ildj on the un-normalised x0)
        ldj = ldj + ild
        x0 = y
        s = node_block_gnn(x0, senders, receivers, _st(params, "s", 0, i), cfg)
        t = node_block_gnn(x0, senders, receivers, _st(params, "t", 0, i), cfg)
        ldj = ldj + np.sum(s, dtype=dt)
        x1 = x1 * np.exp(s) + t
        y, ild = bn_inverse(x1, bns[1][i], update)            # gnn.py:325-328
        ldj = ldj + ild
        x1 = y
        s = node_block_gnn(x1, senders, receivers, _st(params, "s", 1, i), cfg)
        t = node_block_gnn(x1, senders, receivers, _st(params, "t", 1, i), cfg)
        ldj = ldj + np.sum(s, dtype=dt)
        x0 = x0 * np.exp(s) + t
    return np.concatenate([x0, x1], axis=1), ldj


def grevnet_g_bn(nodes, senders, receivers, params, bns):
    """GRevNet.g with use_batch_norm=True (gnn.py:343-373): s,t from the normalised-domain half,
    THEN that half is de-normalised with the moving statistics."""
    cfg = params["cfg"]
    h = nodes.shape[1] // 2
    z0, z1 = nodes[:, :h].copy(), nodes[:, h:].copy()
    for i in reversed(range(params["T"])):
        s = node_block_gnn(z1, senders, receivers, _st(params, "s", 1, i), cfg)
        t = node_block_gnn(z1, senders, receivers, _st(params, "t", 1, i), cfg)
        z1 = bn_forward(z1, bns[1][i])                        # gnn.py:356-358
        z0 = (z0 - t) * np.exp(-s)
        s = node_block_gnn(z0, senders, receivers, _st(params, "s", 0, i), cfg)
        t = node_block_gnn(z0, senders, receivers, _st(params, "t", 0, i), cfg)
        z0 = bn_forward(z0, bns[0][i])                        # gnn.py:369-371
        z1 = (z1 - t) * np.exp(-s)
    return np.concatenate([z0, z1], axis=1)


def grevnet_call(graph, params, inverse=True):
    """GRevNet._build (gnn.py:379-381): inverse=True is the data->latent (density) direction."""
    if inverse:
        z, ldj = grevnet_f(graph.nodes, graph.senders, graph.receivers, params)
        return graph.replace(nodes=z), ldj
    return graph.replace(nodes=grevnet_g(graph.nodes, graph.senders, graph.receivers, params))


# --------------------------------------------------------------------------- #
# a8  log-prob assembly  (run_grevnet.py:290-302, train_grevnet_with_data.py:346-355)
# --------------------------------------------------------------------------- #
def log_prob(z_nodes, ldj, n_node):
    """Returns the scalars the scripts log: log_prob_zs, log_prob_xs, total_loss and
    the per-node normalisations by sum(n_node)."""
    dt = z_nodes.dtype
    d = z_nodes.shape[1]
    per_node = dt.type(-0.5) * np.sum(z_nodes * z_nodes, axis=1, dtype=dt) - dt.type(0.5 * d * LOG_2PI)
    log_prob_zs = np.sum(per_node, dtype=dt)                  # run_grevnet.py:294
    log_prob_xs = log_prob_zs + dt.type(ldj)                  # run_grevnet.py:295
    total_loss = -log_prob_xs                                 # run_grevnet.py:296
    num_nodes = dt.type(np.sum(n_node))                       # run_grevnet.py:298
    return {
        "log_prob_zs": log_prob_zs,
        "log_det_jacobian": dt.type(ldj),
        "log_prob_xs": log_prob_xs,
        "total_loss": total_loss,
        "num_nodes": num_nodes,
        "loss_per_node": total_loss / num_nodes,
        "log_prob_xs_per_node": log_prob_xs / num_nodes,
        "log_prob_zs_per_node": log_prob_zs / num_nodes,
        "log_det_jacobian_per_node": dt.type(ldj) / num_nodes,
    }


# --------------------------------------------------------------------------- #
# a2  GRevNet.__init__  (gnn.py:274-302): 4*T MLPs (or 4 when weight_sharing)
# --------------------------------------------------------------------------- #
def make_params(seed, T, D, latent_dim=256, num_layers=5, agg="sum", block="concat", eps=1.0,
                act="leaky_relu", bias_init_stddev=0.1, last_layer_scale=1.0,
                weight_sharing=False, dtype=np.float32, attn=None):
    """attn (block == "dm_attn"): dict(num_heads, kq_dim, v_dim, out_dim, concat, residual, kq_dim_division)."""
    if D % 2:
        raise ValueError("node_embedding_dim must be even (tf.split, gnn.py:306)")
    h = D // 2
    in_dim = D if block == "concat" else h
    rng = np.random.default_rng(seed)
    if block == "dm_attn":
        in_dim = h + attn["out_dim"] if attn["concat"] else attn["out_dim"]

    def xavier(i, o):        # tf.contrib.layers.xavier_initializer(uniform=True), gnn.py:504-506
        lim = math.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, (i, o)).astype(dtype)

    def mk():
        if block == "dm_attn":
            qk, hv = attn["num_heads"] * attn["kq_dim"], attn["num_heads"] * attn["v_dim"]
            g = {"wq": xavier(h, qk), "wk": xavier(h, qk), "wv": xavier(h, attn["v_dim"]),
                 "wo": truncated_normal(rng, (hv, attn["out_dim"]), 1.0 / math.sqrt(hv), dtype)}
            g["mlp"] = init_mlp(rng, in_dim, latent_dim, h, num_layers, bias_init_stddev, last_layer_scale, dtype)
            if attn.get("layer_norm"):           # init is gamma = 1, beta = 0; perturbed so that tests see them
                g["ln_gamma"] = (1.0 + 0.1 * rng.standard_normal(h)).astype(dtype)
                g["ln_beta"] = (0.1 * rng.standard_normal(h)).astype(dtype)
            return g
        return init_mlp(rng, in_dim, latent_dim, h, num_layers, bias_init_stddev,
                        last_layer_scale, dtype)

    if weight_sharing:
        s = [mk(), mk()]
        t = [mk(), mk()]
    else:                                   # construction order of gnn.py:292-299
        s = [[mk() for _ in range(T)], [mk() for _ in range(T)]]
        t = [[mk() for _ in range(T)], [mk() for _ in range(T)]]
    cfg = {"agg": agg, "block": block, "eps": float(eps), "act": act}
    if block == "dm_attn":
        cfg.update(attn)
    return {"T": T, "D": D, "s": s, "t": t, "weight_sharing": weight_sharing, "cfg": cfg}


def cast_params(params, dtype):
    def c(m):
        if isinstance(m, dict):
            out = {k: m[k].astype(dtype) for k in ("wq", "wk", "wv", "wo", "ln_gamma", "ln_beta") if k in m}
            out["mlp"] = [(w.astype(dtype), b.astype(dtype)) for (w, b) in m["mlp"]]
            return out
        return [(w.astype(dtype), b.astype(dtype)) for (w, b) in m]
    out = dict(params)
    for k in ("s", "t"):
        if params["weight_sharing"]:
            out[k] = [c(m) for m in params[k]]
        else:
            out[k] = [[c(m) for m in half] for half in params[k]]
    return out


# --------------------------------------------------------------------------- #
# f3  decode tail: pred_adj + scaled_hacky_sigmoid_l2   (loss.py:154-159,45-53,131-151,83-85)
# --------------------------------------------------------------------------- #
def scaled_hacky_sigmoid_l2(nodes):
    """loss.py:45-53."""
    dt = nodes.dtype
    dim = nodes.shape[1]
    r = np.sum(np.square(nodes), 1, dtype=dt).reshape(-1, 1)
    D = r - dt.type(2) * (nodes @ nodes.T) + r.T
    D = D / np.sqrt(dt.type(dim))
    return dt.type(1) / (dt.type(1) + np.exp(-dt.type(10) * (dt.type(1) - D)))


def loss_mask(n_node):
    """loss.py:131-151: block-diagonal ones."""
    n = int(np.sum(n_node))
    m = np.zeros((n, n), np.float32)
    lo = 0
    for k in n_node:
        m[lo:lo + int(k), lo:lo + int(k)] = 1
        lo += int(k)
    return m


def pred_adj(nodes, n_node):
    """loss.py:154-159: distance_fn(nodes) * loss_mask, diagonal removed (loss.py:83-85)."""
    p = scaled_hacky_sigmoid_l2(nodes) * loss_mask(n_node).astype(nodes.dtype)
    return p * (1 - np.eye(p.shape[0], dtype=nodes.dtype))
