"""Host-side mirror of the reference's GRevNet / message-passing-GNN interface (gnn.py).

Same names, constructor arguments, keyword semantics and return conventions as
/root/reference/gnn.py for the hot path:

    make_mlp_model (gnn.py:159-180)            ConcatThenMLPBlock (gnn.py:100-111)
    AggThenMLPBlock (gnn.py:114-126)           NodeBlockGNN (gnn.py:143-156)
    sum_concat_then_mlp_gnn / avg_concat_then_mlp_gnn / sum_then_mlp_gnn / avg_then_mlp_gnn
    (gnn.py:238-257)                           GRevNet (gnn.py:273-381)

The objects are thin parameter containers; all arithmetic runs in libgnf_b200.so
(hand-written sm_100a kernels, include/gnf_b200.h).  There is no CPU or eager-PyTorch
fallback: calling any of these without the library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional

import torch
from torch import nn

from . import _lib
from .graphs import BatchStructure, GraphsTuple, structure_of, transposed_structure_of

# activation tokens (the reference passes tf.nn.leaky_relu / tf.nn.relu, run_grevnet.py:158, gnn.py:162)
leaky_relu = "leaky_relu"
relu = "relu"


def _act_name(activation) -> str:
    if activation in ("leaky_relu", "relu"):
        return activation
    name = getattr(activation, "__name__", None)
    if name in ("leaky_relu", "relu"):
        return name
    raise ValueError(f"unsupported activation {activation!r}: the fused kernels implement "
                     "leaky_relu(alpha=0.2) and relu")


# reducer tokens (the reference passes tf.unsorted_segment_sum / tf.unsorted_segment_mean, gnn.py:239-257)
def unsorted_segment_sum(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int) -> torch.Tensor:
    """tf.unsorted_segment_sum on the device: serial in-order accumulation per segment
    (bit-exact with the TF CPU kernel)."""
    return _segment_reduce(data, segment_ids, num_segments, "sum")


def unsorted_segment_mean(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int) -> torch.Tensor:
    """tf.unsorted_segment_mean: segment_sum / max(count, 1)."""
    return _segment_reduce(data, segment_ids, num_segments, "mean")


unsorted_segment_sum.agg = "sum"
unsorted_segment_mean.agg = "mean"


def _agg_name(aggn_fn) -> str:
    agg = getattr(aggn_fn, "agg", None) or (aggn_fn if aggn_fn in ("sum", "mean") else None)
    if agg is None:
        raise ValueError("aggn_fn must be unsorted_segment_sum or unsorted_segment_mean")
    return agg


def _segment_reduce(data, segment_ids, num_segments, agg):
    lib = _lib.load()
    _lib.require_cuda(data, "data", torch.float32)
    _lib.require_cuda(segment_ids, "segment_ids", torch.int32)
    if data.dim() != 2 or data.shape[0] != segment_ids.shape[0]:
        raise ValueError("data must be [E, H] with one segment id per row")
    st = BatchStructure(segment_ids, segment_ids, int(num_segments))
    out = torch.empty(int(num_segments), data.shape[1], dtype=torch.float32, device=data.device)
    _lib.check(lib.gnf_segment_sum(_lib.ptr(data), data.shape[1], _lib.ptr(st.rowptr), _lib.ptr(st.perm),
                                   int(num_segments), _lib.AGG[agg], _lib.ptr(out), _lib.stream_ptr(data.device)),
               "gnf_segment_sum")
    return out


def gather_segment_reduce(nodes: torch.Tensor, structure: BatchStructure, agg: str = "sum") -> torch.Tensor:
    """EdgeBlock(use_sender_nodes) + ReceivedEdgesToNodesAggregator fused (gnn.py:151-156,103-104)."""
    lib = _lib.load()
    _lib.require_cuda(nodes, "nodes", torch.float32)
    out = torch.empty_like(nodes)
    _lib.check(lib.gnf_gather_segment_sum(_lib.ptr(nodes), nodes.shape[1], _lib.ptr(structure.rowptr),
                                          _lib.ptr(structure.csr_senders), nodes.shape[0], _lib.AGG[agg],
                                          _lib.ptr(out), _lib.stream_ptr(nodes.device)),
               "gnf_gather_segment_sum")
    return out


# --------------------------------------------------------------------------------------------
# make_mlp_model  (gnn.py:159-180)
# --------------------------------------------------------------------------------------------
def _truncated_normal_(t: torch.Tensor, std: float, generator=None):
    """tf truncated_normal: N(0, std) re-drawn outside 2 std."""
    with torch.no_grad():
        t.normal_(0.0, 1.0, generator=generator)
        bad = t.abs() > 2.0
        while bool(bad.any()):
            t[bad] = torch.randn(int(bad.sum()), generator=generator, dtype=t.dtype, device=t.device)
            bad = t.abs() > 2.0
        t.mul_(std)
    return t


class MLP(nn.Module):
    """snt.nets.MLP([latent]*(K-1)+[out], activate_final=False) with Glorot-normal weights and
    truncated-normal biases (gnn.py:165-180).  The input width is bound on first use
    (Sonnet infers it at connection time)."""

    def __init__(self, latent_dim, output_dim, num_layers, activation=relu, l2_regularizer_weight=0.01,
                 bias_init_stddev=0.1):
        super().__init__()
        self.latent_dim = int(latent_dim)
        self.output_dim = int(output_dim)          # the scripts pass node_embedding_dim / 2 (a float)
        self.num_layers = int(num_layers)
        if self.num_layers < 2:
            raise ValueError("num_layers must be >= 2")
        self.activation = _act_name(activation)
        self.l2_regularizer_weight = l2_regularizer_weight   # unused, as in the reference (commented out)
        self.bias_init_stddev = float(bias_init_stddev)
        self.input_dim: Optional[int] = None
        self.weights: List[torch.Tensor] = []
        self.biases: List[torch.Tensor] = []

    def layer_shapes(self, input_dim):
        sizes = [self.latent_dim] * (self.num_layers - 1) + [self.output_dim]
        shapes, d = [], int(input_dim)
        for o in sizes:
            shapes.append((d, o))
            d = o
        return shapes

    def param_count(self, input_dim):
        return sum(i * o + o for i, o in self.layer_shapes(input_dim))

    def bind(self, input_dim: int, flat: torch.Tensor, generator=None, init=True):
        """Attach this MLP to a slice of a flat parameter buffer (layout of
        include/gnf_b200.h: W0 b0 W1 b1 ...) and initialise it."""
        self.input_dim = int(input_dim)
        self.weights, self.biases = [], []
        off = 0
        for (i, o) in self.layer_shapes(input_dim):
            w = flat[off:off + i * o].view(i, o)
            off += i * o
            b = flat[off:off + o]
            off += o
            if init:
                # tf.initializers.glorot_normal = VarianceScaling(1, fan_avg, truncated_normal)
                std = math.sqrt(2.0 / (i + o)) / 0.87962566103423978
                _truncated_normal_(w, std, generator)
                _truncated_normal_(b, self.bias_init_stddev, generator)
            self.weights.append(w)
            self.biases.append(b)
        return off

    def signature(self):
        return (self.latent_dim, self.output_dim, self.num_layers, self.activation)


def make_mlp_model(latent_dim, output_dim, num_layers, activation=relu, l2_regularizer_weight=0.01,
                   bias_init_stddev=0.1) -> MLP:
    return MLP(latent_dim, output_dim, num_layers, activation, l2_regularizer_weight, bias_init_stddev)


# --------------------------------------------------------------------------------------------
# node blocks and NodeBlockGNN  (gnn.py:100-156)
# --------------------------------------------------------------------------------------------
class ConcatThenMLPBlock(nn.Module):
    """nodes <- MLP(concat([nodes, aggregate(received edges)], 1))   (gnn.py:100-111)"""
    block = "concat"

    def __init__(self, aggn_fn, make_mlp_fn, name="AggThenMLPBlock"):
        super().__init__()
        self.name = name
        self.agg = _agg_name(aggn_fn)
        self._mlp = make_mlp_fn()
        self.epsilon = 1.0

    def input_dim(self, half_dim):
        return 2 * half_dim


class AggThenMLPBlock(nn.Module):
    """nodes <- MLP(epsilon * nodes + aggregate(received edges))   (gnn.py:114-126)"""
    block = "agg_then"

    def __init__(self, aggn_fn, make_mlp_fn, epsilon, name="AggThenMLPBlock"):
        super().__init__()
        self.name = name
        self.agg = _agg_name(aggn_fn)
        self._mlp = make_mlp_fn()
        self.epsilon = float(epsilon)

    def input_dim(self, half_dim):
        return half_dim


EDGE_BLOCK_OPT = {          # gnn.py:135-140: edges = nodes[senders], nothing else
    "use_edges": False,
    "use_receiver_nodes": False,
    "use_sender_nodes": True,
    "use_globals": False,
}


class _Flow:
    """Owner of one gnf_flow handle (packed device weights)."""

    def __init__(self, T, D, mlp: MLP, block: str, agg: str, eps: float, weight_sharing: bool, attn=None):
        self.lib = _lib.load()
        attn = attn or {}
        self.desc = _lib.FlowDesc(
            num_timesteps=T, node_embedding_dim=D, latent_dim=mlp.latent_dim, num_layers=mlp.num_layers,
            agg=_lib.AGG[agg], block=_lib.BLOCK[block], act=_lib.ACT[mlp.activation],
            weight_sharing=int(bool(weight_sharing)), eps=eps,
            attn_num_heads=attn.get("num_heads", 0), attn_kq_dim=attn.get("kq_dim", 0),
            attn_v_dim=attn.get("v_dim", 0), attn_out_dim=attn.get("out_dim", 0), attn_flags=attn.get("flags", 0))
        n = self.lib.gnf_flow_param_count(C.byref(self.desc))
        if n < 0:
            raise ValueError(_lib.last_error())
        self.param_count = int(n)
        self.handle = None
        self._packed_version = None
        self._packed_ptr = None

    def ensure(self, flat: torch.Tensor):
        """Create the handle on first use and (re)pack when the flat buffer changed."""
        _lib.require_cuda(flat, "GRevNet parameters", torch.float32)
        if self.handle is None:
            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device: the GRevNet hot path has no CPU fallback")
            h = C.c_void_p()
            with torch.cuda.device(flat.device):
                _lib.check(self.lib.gnf_flow_create(C.byref(h), C.byref(self.desc)), "gnf_flow_create")
            self.handle = h
        if self._packed_version != flat._version or self._packed_ptr != flat.data_ptr():
            _lib.check(self.lib.gnf_flow_set_params(self.handle, _lib.ptr(flat), _lib.stream_ptr(flat.device)),
                       "gnf_flow_set_params")
            self._packed_version, self._packed_ptr = flat._version, flat.data_ptr()
        return self.handle

    def supports(self, math_name: str) -> bool:
        return bool(self.handle is not None and self.lib.gnf_flow_supports(self.handle, _lib.MATH[math_name]))

    def supports_backward(self, math_name: str) -> bool:
        return bool(self.handle is not None and
                    self.lib.gnf_flow_supports_backward(self.handle, _lib.MATH[math_name]))

    def __del__(self):
        try:
            if self.handle is not None:
                self.lib.gnf_flow_destroy(self.handle)
        except Exception:
            pass


class NodeBlockGNN(nn.Module):
    """node_block(edge_block(graph)) with an identity edge model over sender nodes
    (gnn.py:143-156).  Callable GraphsTuple -> GraphsTuple; inside a GRevNet its weights are
    adopted into the flow's packed buffer and the fused kernels do the work."""

    def __init__(self, node_block, edge_block_opt=EDGE_BLOCK_OPT, name="NodeBlockGNN"):
        super().__init__()
        self.name = name
        self._node_block = node_block
        self._flow: Optional[_Flow] = None
        self.params: Optional[nn.Parameter] = None

    @property
    def mlp(self) -> MLP:
        return self._node_block._mlp

    def config(self):
        nb = self._node_block
        return (nb.block, nb.agg, nb.epsilon) + self.mlp.signature()

    # -- what a flow needs from one of its GNNs ----------------------------------------------------
    def flow_kwargs(self):
        nb = self._node_block
        return dict(block=nb.block, agg=nb.agg, eps=nb.epsilon, attn=None)

    def gnn_param_count(self, half_dim):
        return self.mlp.param_count(self._node_block.input_dim(half_dim))

    def bind_params(self, half_dim, flat, generator=None, init=True):
        self._bound = flat
        return self.mlp.bind(self._node_block.input_dim(half_dim), flat, generator=generator, init=init)

    def _standalone(self, half_dim, device):
        if self._flow is None:
            mlp = self.mlp
            if mlp.output_dim != half_dim:
                raise ValueError(f"MLP output_dim={mlp.output_dim} must equal the node width {half_dim}")
            self._flow = _Flow(1, 2 * half_dim, mlp, weight_sharing=True, **self.flow_kwargs())
            per = self.gnn_param_count(half_dim)
            flat = torch.empty(4 * per, dtype=torch.float32, device=device)
            if getattr(self, "_bound", None) is None:
                self.bind_params(half_dim, flat[:per])
            else:                      # already bound elsewhere (e.g. adopted by a GRevNet): copy
                flat[:per].copy_(self._bound.detach().to(device))
            for k in range(1, 4):
                flat[k * per:(k + 1) * per].copy_(flat[:per])
            self.params = nn.Parameter(flat, requires_grad=False)
        return self._flow

    def forward(self, graph: GraphsTuple) -> GraphsTuple:
        nodes = _lib.require_cuda(graph.nodes, "graph.nodes", torch.float32)
        n, h = nodes.shape
        st = structure_of(graph)
        flow = self._standalone(h, nodes.device)
        handle = flow.ensure(self.params.detach())
        lib = flow.lib
        wsb = lib.gnf_grevnet_workspace(handle, n, _lib.MATH["fp32"])
        ws = _lib.workspace(wsb, nodes.device)
        out = torch.empty_like(nodes)
        _lib.check(lib.gnf_gnn_forward(handle, 0, 0, 0, _lib.ptr(nodes), n, st.n_edges, _lib.ptr(st.rowptr),
                                       _lib.ptr(st.csr_senders), _lib.ptr(out), _lib.ptr(ws), wsb,
                                       _lib.stream_ptr(nodes.device)), "gnf_gnn_forward")
        return graph.replace(nodes=out)


def avg_then_mlp_gnn(make_mlp_fn, epsilon):            # gnn.py:238-241
    return NodeBlockGNN(AggThenMLPBlock(unsorted_segment_mean, make_mlp_fn, epsilon))


def sum_then_mlp_gnn(make_mlp_fn, epsilon):            # gnn.py:244-247
    return NodeBlockGNN(AggThenMLPBlock(unsorted_segment_sum, make_mlp_fn, epsilon))


def sum_concat_then_mlp_gnn(make_mlp_fn):              # gnn.py:250-252
    return NodeBlockGNN(ConcatThenMLPBlock(unsorted_segment_sum, make_mlp_fn))


def avg_concat_then_mlp_gnn(make_mlp_fn):              # gnn.py:255-257
    return NodeBlockGNN(ConcatThenMLPBlock(unsorted_segment_mean, make_mlp_fn))


class DMSelfAttentionMLP(NodeBlockGNN):
    """DMSelfAttentionMLP (gnn.py:480-553), the default GNN of both GRevNet scripts
    (run_grevnet.py:56, train_grevnet_with_data.py:42): bias-free q/k/v projections, multi-head
    edge-softmax attention over each receiver's in-edges (DMSelfAttention, gnn.py:385-477; the value
    projection is shared by all heads, gnn.py:528), head-concat projection to
    `concat_heads_output_dim`, concat with the input, MLP, optional residual and snt.LayerNorm
    (gamma = 1, beta = 0 at init, eps 1e-5 [upstream]).  fp32 kernels."""

    def __init__(self, kq_dim, v_dim, make_mlp_fn, num_heads=8, concat_heads_output_dim=20, concat=True,
                 residual=False, layer_norm=False, kq_dim_division=False, name="dm_self_attention"):
        nn.Module.__init__(self)
        self.name = name
        self.kq_dim, self.v_dim, self.num_heads = int(kq_dim), int(v_dim), int(num_heads)
        self.concat_heads_output_dim = int(concat_heads_output_dim)
        self.concat, self.residual, self.layer_norm = bool(concat), bool(residual), bool(layer_norm)
        self.kq_dim_division = bool(kq_dim_division)
        self._mlp = make_mlp_fn()
        self._flow = None
        self.params = None
        self.wq = self.wk = self.wv = self.wo = None
        self.ln_gamma = self.ln_beta = None

    @property
    def mlp(self) -> MLP:
        return self._mlp

    def config(self):
        return ("dm_attn", self.kq_dim, self.v_dim, self.num_heads, self.concat_heads_output_dim, self.concat,
                self.residual, self.kq_dim_division, self.layer_norm) + self.mlp.signature()

    def flow_kwargs(self):
        flags = (_lib.ATTN_CONCAT if self.concat else 0) | (_lib.ATTN_RESIDUAL if self.residual else 0) | \
                (_lib.ATTN_KQ_DIV if self.kq_dim_division else 0) | (_lib.ATTN_LAYER_NORM if self.layer_norm else 0)
        return dict(block="dm_attn", agg="sum", eps=1.0,
                    attn=dict(num_heads=self.num_heads, kq_dim=self.kq_dim, v_dim=self.v_dim,
                              out_dim=self.concat_heads_output_dim, flags=flags))

    def mlp_input_dim(self, half_dim):
        return half_dim + self.concat_heads_output_dim if self.concat else self.concat_heads_output_dim

    def _attn_shapes(self, half_dim):
        qk, hv = self.num_heads * self.kq_dim, self.num_heads * self.v_dim
        return [(half_dim, qk), (half_dim, qk), (half_dim, self.v_dim), (hv, self.concat_heads_output_dim)]

    def gnn_param_count(self, half_dim):
        return sum(i * o for i, o in self._attn_shapes(half_dim)) + self.mlp.param_count(self.mlp_input_dim(half_dim)) + \
            (2 * half_dim if self.layer_norm else 0)

    def bind_params(self, half_dim, flat, generator=None, init=True):
        self._bound = flat
        off, mats = 0, []
        for k, (i, o) in enumerate(self._attn_shapes(half_dim)):
            w = flat[off:off + i * o].view(i, o)
            off += i * o
            if init:
                with torch.no_grad():
                    if k < 3:      # tf.contrib.layers.xavier_initializer(uniform=True)  gnn.py:504-506
                        lim = math.sqrt(6.0 / (i + o))
                        w.uniform_(-lim, lim, generator=generator)
                    else:          # snt.Linear default: truncated normal, stddev 1/sqrt(input_size)
                        _truncated_normal_(w, 1.0 / math.sqrt(i), generator)
            mats.append(w)
        self.wq, self.wk, self.wv, self.wo = mats
        off += self.mlp.bind(self.mlp_input_dim(half_dim), flat[off:], generator=generator, init=init)
        if self.layer_norm:                      # snt.LayerNorm variables, created after the MLP's (gnn.py:554-556)
            self.ln_gamma, self.ln_beta = flat[off:off + half_dim], flat[off + half_dim:off + 2 * half_dim]
            if init:
                with torch.no_grad():
                    self.ln_gamma.fill_(1.0)
                    self.ln_beta.zero_()
            off += 2 * half_dim
        return off


def dm_self_attn_gnn(kq_dim, v_dim, make_mlp_fn, num_heads, concat_heads_output_dim, concat=True, residual=False,
                     layer_norm=False, kq_dim_division=False):            # gnn.py:555-573
    return DMSelfAttentionMLP(kq_dim=kq_dim, v_dim=v_dim, make_mlp_fn=make_mlp_fn, num_heads=num_heads,
                              concat_heads_output_dim=concat_heads_output_dim, concat=concat, residual=residual,
                              layer_norm=layer_norm, kq_dim_division=kq_dim_division)


def get_gnns(num_timesteps, make_gnn_fn):              # gnn.py:266-267
    return [make_gnn_fn() for _ in range(num_timesteps)]


# --------------------------------------------------------------------------------------------
# GRevNet  (gnn.py:273-381)
# --------------------------------------------------------------------------------------------
class GRevNet(nn.Module):
    """Affine-coupling flow over node features.

    grevnet(graph, inverse=True)  -> (GraphsTuple z, scalar log_det_jacobian)   [GRevNet.f]
    grevnet(graph, inverse=False) -> GraphsTuple x                               [GRevNet.g]

    NOTE the reference's naming: inverse=True is the data -> latent (density) direction
    (gnn.py:379-381).

    `math` selects the arithmetic of the MLP contraction: "tc3x" (tcgen05, fp16 hi/lo split,
    fp32 accumulate; the default when the shape is supported), "fp32" (FFMA, any shape),
    "bf16" (tcgen05 single pass), "tc3x_bf16", "tc2x".

    fp16 range guard: the reference computes in fp32 and never clamps exp(s) (gnn.py:323), so node
    features / hidden activations beyond the fp16 range (|v| > 65504) are legal inputs, but the fp16
    hi/lo split of "tc3x"/"tc2x" turns them into inf/NaN.  The kernels raise a sticky device flag when
    that happens; it is copied back asynchronously after every call and `check_numerics()` (called
    without blocking at the start of every later call, or by the user with block=True) raises
    FloatingPointError.  With math=None (auto) the module then switches to "tc3x_bf16" (same 3-MMA
    scheme, fp32 exponent range, 2^-17 instead of 2^-22 operand error) for all later calls.
    """

    def __init__(self, make_gnn_fn: Callable[[], NodeBlockGNN], num_timesteps, node_embedding_dim,
                 use_batch_norm=False, weight_sharing=False, name="GRevNet", math: Optional[str] = None,
                 device=None, seed: Optional[int] = None):
        super().__init__()
        self.name = name
        self.num_timesteps = int(num_timesteps)
        self.weight_sharing = bool(weight_sharing)
        self.node_embedding_dim = int(node_embedding_dim)
        if self.node_embedding_dim % 2:
            raise ValueError("node_embedding_dim must be even (tf.split at gnn.py:306)")
        self.use_batch_norm = bool(use_batch_norm)
        T = self.num_timesteps
        # construction order of gnn.py:283-299
        if self.weight_sharing:
            self.s = nn.ModuleList([make_gnn_fn(), make_gnn_fn()])
            self.t = nn.ModuleList([make_gnn_fn(), make_gnn_fn()])
            order = [self.s[0], self.s[1], self.t[0], self.t[1]]
        else:
            self.s = nn.ModuleList([nn.ModuleList(get_gnns(T, make_gnn_fn)), nn.ModuleList(get_gnns(T, make_gnn_fn))])
            self.t = nn.ModuleList([nn.ModuleList(get_gnns(T, make_gnn_fn)), nn.ModuleList(get_gnns(T, make_gnn_fn))])
            order = [g for which in (self.s, self.t) for half in which for g in half]
        cfgs = {g.config() for g in order}
        if len(cfgs) != 1:
            raise ValueError("make_gnn_fn must return identically configured GNNs")
        g0 = order[0]
        if not isinstance(g0, NodeBlockGNN):
            raise TypeError("make_gnn_fn must return a NodeBlockGNN (sum/avg concat_then_mlp or then_mlp) "
                            "or a dm_self_attn_gnn")
        H = self.node_embedding_dim // 2
        mlp = g0.mlp
        if mlp.output_dim != H:
            raise ValueError(f"MLP output_dim={mlp.output_dim} must be node_embedding_dim/2={H}")
        self._flow = _Flow(T, self.node_embedding_dim, mlp, weight_sharing=self.weight_sharing, **g0.flow_kwargs())
        in_dim = H
        per = g0.gnn_param_count(H)
        assert per * len(order) == self._flow.param_count, (per, len(order), self._flow.param_count)
        dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        gen = None
        if seed is not None:
            gen = torch.Generator(device="cpu")
            gen.manual_seed(int(seed))
        flat = torch.empty(self._flow.param_count, dtype=torch.float32)
        for k, g in enumerate(order):          # flat order: which -> half -> step (include/gnf_b200.h)
            g.bind_params(H, flat[k * per:(k + 1) * per], generator=gen)
        self.params = nn.Parameter(flat.to(dev), requires_grad=False)
        self._order = order
        self._per = per
        self._in_dim = in_dim
        self._rebind()
        self._math = math
        self._range_host = None               # pinned int32[1]: last copy of the flow's fp16 range flag
        self._range_event = None              # recorded after that copy
        self._range_tripped = False           # auto mode: an overflow was seen -> tc3x_bf16 from then on
        # bns[2][T] = make_batch_norm() (gnn.py:260-263,301-302): tf.layers.BatchNormalization(axis=-1)
        # defaults gamma=1, beta=0, moving_mean=0, moving_variance=1, epsilon=1e-3, momentum=0.99,
        # wrapped in tfb.BatchNormalization(training=True).  Always constructed, as in the reference.
        self.bn_epsilon, self.bn_momentum = 1e-3, 0.99
        # trainable, like the tf.layers variables they stand for, when the bijector is in use (gradients come from
        # _backward_bn); the gamma constraint relu(gamma)+1e-6 is a Keras constraint = a projection after each update
        self.bn_gamma = nn.Parameter(torch.ones(2, T, H, device=dev), requires_grad=bool(use_batch_norm))
        self.bn_beta = nn.Parameter(torch.zeros(2, T, H, device=dev), requires_grad=bool(use_batch_norm))
        self.register_buffer("bn_moving_mean", torch.zeros(2, T, H, device=dev))
        self.register_buffer("bn_moving_var", torch.ones(2, T, H, device=dev))
        self.bn_update_moving = True          # the scripts run UPDATE_OPS with every step (run_grevnet.py:362)
        self._bn_saved = {}                   # (half, step) -> (mean, var, N) of the last density pass
        self.bn_group = None                  # process group for cross-rank batch statistics (sharding.py)
        self.bn_sync = True                   # all-reduce the batch statistics when running sharded
        self.bn_chain = True                  # single-rank batches: whole batch-norm flow in one library call

    # -- parameter plumbing -------------------------------------------------------------------
    def _rebind(self):
        for k, g in enumerate(self._order):
            g.bind_params(self._in_dim, self.params.detach()[k * self._per:(k + 1) * self._per], init=False)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._rebind()
        return out

    def mlp_of(self, which: str, half: int, step: int) -> MLP:
        grp = self.s if which == "s" else self.t
        return (grp[half] if self.weight_sharing else grp[half][step]).mlp

    def scale_last_layers_(self, factor: float):
        """Damp the last layer of every MLP (synthetic benchmarks only: keeps the reference's
        unclamped exp(s), gnn.py:323, finite; SURVEY §7 hard part 3)."""
        with torch.no_grad():
            for g in self._order:
                g.mlp.weights[-1].mul_(factor)
                g.mlp.biases[-1].mul_(factor)
        return self

    @property
    def math(self) -> str:
        if self._math is not None:
            return self._math
        self._flow.ensure(self.params.detach())
        if not self._flow.supports("tc3x"):
            return "fp32"
        return "tc3x_bf16" if self._range_tripped else "tc3x"

    @math.setter
    def math(self, value):
        if value is not None and value not in _lib.MATH:
            raise ValueError(f"math must be one of {sorted(_lib.MATH)}")
        self._math = value

    # -- fp16 range guard -------------------------------------------------------------------------
    def _arm_range_guard(self, handle, math_name: str, device):
        """After a tc3x / tc2x call: queue an async copy of the sticky device flag + an event."""
        if math_name not in ("tc3x", "tc2x") or torch.cuda.is_current_stream_capturing():
            return
        if self._range_host is None:
            self._range_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        _lib.check(self._flow.lib.gnf_flow_range_flag(handle, C.c_void_p(self._range_host.data_ptr()), 0,
                                                      _lib.stream_ptr(device)), "gnf_flow_range_flag")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._range_event = ev

    def check_numerics(self, block: bool = True):
        """Raise FloatingPointError if an earlier "tc3x"/"tc2x" call left the fp16 range.  block=False only
        looks at copies that have already landed (no host sync)."""
        ev = self._range_event
        if ev is None:
            return
        if block:
            ev.synchronize()
        elif not ev.query():
            return
        self._range_event = None
        if int(self._range_host[0]) == 0:
            return
        self._range_host[0] = 0
        if self._flow.handle is not None:       # clear the sticky flag (stream ordered)
            self._flow.lib.gnf_flow_range_flag(self._flow.handle, None, 1, _lib.stream_ptr(self.params.device))
        self._range_tripped = True
        raise FloatingPointError(
            "GRevNet: an MLP input or hidden activation exceeded the fp16 range (|v| > 65504) in an earlier "
            "'tc3x'/'tc2x' call; its outputs contain inf/NaN where the reference's fp32 arithmetic stays finite. "
            "Re-run that batch: math=None (auto) has switched to 'tc3x_bf16', or set math='tc3x_bf16' / 'fp32'.")

    # -- the hot path ---------------------------------------------------------------------------
    def _run(self, x: GraphsTuple, inverse_kernel: bool):
        if self._range_event is not None and not torch.cuda.is_current_stream_capturing():
            self.check_numerics(block=False)                  # (cudaEventQuery would invalidate a capture)
        if self.use_batch_norm:
            return self._run_bn(x, inverse_kernel)
        nodes = _lib.require_cuda(x.nodes, "graph.nodes", torch.float32)
        if nodes.dim() != 2 or nodes.shape[1] != self.node_embedding_dim:
            raise ValueError(f"graph.nodes must be [N, {self.node_embedding_dim}], got {tuple(nodes.shape)}")
        if nodes.device != self.params.device:
            raise RuntimeError(f"graph is on {nodes.device} but the GRevNet parameters are on {self.params.device}")
        st = structure_of(x)
        handle = self._flow.ensure(self.params.detach())
        lib = self._flow.lib
        math_name = self.math
        m = _lib.MATH[math_name]
        n = nodes.shape[0]
        wsb = lib.gnf_grevnet_workspace(handle, n, m)
        ws = _lib.workspace(wsb, nodes.device)
        out = torch.empty_like(nodes)
        stream = _lib.stream_ptr(nodes.device)
        if not inverse_kernel:
            ldj = torch.empty(1, dtype=torch.float64, device=nodes.device)
            _lib.check(lib.gnf_grevnet_forward(handle, _lib.ptr(nodes), n, st.n_edges, _lib.ptr(st.rowptr),
                                               _lib.ptr(st.csr_senders), _lib.ptr(out), _lib.ptr(ldj), m,
                                               _lib.ptr(ws), wsb, stream), "gnf_grevnet_forward")
            self._arm_range_guard(handle, math_name, nodes.device)
            return out, ldj
        _lib.check(lib.gnf_grevnet_inverse(handle, _lib.ptr(nodes), n, st.n_edges, _lib.ptr(st.rowptr),
                                           _lib.ptr(st.csr_senders), _lib.ptr(out), m, _lib.ptr(ws), wsb,
                                           stream), "gnf_grevnet_inverse")
        self._arm_range_guard(handle, math_name, nodes.device)
        return out, None

    # -- a9: the batch-norm variant, half step by half step ---------------------------------------
    def _bn_inverse_(self, xp, n_local, half, i, ldj):
        """bn.inverse + bn.inverse_log_det_jacobian(x, 2) on a planar half, in place (gnn.py:310-313):
        batch statistics over ALL nodes of the batch (all ranks when sharded).  Three launches on the device
        (moments, [H]-sized bookkeeping, affine) and, when sharded, one all-reduce of the [2H+1] sums."""
        lib, H, dev = self._flow.lib, self.node_embedding_dim // 2, xp.device
        stream = _lib.stream_ptr(dev)
        sums = torch.empty(2 * H + 1, dtype=torch.float64, device=dev)
        wsb = lib.gnf_bn_moments_workspace(H)
        ws = _lib.workspace(wsb, dev)
        _lib.check(lib.gnf_bn_moments(_lib.ptr(xp), n_local, H, _lib.ptr(sums), _lib.ptr(ws), wsb, stream), "gnf_bn_moments")
        dist = torch.distributed
        if self.bn_sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.bn_group) > 1:
            dist.all_reduce(sums, group=self.bn_group)                   # the extra [2H+1] all-reduce of row a9
        scale_shift = torch.empty(2 * H, dtype=torch.float32, device=dev)
        stats = torch.empty(2 * H + 1, dtype=torch.float64, device=dev)  # mean, var, N: the backward undoes the bijector with these
        gamma, beta = self.bn_gamma.detach()[half, i], self.bn_beta.detach()[half, i]
        mm, mv = self.bn_moving_mean[half, i], self.bn_moving_var[half, i]
        _lib.check(lib.gnf_bn_finalize(_lib.ptr(sums), H, _lib.ptr(gamma), _lib.ptr(beta), float(self.bn_epsilon),
                                       float(n_local), _lib.ptr(ldj), _lib.ptr(scale_shift), _lib.ptr(stats),
                                       _lib.ptr(mm), _lib.ptr(mv),
                                       float(self.bn_momentum) if self.bn_update_moving else -1.0, stream), "gnf_bn_finalize")
        self._bn_saved[(half, i)] = stats
        _lib.check(lib.gnf_affine_rows(_lib.ptr(xp), n_local, H, _lib.ptr(scale_shift), _lib.ptr(scale_shift[H:]), stream),
                   "gnf_affine_rows")

    def _bn_forward_(self, zp, n_local, half, i):
        """bn.forward = de-normalise with the MOVING statistics (gnn.py:356-358,369-371)."""
        lib, H, dev = self._flow.lib, self.node_embedding_dim // 2, zp.device
        scale = torch.sqrt(self.bn_moving_var[half, i] + self.bn_epsilon) / self.bn_gamma.detach()[half, i]
        shift = self.bn_moving_mean[half, i] - self.bn_beta.detach()[half, i] * scale
        sc, sh = scale.float().contiguous(), shift.float().contiguous()
        _lib.check(lib.gnf_affine_rows(_lib.ptr(zp), n_local, H, _lib.ptr(sc), _lib.ptr(sh), _lib.stream_ptr(dev)),
                   "gnf_affine_rows")

    def _run_bn(self, x: GraphsTuple, inverse_kernel: bool):
        nodes = _lib.require_cuda(x.nodes, "graph.nodes", torch.float32)
        D = self.node_embedding_dim
        if nodes.dim() != 2 or nodes.shape[1] != D:
            raise ValueError(f"graph.nodes must be [N, {D}], got {tuple(nodes.shape)}")
        st = structure_of(x)
        handle = self._flow.ensure(self.params.detach())
        lib, dev = self._flow.lib, nodes.device
        math_name = self.math
        m = _lib.MATH[math_name]
        n = nodes.shape[0]
        hp = lib.gnf_padded_half(D // 2)
        wsb = lib.gnf_grevnet_workspace(handle, n, m)
        ws = _lib.workspace(wsb, dev)
        stream = _lib.stream_ptr(dev)
        dist = torch.distributed
        sharded = self.bn_sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.bn_group) > 1
        if self.bn_chain and not sharded:
            # single-rank batch: the whole chain in one library call (no per-half-step host work)
            H, T = D // 2, self.num_timesteps
            bwsb = lib.gnf_grevnet_bn_workspace(handle)
            bws = _lib.workspace(bwsb, dev)
            out = torch.empty_like(nodes)
            gamma, beta = self.bn_gamma.detach(), self.bn_beta.detach()
            for t in (gamma, beta, self.bn_moving_mean, self.bn_moving_var):
                _lib.require_cuda(t, "batch-norm parameter", torch.float32)
            if not inverse_kernel:
                ldj = torch.empty(1, dtype=torch.float64, device=dev)
                stats = torch.empty(2, T, 2 * H + 1, dtype=torch.float64, device=dev)
                _lib.check(lib.gnf_grevnet_forward_bn(
                    handle, _lib.ptr(nodes), n, st.n_edges, _lib.ptr(st.rowptr), _lib.ptr(st.csr_senders), _lib.ptr(gamma),
                    _lib.ptr(beta), _lib.ptr(self.bn_moving_mean), _lib.ptr(self.bn_moving_var), float(self.bn_epsilon),
                    float(self.bn_momentum) if self.bn_update_moving else -1.0, _lib.ptr(out), _lib.ptr(ldj),
                    _lib.ptr(stats), m, _lib.ptr(ws), wsb, _lib.ptr(bws), bwsb, stream), "gnf_grevnet_forward_bn")
                for half in range(2):
                    for i in range(T):
                        self._bn_saved[(half, i)] = stats[half, i]
                self._arm_range_guard(handle, math_name, dev)
                return out, ldj
            _lib.check(lib.gnf_grevnet_inverse_bn(
                handle, _lib.ptr(nodes), n, st.n_edges, _lib.ptr(st.rowptr), _lib.ptr(st.csr_senders), _lib.ptr(gamma),
                _lib.ptr(beta), _lib.ptr(self.bn_moving_mean), _lib.ptr(self.bn_moving_var), float(self.bn_epsilon),
                _lib.ptr(out), m, _lib.ptr(ws), wsb, _lib.ptr(bws), bwsb, stream), "gnf_grevnet_inverse_bn")
            self._arm_range_guard(handle, math_name, dev)
            return out, None
        x0 = torch.empty(max(n, 1), hp, dtype=torch.float32, device=dev)
        x1 = torch.empty(max(n, 1), hp, dtype=torch.float32, device=dev)
        _lib.check(lib.gnf_split_halves(_lib.ptr(nodes), n, D, _lib.ptr(x0), _lib.ptr(x1), stream), "gnf_split_halves")
        halves = (x0, x1)

        def half_step(half, i, inverse, ldj):
            xa, xb = halves[half], halves[1 - half]
            _lib.check(lib.gnf_coupling_half(handle, half, i, int(inverse), _lib.ptr(xa), _lib.ptr(xb), n, st.n_edges,
                                             _lib.ptr(st.rowptr), _lib.ptr(st.csr_senders), _lib.ptr(ldj), m,
                                             _lib.ptr(ws), wsb, stream), "gnf_coupling_half")

        ldj = None
        if not inverse_kernel:
            ldj = torch.zeros(1, dtype=torch.float64, device=dev)
            for i in range(self.num_timesteps):                      # gnn.py:309-338
                self._bn_inverse_(x0, n, 0, i, ldj)
                half_step(0, i, False, ldj)
                self._bn_inverse_(x1, n, 1, i, ldj)
                half_step(1, i, False, ldj)
        else:
            for i in reversed(range(self.num_timesteps)):            # gnn.py:347-372
                half_step(1, i, True, None)
                self._bn_forward_(x1, n, 1, i)
                half_step(0, i, True, None)
                self._bn_forward_(x0, n, 0, i)
        out = torch.empty_like(nodes)
        _lib.check(lib.gnf_merge_halves(_lib.ptr(x0), _lib.ptr(x1), n, D, _lib.ptr(out), stream), "gnf_merge_halves")
        self._arm_range_guard(handle, math_name, dev)
        return out, ldj

    def f(self, x: GraphsTuple):
        """x -> (z, log_det_jacobian)   (gnn.py:304-341)"""
        z, ldj = self._run(x, inverse_kernel=False)
        self._last_ldj64 = ldj
        return x.replace(nodes=z), ldj[0].to(torch.float32)

    def f64(self, x: GraphsTuple):
        """As f, but returns the log-det as the device float64 [1] tensor the kernels produce."""
        z, ldj = self._run(x, inverse_kernel=False)
        return x.replace(nodes=z), ldj

    def g(self, z: GraphsTuple):
        """z -> x   (gnn.py:343-373)"""
        x, _ = self._run(z, inverse_kernel=True)
        return z.replace(nodes=x)

    # -- f2: training-step gradients (reversible, nothing stored by the forward pass) --------------
    def backward_from_z(self, graph: GraphsTuple, z_nodes: torch.Tensor, loss_scale: float,
                        grads: Optional[torch.Tensor] = None, return_x: bool = False, math: Optional[str] = None):
        """d(-loss_scale * log_prob_xs)/d(params), accumulated into `grads` (flat, same layout as
        self.params).  `z_nodes` must be f(graph).nodes.  `math` defaults to the forward's mode:
        "fp32" = layered FFMA kernels, anything else = the tcgen05 backward (backward_tc.cu)."""
        if self.use_batch_norm:
            return self._backward_bn(graph, z_nodes, loss_scale, grads, return_x, math)
        lib = self._flow.lib
        _lib.require_cuda(z_nodes, "z.nodes", torch.float32)
        handle = self._flow.ensure(self.params.detach())
        st, stt = structure_of(graph), transposed_structure_of(graph)
        n, dev = z_nodes.shape[0], z_nodes.device
        if grads is None:
            grads = torch.zeros_like(self.params.detach())
        _lib.require_cuda(grads, "grads", torch.float32)
        x_out = torch.empty_like(z_nodes) if return_x else None
        m = _lib.MATH[self._backward_math(math)]
        wsb = lib.gnf_grevnet_backward_workspace(handle, n, m)
        ws = _lib.workspace(wsb, dev)
        _lib.check(lib.gnf_grevnet_backward(handle, _lib.ptr(z_nodes), n, st.n_edges, _lib.ptr(st.rowptr),
                                            _lib.ptr(st.csr_senders), _lib.ptr(stt.rowptr), _lib.ptr(stt.csr_senders),
                                            float(loss_scale), _lib.ptr(grads), _lib.ptr(x_out), m, _lib.ptr(ws), wsb,
                                            _lib.stream_ptr(dev)), "gnf_grevnet_backward")
        self._last_backward_workspace = ws if getattr(self, "_keep_backward_workspace", False) else None
        return (grads, x_out) if return_x else grads

    def _backward_math(self, math: Optional[str]) -> str:
        """Arithmetic of the backward: an explicit `math` is taken as is (and rejected by the library if the shape
        does not support it); otherwise the forward's mode when the tensor-core backward serves this flow, else fp32."""
        if math is not None:
            return math
        name = self.math
        return name if self._flow.supports_backward(name) else "fp32"

    def _backward_bn(self, graph, z_nodes, loss_scale, grads, return_x, math):
        """backward_from_z with use_batch_norm=True: the reversed half steps (gnf_coupling_half_backward)
        interleaved with the backward of the TFP batch-norm bijector in training mode.  The bijector is undone
        with the batch statistics the density pass saved (`f` must have been called on this batch); its gamma
        and beta get gradients in `self.bn_gamma.grad` / `self.bn_beta.grad`:
            y = gamma * xhat + beta,  xhat = (x - mu) / s,  s = sqrt(var + eps),  ldj += N (log gamma - log s)
            dL/dbeta = S1,  dL/dgamma = S2 - loss_scale * N / gamma          (S1 = sum G_y, S2 = sum G_y xhat)
            dL/dx = (gamma G_y - gamma S1/N - xhat gamma S2/N) / s + loss_scale * xhat / s
        Sharded runs all-reduce (S1, S2) like the forward's moments."""
        lib = self._flow.lib
        _lib.require_cuda(z_nodes, "z.nodes", torch.float32)
        T, D = self.num_timesteps, self.node_embedding_dim
        H = D // 2
        if len(self._bn_saved) != 2 * T:
            raise RuntimeError("backward with use_batch_norm=True needs the batch statistics of the density pass: "
                               "call f(graph) (or loss_and_grad) on this batch first")
        handle = self._flow.ensure(self.params.detach())
        st, stt = structure_of(graph), transposed_structure_of(graph)
        n, dev = z_nodes.shape[0], z_nodes.device
        if grads is None:
            grads = torch.zeros_like(self.params.detach())
        _lib.require_cuda(grads, "grads", torch.float32)
        m = _lib.MATH[self._backward_math(math)]
        stream = _lib.stream_ptr(dev)
        hp = lib.gnf_padded_half(H)
        wsb = lib.gnf_grevnet_backward_workspace(handle, n, m)
        ws = _lib.workspace(wsb, dev)
        bwsb = lib.gnf_bn_moments_workspace(H)
        bws = _lib.workspace(bwsb, dev)
        x = [torch.empty(max(n, 1), hp, dtype=torch.float32, device=dev) for _ in range(2)]
        _lib.check(lib.gnf_split_halves(_lib.ptr(z_nodes), n, D, _lib.ptr(x[0]), _lib.ptr(x[1]), stream), "gnf_split_halves")
        g = [x[0] * float(loss_scale), x[1] * float(loss_scale)]          # dL/dz = loss_scale * z
        g_gamma = torch.zeros_like(self.bn_gamma.detach(), dtype=torch.float64)
        g_beta = torch.zeros_like(self.bn_beta.detach(), dtype=torch.float64)
        dist = torch.distributed
        sync = self.bn_sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.bn_group) > 1

        coef = torch.empty(7 * H, dtype=torch.float32, device=dev)
        inv_gamma = torch.empty(H, dtype=torch.float32, device=dev)
        sums = torch.empty(2 * H, dtype=torch.float64, device=dev)

        def bn_backward(half, i):
            stats = self._bn_saved[(half, i)]
            gamma, beta = self.bn_gamma.detach()[half, i], self.bn_beta.detach()[half, i]
            _lib.check(lib.gnf_bn_backward_coef(None, None, H, _lib.ptr(gamma), _lib.ptr(beta), float(self.bn_epsilon),
                                                float(loss_scale), None, _lib.ptr(inv_gamma), None, None, stream),
                       "gnf_bn_backward_coef")
            _lib.check(lib.gnf_bn_backward_sums(_lib.ptr(x[half]), _lib.ptr(g[half]), n, H, _lib.ptr(beta), _lib.ptr(inv_gamma),
                                                _lib.ptr(sums), _lib.ptr(bws), bwsb, stream), "gnf_bn_backward_sums")
            if sync:
                dist.all_reduce(sums, group=self.bn_group)
            _lib.check(lib.gnf_bn_backward_coef(_lib.ptr(sums), _lib.ptr(stats), H, _lib.ptr(gamma), _lib.ptr(beta),
                                                float(self.bn_epsilon), float(loss_scale), _lib.ptr(coef), None,
                                                _lib.ptr(g_gamma[half, i]), _lib.ptr(g_beta[half, i]), stream),
                       "gnf_bn_backward_coef")
            _lib.check(lib.gnf_bn_backward_apply(_lib.ptr(x[half]), _lib.ptr(g[half]), n, H, _lib.ptr(coef), stream),
                       "gnf_bn_backward_apply")

        def half_backward(half, i):
            a, b = half, 1 - half
            _lib.check(lib.gnf_coupling_half_backward(handle, half, i, _lib.ptr(x[a]), _lib.ptr(x[b]), _lib.ptr(g[a]),
                                                      _lib.ptr(g[b]), n, st.n_edges, _lib.ptr(st.rowptr),
                                                      _lib.ptr(st.csr_senders), _lib.ptr(stt.rowptr),
                                                      _lib.ptr(stt.csr_senders), float(loss_scale), _lib.ptr(grads), m,
                                                      _lib.ptr(ws), wsb, stream), "gnf_coupling_half_backward")

        for i in reversed(range(T)):                                  # gnn.py:309-338 reversed
            half_backward(1, i)
            bn_backward(1, i)
            half_backward(0, i)
            bn_backward(0, i)
        # a sharded caller all-reduces parameter gradients itself; gamma/beta data terms were already global sums
        self.bn_gamma.requires_grad_(True)
        self.bn_beta.requires_grad_(True)
        self.bn_gamma.grad = g_gamma.float()
        self.bn_beta.grad = g_beta.float()
        if return_x:
            x_out = torch.empty_like(z_nodes)
            _lib.check(lib.gnf_merge_halves(_lib.ptr(x[0]), _lib.ptr(x[1]), n, D, _lib.ptr(x_out), stream), "gnf_merge_halves")
            return grads, x_out
        return grads

    def loss_and_grad(self, graph: GraphsTuple, per_node: bool = True, backward_math: Optional[str] = None):
        """One training-step evaluation: the scalars of run_grevnet.py:292-302 and the gradient of
        total_loss (per_node=False, run_grevnet.py:362-364) or loss_per_node (per_node=True,
        train_grevnet_with_data.py:353-380) w.r.t. the flat parameter vector; also stored in
        self.params.grad for torch optimisers."""
        from .loss import mvn_log_prob_sum, scalars_from_vector
        z, ldj64 = self.f64(graph)
        out = scalars_from_vector(mvn_log_prob_sum(z.nodes, ldj64))
        n = max(int(z.nodes.shape[0]), 1)
        grads = self.backward_from_z(graph, z.nodes, 1.0 / n if per_node else 1.0, math=backward_math)
        self.params.grad = grads
        out["z"] = z
        return out, grads

    def log_prob(self, x: GraphsTuple):
        """sum(prior.log_prob(z)) + log_det_jacobian with prior = N(0, I): what the dead
        GRevNet.log_prob (gnn.py:375-377, self.prior is never set) was meant to compute and the
        scripts inline (run_grevnet.py:291-295)."""
        from .loss import log_prob as _log_prob
        return _log_prob(self, x)["log_prob_xs"]

    def forward(self, input: GraphsTuple, inverse=True):
        func = self.f if inverse else self.g      # gnn.py:379-381
        return func(input)
