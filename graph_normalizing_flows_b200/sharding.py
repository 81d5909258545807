"""Graph-sharded data parallelism for the GRevNet hot path.

Graphs in a packed batch never exchange data (gather / segment indices stay inside a graph's
node range, the MLPs are row-wise, log-det and log-prob are plain sums; gnn.py:322,337 and
run_grevnet.py:294), so whole graphs are assigned to ranks, each rank runs the fused kernels
on its own packed sub-batch with replicated weights, and ONE all-reduce(SUM) of the fp64
vector (log_prob_zs, log_det_jacobian, log_prob_xs, num_nodes) assembles the batch
log-likelihood.  z stays sharded.  No data-path collective.  (use_batch_norm=True would couple
graphs through batch statistics and is not on this path.)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .graphs import GraphsTuple


def graph_costs(n_node, n_edge, flops_per_node_update: float = 807936.0, alpha: float = 1.0,
                beta: float = 64.0) -> np.ndarray:
    """Per-graph cost model alpha * n_node * F + beta * n_edge (MLP FLOPs dominate; the edge term
    keeps very dense graphs from piling up on one rank)."""
    return alpha * np.asarray(n_node, np.float64) * flops_per_node_update + beta * np.asarray(n_edge, np.float64)


def partition_graphs(n_node: Sequence[int], n_edge: Sequence[int], world_size: int, **cost_kw) -> List[np.ndarray]:
    """Greedy longest-processing-time assignment of whole graphs to ranks.  Deterministic (stable
    sort, lowest rank wins ties); each rank's graph ids are returned in ascending order."""
    cost = graph_costs(n_node, n_edge, **cost_kw)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world_size, np.float64)
    buckets: List[List[int]] = [[] for _ in range(world_size)]
    for gidx in order:
        r = int(np.argmin(load))
        buckets[r].append(int(gidx))
        load[r] += cost[gidx]
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def shard_graphs_tuple(graph: GraphsTuple, graph_ids: np.ndarray) -> GraphsTuple:
    """Sub-batch of the selected graphs (host arrays), indices re-based to the shard's own packed
    node array -- integer-exact, edge order inside every graph preserved."""
    n_node = np.asarray(graph.n_node, np.int64)
    n_edge = np.asarray(graph.n_edge, np.int64)
    node_off = np.concatenate([[0], np.cumsum(n_node)])
    edge_off = np.concatenate([[0], np.cumsum(n_edge)])
    ids = np.asarray(graph_ids, np.int64)
    if len(ids) == 0:
        d = graph.nodes.shape[1] if graph.nodes is not None else 0
        return GraphsTuple(np.zeros((0, d), np.float32), None, np.zeros(0, np.int32), np.zeros(0, np.int32), None,
                           np.zeros(0, np.int32), np.zeros(0, np.int32))
    node_idx = np.concatenate([np.arange(node_off[g], node_off[g + 1]) for g in ids])
    edge_idx = np.concatenate([np.arange(edge_off[g], edge_off[g + 1]) for g in ids])
    new_off = np.concatenate([[0], np.cumsum(n_node[ids])[:-1]])
    shift = np.repeat(new_off - node_off[ids], n_edge[ids])
    senders = (np.asarray(graph.senders, np.int64)[edge_idx] + shift).astype(np.int32)
    receivers = (np.asarray(graph.receivers, np.int64)[edge_idx] + shift).astype(np.int32)
    nodes = np.asarray(graph.nodes)[node_idx] if graph.nodes is not None else None
    return GraphsTuple(nodes=nodes, edges=None, receivers=receivers, senders=senders, globals=None,
                       n_node=n_node[ids].astype(np.int32), n_edge=n_edge[ids].astype(np.int32))


def all_reduce_log_prob(vec4: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place SUM of (log_prob_zs, ldj, log_prob_xs, num_nodes) over the ranks (fp64).  NCCL on
    device tensors, gloo on host tensors (tests)."""
    if vec4.dtype != torch.float64 or vec4.numel() != 4:
        raise ValueError("expected a float64 4-vector")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec4, op=dist.ReduceOp.SUM, group=group)
    return vec4


class PeerComm:
    """The path's one collective over NVLink peer memory (include/gnf_b200.h section e): every rank exports a small
    slot buffer with cudaIpc, peers store their 4-vector into it, the sum is taken in rank order.  One node, world <= 8.

        comm = PeerComm(group)                 # collective: every rank of the group must construct it
        comm.all_reduce_(vec4)                 # device float64 [4], in place, one one-warp kernel on the current stream

    Construction raises (RuntimeError) where peer access is not available; GraphShardedGRevNet then keeps NCCL."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device=None):
        import ctypes as C
        import socket
        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        handle = (C.c_uint8 * 64)()
        peer = C.c_void_p()
        err = None
        with torch.cuda.device(self.device):
            rc = self.lib.gnf_peer_create(C.byref(peer), self.rank, self.world, handle) if self.world <= 8 else -1
        if rc != 0:
            err = _lib.last_error() if self.world <= 8 else "more than 8 ranks"
        mine = (socket.gethostname(), bytes(handle) if rc == 0 else None)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)              # the exchange IS the host-side all-gather
        self._peer = peer if rc == 0 else None
        if err is None and any(h is None for _, h in everyone):
            err = "a peer could not export its buffer"
        if err is None and len({host for host, _ in everyone}) != 1:
            err = "ranks span several hosts"
        if err is None:
            blob = b"".join(h for _, h in everyone)
            with torch.cuda.device(self.device):
                if self.lib.gnf_peer_connect(self._peer, blob) != 0:
                    err = _lib.last_error()
        ok = torch.tensor([0 if err else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # all or nothing: every rank takes the same path
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer-memory all-reduce unavailable: {err or 'a peer failed'}")

    def all_reduce_(self, vec4: torch.Tensor) -> torch.Tensor:
        from . import _lib
        if vec4.dtype != torch.float64 or vec4.numel() != 4 or not vec4.is_cuda:
            raise ValueError("expected a CUDA float64 4-vector")
        _lib.check(self.lib.gnf_peer_allreduce4(self._peer, _lib.ptr(vec4), _lib.stream_ptr(vec4.device)),
                   "gnf_peer_allreduce4")
        return vec4

    def log_prob_all_reduce(self, z_nodes: torch.Tensor, ldj64: Optional[torch.Tensor]) -> torch.Tensor:
        """gnf_log_prob and the all-reduce of its 4-vector in ONE kernel: the global scalars on every rank."""
        from . import _lib
        _lib.require_cuda(z_nodes, "z.nodes", torch.float32)
        n, d = z_nodes.shape
        out = torch.empty(4, dtype=torch.float64, device=z_nodes.device)
        wsb = self.lib.gnf_log_prob_workspace(n, d)
        ws = _lib.workspace(wsb, z_nodes.device)
        _lib.check(self.lib.gnf_log_prob_allreduce(_lib.ptr(z_nodes), n, d, _lib.ptr(ldj64), _lib.ptr(out), _lib.ptr(ws),
                                                   wsb, self._peer, _lib.stream_ptr(z_nodes.device)),
                   "gnf_log_prob_allreduce")
        return out

    def close(self):
        if getattr(self, "_peer", None) is not None:
            self.lib.gnf_peer_destroy(self._peer)
            self._peer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PendingLogProb:
    """Result of GraphShardedGRevNet.log_prob_async: the local latent and the (in flight) all-reduced 4-vector."""

    def __init__(self, vec, z, stream, t0, t1):
        self.vec, self.z, self._stream, self._t0, self._t1 = vec, z, stream, t0, t1

    def wait(self) -> dict:
        from .loss import scalars_from_vector
        if self._stream is not None:
            torch.cuda.current_stream(self.vec.device).wait_stream(self._stream)
            self._stream = None
        return scalars_from_vector(self.vec)

    def all_reduce_ms(self) -> float:
        """Device time between the all-reduce being reached on the side stream and its completion on this rank
        (includes waiting for the slowest rank); call after a synchronize."""
        return self._t0.elapsed_time(self._t1) if self._t0 is not None else 0.0


class GraphShardedGRevNet:
    """One process per GPU; weights replicated; graphs sharded.

        sharded = GraphShardedGRevNet(grevnet)                    # after init_process_group("nccl")
        local = sharded.local_shard(host_batch).to(device)        # this rank's graphs
        scalars = sharded.log_prob(local)                          # global batch scalars on every rank
    """

    def __init__(self, grevnet, group: Optional[dist.ProcessGroup] = None, peer_memory="auto"):
        """peer_memory: "auto" = the 4-vector all-reduce goes over NVLink peer memory (PeerComm) when every rank sits
        on one node with CUDA peer access and the backend is NCCL, else NCCL; True = require it; False = NCCL."""
        self.grevnet = grevnet
        self.group = group
        if hasattr(grevnet, "bn_group"):
            grevnet.bn_group = group          # batch-norm bijector: batch statistics / gradient sums over the same ranks
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world_size = dist.get_world_size(group) if on else 1
        self._comm_stream = None              # side stream of log_prob_async
        self.peer = None
        if peer_memory and self.world_size > 1 and dist.get_backend(group) == "nccl":
            try:
                self.peer = PeerComm(group, device=grevnet.params.device)
            except RuntimeError:
                if peer_memory is True:
                    raise

    def local_shard(self, host_batch: GraphsTuple) -> GraphsTuple:
        parts = partition_graphs(host_batch.n_node, host_batch.n_edge, self.world_size)
        return shard_graphs_tuple(host_batch, parts[self.rank])

    def broadcast_parameters(self, src: int = 0):
        """Replicate rank `src`'s flow parameters, batch-norm gamma/beta and moving statistics on every rank,
        and force the packed weight images to be rebuilt from them."""
        if self.world_size > 1:
            net = self.grevnet
            with torch.no_grad():
                for t in (net.params, net.bn_gamma, net.bn_beta, net.bn_moving_mean, net.bn_moving_var):
                    dist.broadcast(t.detach(), src=src, group=self.group)
            net._flow._packed_version = None

    def log_prob(self, local_graph: GraphsTuple, return_z: bool = False) -> dict:
        from .loss import mvn_log_prob_sum, scalars_from_vector
        z, ldj64 = self.grevnet.f64(local_graph)
        if self.peer is not None:             # log-prob assembly + all-reduce in ONE kernel over NVLink peer memory
            vec = self.peer.log_prob_all_reduce(z.nodes, ldj64)
        else:
            vec = all_reduce_log_prob(mvn_log_prob_sum(z.nodes, ldj64), self.group)
        out = scalars_from_vector(vec)
        if return_z:
            out["z"] = z
        return out

    def log_prob_async(self, local_graph: GraphsTuple) -> "PendingLogProb":
        """As log_prob, but the one collective of the path (32 bytes) is enqueued on a side stream, so the kernels
        of the NEXT batch do not wait for the slowest rank of this one (the overlap a DDP gradient bucket gets).
        Returns a handle; `.wait()` orders the current stream after the all-reduce and returns the scalars."""
        from .loss import mvn_log_prob_sum
        z, ldj64 = self.grevnet.f64(local_graph)
        vec = mvn_log_prob_sum(z.nodes, ldj64)
        dev = vec.device
        if self.world_size <= 1:
            return PendingLogProb(vec, z, None, None, None)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(dev)
        cs = self._comm_stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(cs):
            t0.record(cs)
            if self.peer is not None:         # one-warp kernel: co-resides with the next batch's fused CTAs
                self.peer.all_reduce_(vec)
            else:
                dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=self.group)
            t1.record(cs)
        vec.record_stream(cs)
        return PendingLogProb(vec, z, cs, t0, t1)

    def loss_and_grad(self, local_graph: GraphsTuple, per_node: bool = True):
        """Sharded training-step evaluation: local density pass, all-reduce of the 4-vector (gives the
        global node count the per-node loss is normalised by), local reversible backward, then the
        gradient all-reduce(SUM) -- every rank ends with the global-batch gradient.  With use_batch_norm the
        bijector's gamma/beta gradients (net.bn_gamma.grad / net.bn_beta.grad) are already global on every rank:
        they come from sums that _backward_bn all-reduced."""
        from .loss import mvn_log_prob_sum, scalars_from_vector
        net = self.grevnet
        z, ldj64 = net.f64(local_graph)
        if self.peer is not None:
            vec = self.peer.log_prob_all_reduce(z.nodes, ldj64)
        else:
            vec = all_reduce_log_prob(mvn_log_prob_sum(z.nodes, ldj64), self.group)
        out = scalars_from_vector(vec)
        n_global = max(float(vec[3].item()), 1.0)
        grads = net.backward_from_z(local_graph, z.nodes, 1.0 / n_global if per_node else 1.0)
        if self.world_size > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.group)      # 4 * params bytes (19.5 MB at T=6)
        net.params.grad = grads
        return out, grads

    def sample(self, local_latent: GraphsTuple) -> GraphsTuple:
        return self.grevnet.g(local_latent)
