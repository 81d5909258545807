"""Packed graph batches.

`GraphsTuple` mirrors graph_nets.graphs.GraphsTuple as the reference uses it
(train_grevnet_with_data.py:265-271; `.replace(nodes=...)` at gnn.py:307-308): an immutable
namedtuple `(nodes, edges, receivers, senders, globals, n_node, n_edge)` holding G graphs
concatenated, with senders/receivers already offset into the packed node array.

Layout contract at the kernel boundary:  nodes [N, D] float32 row-major; senders, receivers
[E] int32; n_node, n_edge [G] int32.  `edges` / `globals` are carried but never read by the
hot path (the reference fills them with zeros, graph_data.py:33-50).
"""
from __future__ import annotations

import collections
import weakref
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib

_FIELDS = ("nodes", "edges", "receivers", "senders", "globals", "n_node", "n_edge")


class GraphsTuple(collections.namedtuple("GraphsTuple", _FIELDS)):
    __slots__ = ()

    def replace(self, **kwargs):
        return self._replace(**kwargs)

    def map(self, fn, fields=_FIELDS):
        return self._replace(**{k: fn(getattr(self, k)) for k in fields if getattr(self, k) is not None})

    def to(self, device, non_blocking=True):
        """Host (numpy / CPU tensor) -> device tensors; the feed_dict H2D of the reference
        (run_grevnet.py:447)."""
        def mv(v):
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(v)
            return v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        return self.map(mv)


def data_dicts_to_graphs_tuple(data_dicts: Sequence[dict]) -> GraphsTuple:
    """graph_nets.utils_np.data_dicts_to_graphs_tuple as used at grevnet_synthetic_data.py:45-47:
    concatenate per-graph arrays, offset senders/receivers by the cumulative node count."""
    n_node = np.array([int(d["n_node"]) for d in data_dicts], dtype=np.int32)
    n_edge = np.array([len(d["senders"]) for d in data_dicts], dtype=np.int32)
    offsets = np.concatenate([[0], np.cumsum(n_node)[:-1]]).astype(np.int64) if len(n_node) else np.zeros(0, np.int64)
    rep = np.repeat(offsets, n_edge)
    cat = lambda key, dt: (np.concatenate([np.asarray(d[key], dtype=dt).reshape(-1) for d in data_dicts])
                           if data_dicts else np.zeros(0, dt))
    senders = (cat("senders", np.int64) + rep).astype(np.int32)
    receivers = (cat("receivers", np.int64) + rep).astype(np.int32)
    nodes = None
    if data_dicts and data_dicts[0].get("nodes") is not None:
        nodes = np.concatenate([np.asarray(d["nodes"], dtype=np.float32) for d in data_dicts], axis=0)
    return GraphsTuple(nodes=nodes, edges=None, receivers=receivers, senders=senders, globals=None,
                       n_node=n_node, n_edge=n_edge)


def concat_structures(structs: Iterable[tuple], nodes=None) -> GraphsTuple:
    """structs: iterable of (n_node, senders_local, receivers_local) per graph."""
    return data_dicts_to_graphs_tuple(
        [{"n_node": n, "senders": s, "receivers": r} for (n, s, r) in structs]).replace(nodes=nodes)


def networkx_to_arrays(g):
    """graph_nets.utils_np.networkx_to_data_dict [upstream], index part: nodes keyed 0..n-1 in
    iteration order; edge order = networkx iteration order (sender-major)."""
    n = g.number_of_nodes()
    if n and (list(g.nodes()) != list(range(n))):
        raise ValueError("graph nodes must be sequentially numbered 0..n-1 in iteration order")
    ed = np.fromiter((x for e in g.edges() for x in e), dtype=np.int32, count=2 * g.number_of_edges())
    ed = ed.reshape(-1, 2)
    return n, np.ascontiguousarray(ed[:, 0]), np.ascontiguousarray(ed[:, 1])


def networkxs_to_graphs_tuple(graphs, nodes=None) -> GraphsTuple:
    """gn.utils_np.networkxs_to_graphs_tuple (graph_data.py:107,110,122), index part."""
    return concat_structures([networkx_to_arrays(g) for g in graphs], nodes=nodes)


class BatchStructure:
    """Device-side, per-batch index structure: stable CSR by receiver.  Built once per batch
    (cached on the senders/receivers tensors) and shared by all 2T half steps."""

    __slots__ = ("n_nodes", "n_edges", "rowptr", "perm", "csr_senders", "device", "_bad")

    def __init__(self, senders: torch.Tensor, receivers: torch.Tensor, n_nodes: int, validate=True):
        """validate: True = count out-of-range ids and raise now (one host sync); "deferred" = launch
        the check but read it later with .check() (prefetch path); False = trust the caller."""
        lib = _lib.load()
        _lib.require_cuda(senders, "senders", torch.int32)
        _lib.require_cuda(receivers, "receivers", torch.int32)
        if senders.shape != receivers.shape or senders.dim() != 1:
            raise ValueError("senders and receivers must be 1-D and of equal length")
        dev = senders.device
        n, e = int(n_nodes), int(senders.numel())
        st = _lib.stream_ptr(dev)
        self.n_nodes, self.n_edges, self.device = n, e, dev
        self._bad = None
        if validate:
            # reference behaviour: TF raises InvalidArgumentError on out-of-range ids (CPU)
            self._bad = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(lib.gnf_validate_indices(_lib.ptr(senders), _lib.ptr(receivers), n, e, _lib.ptr(self._bad), st),
                       "gnf_validate_indices")
            if validate != "deferred":
                self.check()
        self.rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.perm = torch.empty(e, dtype=torch.int32, device=dev)
        self.csr_senders = torch.empty(e, dtype=torch.int32, device=dev)
        wsb = lib.gnf_build_csr_workspace(n, e)
        ws = _lib.workspace(wsb, dev)
        _lib.check(lib.gnf_build_csr(_lib.ptr(receivers), _lib.ptr(senders), n, e, _lib.ptr(self.rowptr),
                                     _lib.ptr(self.perm), _lib.ptr(self.csr_senders), _lib.ptr(ws), wsb, st),
                   "gnf_build_csr")


    def check(self):
        """Raise if the (possibly deferred) index validation found ids outside [0, N)."""
        if self._bad is not None:
            nbad = int(self._bad.item())
            self._bad = None
            if nbad:
                raise ValueError(f"{nbad} sender/receiver indices outside [0, {self.n_nodes})")


class BatchPrefetcher:
    """Stages the NEXT batch while the current one computes: pinned-host -> device copies, index
    validation and the CSR build run on a side stream; `wait` hands the batch to the current stream.

        pf = BatchPrefetcher(device)
        ticket = pf.submit(host_batch)                 # host_batch: GraphsTuple of (pinned) CPU tensors / numpy
        ...                                            # compute on the previous batch
        graph = pf.wait(ticket)                        # device GraphsTuple, structure cached, ids validated
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def submit(self, host: GraphsTuple):
        n = int(host.nodes.shape[0])
        for name, tot in (("n_node", n), ("n_edge", int(len(host.senders)))):
            v = getattr(host, name)
            if v is not None and int(np.asarray(v).sum()) != tot:       # host side of the error contract
                raise ValueError(f"sum({name})={int(np.asarray(v).sum())} does not match {tot}")
        with torch.cuda.stream(self.stream):
            g = host.to(self.device, non_blocking=True)
            st = BatchStructure(g.senders, g.receivers, n, validate="deferred")
            s, r = g.senders, g.receivers
            s._gnf_structure = ((id(r), r._version, s._version, n), weakref.ref(r), st)
            bad_host = torch.empty(1, dtype=torch.int32).pin_memory()
            bad_host.copy_(st._bad, non_blocking=True)       # lands with the staging work, on the side stream
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return g, st, ev, bad_host

    def wait(self, ticket) -> GraphsTuple:
        g, st, ev, bad_host = ticket
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in list(g) + [st.rowptr, st.perm, st.csr_senders]:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(cur)
        # the host waits for the STAGING stream's event only (copies + CSR build of a batch submitted a step ago),
        # never for the compute stream: no pipeline stall, and the ids are checked before any kernel consumes them
        ev.synchronize()
        st._bad = None
        nbad = int(bad_host[0])
        if nbad:
            raise ValueError(f"{nbad} sender/receiver indices outside [0, {st.n_nodes})")
        return g


def structure_of(graph: GraphsTuple) -> BatchStructure:
    """Cached BatchStructure of a device-resident GraphsTuple.  The cache lives on the senders
    tensor object itself (it travels through `.replace(nodes=...)`), keyed on the identity and
    in-place version of both index tensors.  Checks sum(n_node) == N and sum(n_edge) == E once
    (host side of the error contract)."""
    s, r = graph.senders, graph.receivers
    if not isinstance(s, torch.Tensor) or not isinstance(r, torch.Tensor):
        raise TypeError("graph.senders / graph.receivers must be torch tensors on a CUDA device "
                        "(use GraphsTuple.to(device)); there is no CPU fallback")
    n = int(graph.nodes.shape[0])
    key = (id(r), r._version, s._version, n)
    hit = getattr(s, "_gnf_structure", None)
    if hit is not None and hit[0] == key and hit[1]() is r:
        return hit[2]
    # validation + CSR build are enqueued first; the bad-id count and the two size sums come back in ONE host read
    st = BatchStructure(s, r, n, validate="deferred")
    dev_vals, names, host_checks = [st._bad.to(torch.int64).reshape(1)], ["bad"], []
    for name, v, want in (("n_node", graph.n_node, n), ("n_edge", graph.n_edge, int(s.numel()))):
        if v is None:
            continue
        v = torch.as_tensor(v)
        if v.is_cuda:
            dev_vals.append(v.sum(dtype=torch.int64).reshape(1))
            names.append((name, want))
        else:
            host_checks.append((name, int(v.sum()), want))
    got = torch.cat(dev_vals).tolist()
    st._bad = None
    for name, tot, want in host_checks + [(nm[0], got[i], nm[1]) for i, nm in enumerate(names) if i > 0]:
        if tot != want:
            what = "nodes.shape[0]" if name == "n_node" else "len(senders)"
            raise ValueError(f"sum({name})={tot} does not match {what}={want}")
    if got[0]:
        raise ValueError(f"{got[0]} sender/receiver indices outside [0, {n})")
    s._gnf_structure = (key, weakref.ref(r), st)
    return st


def transposed_structure_of(graph: GraphsTuple) -> BatchStructure:
    """CSR by SENDER (rowptr over out-edges, payload = receivers): the transpose the backward of
    gather + segment-reduce walks.  Cached on the receivers tensor."""
    s, r = graph.senders, graph.receivers
    n = int(graph.nodes.shape[0])
    key = (id(s), s._version, r._version, n)
    hit = getattr(r, "_gnf_structure_t", None)
    if hit is not None and hit[0] == key and hit[1]() is s:
        return hit[2]
    st = BatchStructure(r, s, n, validate=False)      # roles swapped: keyed by sender, carries receivers
    r._gnf_structure_t = (key, weakref.ref(s), st)
    return st
