"""CPU suite: graph sharding + the single all-reduce of the path, world_size 2 over gloo.
The local compute is the oracle (tests only); what is under test is the partitioner, the index
re-basing and the reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import gnf_oracle as O
from graph_normalizing_flows_b200 import sharding as S
from graph_normalizing_flows_b200.graphs import GraphsTuple


def _batch():
    g = H.random_batch(np.random.default_rng(11), 9, 3, 25, D=14)
    return GraphsTuple(*g)


def test_partition_is_balanced_and_complete():
    fam = np.load(os.path.join(H.GOLDEN, "graphs_protein_4_128.npz"))
    n_node, n_edge = fam["n_node"][:256], fam["n_edge"][:256]
    for world in (1, 2, 4, 8):
        parts = S.partition_graphs(n_node, n_edge, world)
        allg = np.sort(np.concatenate(parts))
        assert np.array_equal(allg, np.arange(256))
        cost = S.graph_costs(n_node, n_edge)
        loads = np.array([cost[p].sum() for p in parts])
        # LPT bound: max load <= mean + largest item
        assert loads.max() <= loads.mean() + cost.max() + 1e-6
        if world > 1:
            naive = np.array([cost[c].sum() for c in np.array_split(np.arange(256), world)])
            assert loads.max() <= naive.max() + 1e-6
    assert S.partition_graphs(n_node, n_edge, 4)[0].tolist() == S.partition_graphs(n_node, n_edge, 4)[0].tolist()


def test_shard_rebases_indices_exactly():
    g = _batch()
    parts = S.partition_graphs(g.n_node, g.n_edge, 3)
    node_off = np.concatenate([[0], np.cumsum(g.n_node)])
    edge_off = np.concatenate([[0], np.cumsum(g.n_edge)])
    for ids in parts:
        sh = S.shard_graphs_tuple(g, ids)
        assert sh.senders.dtype == np.int32 and int(sh.n_node.sum()) == sh.nodes.shape[0]
        off = 0
        eo = 0
        for k in ids:
            ne = int(g.n_edge[k])
            assert np.array_equal(sh.senders[eo:eo + ne] - off, g.senders[edge_off[k]:edge_off[k + 1]] - node_off[k])
            assert np.array_equal(sh.receivers[eo:eo + ne] - off, g.receivers[edge_off[k]:edge_off[k + 1]] - node_off[k])
            assert np.array_equal(sh.nodes[off:off + g.n_node[k]], g.nodes[node_off[k]:node_off[k + 1]])
            off += int(g.n_node[k])
            eo += ne
    empty = S.shard_graphs_tuple(g, np.zeros(0, np.int64))
    assert empty.nodes.shape == (0, 14) and len(empty.senders) == 0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = _batch()
        params = O.make_params(5, 2, 14, 32, 3, last_layer_scale=0.1)
        ids = S.partition_graphs(g.n_node, g.n_edge, world)[rank]
        sh = S.shard_graphs_tuple(g, ids)
        z, ldj = O.grevnet_f(sh.nodes.astype(np.float64), sh.senders, sh.receivers,
                             O.cast_params(params, np.float64))
        lp = O.log_prob(z, ldj, sh.n_node)
        vec = torch.tensor([float(lp["log_prob_zs"]), float(ldj), float(lp["log_prob_xs"]), float(sh.n_node.sum())],
                           dtype=torch.float64)
        S.all_reduce_log_prob(vec)
        q.put((rank, vec.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_all_reduce_matches_unsharded():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = _batch()
    params = O.cast_params(O.make_params(5, 2, 14, 32, 3, last_layer_scale=0.1), np.float64)
    z, ldj = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, params)
    lp = O.log_prob(z, ldj, g.n_node)
    want = [float(lp["log_prob_zs"]), float(ldj), float(lp["log_prob_xs"]), float(g.n_node.sum())]
    assert res[0] == res[1]                                   # every rank holds the global scalars
    for a, b in zip(res[0], want):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b))


def test_all_reduce_rejects_wrong_vector():
    with pytest.raises(ValueError):
        S.all_reduce_log_prob(torch.zeros(4, dtype=torch.float32))
