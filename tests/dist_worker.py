"""Worker of the world_size-2 NCCL test (tests/test_gpu_round2.py): one process per GPU, the sharded wrapper
`GraphShardedGRevNet` on real NCCL -- log-prob all-reduce, gradient all-reduce, the batch-norm bijector's [2H+1] /
[2H] all-reduces, parameter broadcast -- checked by the parent against the unsharded single-GPU result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def make_case():
    import helpers as H
    from oracle import gnf_oracle as O
    rng = np.random.default_rng(77)
    g = H.random_batch(rng, 24, 5, 40, D=14)
    g = g._replace(nodes=(g.nodes * 1.3 + 0.2).astype(np.float32))
    params = O.make_params(9, 2, 14, 128, 4, last_layer_scale=0.1)
    gm = (1.0 + 0.2 * rng.standard_normal((2, 2, 7))).astype(np.float32).clip(0.5, 1.5)
    bt = (0.1 * rng.standard_normal((2, 2, 7))).astype(np.float32)
    return g, params, gm, bt


def build_net(params, device, use_bn, gm, bt, perturb=False):
    import torch
    import helpers as H
    net = H.make_grevnet(params, 128, 4, device=device, math="tc3x")
    if use_bn:
        net.use_batch_norm = True
        net.bn_gamma.data.copy_(torch.from_numpy(gm))
        net.bn_beta.data.copy_(torch.from_numpy(bt))
    if perturb:                               # non-source ranks start from different weights: broadcast must fix it
        with torch.no_grad():
            net.params.mul_(1.5)
            net.bn_gamma.mul_(0.7)
    return net


def run(rank, world, port, backend, q):
    import torch
    import torch.distributed as dist
    import graph_normalizing_flows_b200 as G
    from graph_normalizing_flows_b200 import sharding as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group(backend, rank=rank, world_size=world, device_id=dev)
    try:
        g, params, gm, bt = make_case()
        host = G.GraphsTuple(*g)
        res = {}
        for use_bn in (False, True):
            net = build_net(params, dev, use_bn, gm, bt, perturb=(rank != 0))
            net.bn_update_moving = False
            _ = G.loss.log_prob(net, S.shard_graphs_tuple(host, np.arange(2)).to(dev))   # packs the PERTURBED weights first
            net.bn_update_moving = True
            sharded = S.GraphShardedGRevNet(net)                       # peer-memory all-reduce when the GPUs allow it
            sharded.broadcast_parameters(0)
            local = sharded.local_shard(host).to(dev)
            nccl_only = S.GraphShardedGRevNet(net, peer_memory=False)
            net.bn_update_moving = False
            out_nccl = nccl_only.log_prob(local)
            net.bn_update_moving = True
            out = sharded.log_prob(local)
            pend = sharded.log_prob_async(local)
            out_async = pend.wait()
            scal, grads = sharded.loss_and_grad(local, per_node=True)
            torch.cuda.synchronize()
            key = "bn" if use_bn else "plain"
            res[key] = {
                "vec": [float(out[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")],
                "vec_nccl": [float(out_nccl[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")],
                "peer_used": sharded.peer is not None,
                "vec_async": [float(out_async[k]) for k in ("log_prob_zs", "log_det_jacobian", "log_prob_xs", "num_nodes")],
                "loss_per_node": float(scal["loss_per_node"]),
                "grads": grads.double().cpu().numpy(),
                "n_local": int(local.nodes.shape[0]),
            }
            if use_bn:
                res[key]["g_gamma"] = net.bn_gamma.grad.double().cpu().numpy()
                res[key]["g_beta"] = net.bn_beta.grad.double().cpu().numpy()
                res[key]["moving_mean"] = net.bn_moving_mean.double().cpu().numpy()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()
