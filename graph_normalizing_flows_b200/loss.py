"""Log-prob assembly of the GRevNet scripts.

The reference's loss.py holds only the auto-encoder losses; the flow's log-prob lives inline in
the two scripts (run_grevnet.py:290-302, train_grevnet_with_data.py:346-355) plus a dead method
GRevNet.log_prob (gnn.py:375-377).  BASELINE's north star places `log_prob` here, so this
module supplies it, following those script lines exactly:

    z, log_det_jacobian = grevnet(graph, inverse=True)                         run_grevnet.py:291
    log_prob_zs = sum(MultivariateNormalDiag(0, I).log_prob(z.nodes))          run_grevnet.py:292-294
    log_prob_xs = log_prob_zs + log_det_jacobian                               run_grevnet.py:295
    total_loss  = -log_prob_xs                                                 run_grevnet.py:296
    *_per_node  = * / sum(n_node)                                              run_grevnet.py:298-302
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .graphs import GraphsTuple

LOG_2PI = math.log(2.0 * math.pi)


def mvn_log_prob_sum(z_nodes: torch.Tensor, ldj64: torch.Tensor = None) -> torch.Tensor:
    """Device float64 [4]: (log_prob_zs, log_det_jacobian, log_prob_xs, num_nodes) via gnf_log_prob
    (fp64 fixed-order reduction, so 1/2/4/8-GPU shards agree)."""
    lib = _lib.load()
    _lib.require_cuda(z_nodes, "z.nodes", torch.float32)
    n, d = z_nodes.shape
    out = torch.empty(4, dtype=torch.float64, device=z_nodes.device)
    wsb = lib.gnf_log_prob_workspace(n, d)
    ws = _lib.workspace(wsb, z_nodes.device)
    if ldj64 is not None:
        _lib.require_cuda(ldj64, "log_det_jacobian", torch.float64)
    _lib.check(lib.gnf_log_prob(_lib.ptr(z_nodes), n, d, _lib.ptr(ldj64), _lib.ptr(out), _lib.ptr(ws), wsb,
                                _lib.stream_ptr(z_nodes.device)), "gnf_log_prob")
    return out


def scalars_from_vector(v: torch.Tensor) -> dict:
    """The scalars the scripts log (run_grevnet.py:385-392) from the 4-vector above."""
    log_prob_zs, ldj, log_prob_xs, num_nodes = v[0], v[1], v[2], v[3]
    total_loss = -log_prob_xs
    return {
        "log_prob_zs": log_prob_zs,
        "log_det_jacobian": ldj,
        "log_prob_xs": log_prob_xs,
        "total_loss": total_loss,
        "num_nodes": num_nodes,
        "loss_per_node": total_loss / num_nodes,
        "log_prob_xs_per_node": log_prob_xs / num_nodes,
        "log_prob_zs_per_node": log_prob_zs / num_nodes,
        "log_det_jacobian_per_node": ldj / num_nodes,
    }


def log_prob(grevnet, graph: GraphsTuple, return_z: bool = False):
    """Density pass + log-prob assembly.  Returns a dict of 0-dim device tensors (float64)."""
    z, ldj64 = grevnet.f64(graph)
    out = scalars_from_vector(mvn_log_prob_sum(z.nodes, ldj64))
    if return_z:
        out["z"] = z
    return out
