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


class GraphedLogProb:
    """The density pass + log-prob assembly of ONE batch structure captured in a CUDA graph: the
    2T fused coupling kernels, the reductions and the log-prob kernels replay with a single launch
    (small batches are launch-bound: ~30 kernels of a few microseconds each).

        runner = GraphedLogProb(grevnet, graph)        # warm-up + capture on a side stream
        vec = runner(new_nodes)                         # float64 [4]: log_prob_zs, ldj, log_prob_xs, N
        z = runner.z                                    # [N, D] latent of the last replay

    Node features may change between replays (they are copied into the captured input buffer);
    the structure (senders / receivers) and the parameters' packed image are the captured ones --
    re-create the runner after an optimiser step or for a different batch structure."""

    def __init__(self, grevnet, graph: GraphsTuple, warmup: int = 2):
        from .graphs import structure_of
        structure_of(graph)                                        # validation + CSR outside the capture
        grevnet._flow.ensure(grevnet.params.detach())              # pack weights outside the capture
        _ = grevnet.math
        self.grevnet = grevnet
        self.graph = graph.replace(nodes=graph.nodes.clone())
        dev = graph.nodes.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                z, ldj = grevnet.f64(self.graph)
                mvn_log_prob_sum(z.nodes, ldj)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.cuda_graph):
            z, ldj = grevnet.f64(self.graph)
            self.vec = mvn_log_prob_sum(z.nodes, ldj)
        self.z = z.nodes

    def __call__(self, nodes: torch.Tensor = None) -> torch.Tensor:
        if nodes is not None:
            self.graph.nodes.copy_(nodes, non_blocking=True)
        self.cuda_graph.replay()
        return self.vec


# ------------------------------------------------------------------------------------------------
# f3  decode tail of the sampling pass (train_grevnet_with_data.py:414-416, 529-548)
# ------------------------------------------------------------------------------------------------
def pred_adj(gnn_output: GraphsTuple, temp: float = 10.0, shift: float = 1.0):
    """pred_adj(graph, scaled_hacky_sigmoid_l2) (loss.py:154-159,45-53): per graph the [n, n] block
    sigmoid(temp * (shift - |x_i - x_j|^2 / sqrt(d))) with a zero diagonal.  Returns (blocks, adj_off):
    a flat float32 device tensor holding the blocks back to back and the int64 offsets [G+1].  The
    reference materialises the dense [N, N] matrix times a block mask; `dense_pred_adj` rebuilds that
    view for small batches."""
    lib = _lib.load()
    nodes = _lib.require_cuda(gnn_output.nodes, "graph.nodes", torch.float32)
    n_node = torch.as_tensor(gnn_output.n_node).to(nodes.device, torch.int64)
    zero = torch.zeros(1, dtype=torch.int64, device=nodes.device)
    node_off = torch.cat([zero, torch.cumsum(n_node, 0)]).contiguous()
    adj_off = torch.cat([zero, torch.cumsum(n_node * n_node, 0)]).contiguous()
    out = torch.empty(int(adj_off[-1].item()), dtype=torch.float32, device=nodes.device)
    _lib.check(lib.gnf_pred_adj(_lib.ptr(nodes), nodes.shape[1], _lib.ptr(node_off), _lib.ptr(adj_off),
                                int(n_node.numel()), float(temp), float(shift), _lib.ptr(out),
                                _lib.stream_ptr(nodes.device)), "gnf_pred_adj")
    return out, adj_off


def dense_pred_adj(blocks: torch.Tensor, adj_off: torch.Tensor, n_node) -> torch.Tensor:
    """The reference's dense [N, N] layout (zeros off the block diagonal) from the packed blocks."""
    n_node = [int(v) for v in torch.as_tensor(n_node).tolist()]
    off = [int(v) for v in adj_off.tolist()]
    return torch.block_diag(*[blocks[off[g]:off[g + 1]].view(n, n) for g, n in enumerate(n_node)])


def sampled_graphs(blocks: torch.Tensor, adj_off: torch.Tensor, n_node, threshold: float = 0.5):
    """train_grevnet_with_data.py:529-548: adjacency = pred_adj > 0.5 per sampled graph -> networkx."""
    import networkx as nx
    n_node = [int(v) for v in torch.as_tensor(n_node).tolist()]
    off = [int(v) for v in adj_off.tolist()]
    host = blocks.cpu().numpy()
    return [nx.from_numpy_array((host[off[g]:off[g + 1]].reshape(n, n) > threshold).astype(float))
            for g, n in enumerate(n_node)]
