"""B200-native GRevNet hot path with the call surface of jliu/graph-normalizing-flows.

    from graph_normalizing_flows_b200 import gnn, loss, graphs
    grevnet = gnn.GRevNet(make_gnn_fn, num_timesteps, node_embedding_dim)
    z, log_det_jacobian = grevnet(graph, inverse=True)     # density direction (gnn.py:379-381)
    x = grevnet(graph, inverse=False)                      # sampling direction

All arithmetic runs in libgnf_b200.so (include/gnf_b200.h).  No CPU fallback.
"""
from . import _lib, gnn, graphs, loss, utils  # noqa: F401
from .gnn import (GRevNet, NodeBlockGNN, ConcatThenMLPBlock, AggThenMLPBlock, make_mlp_model,  # noqa: F401
                  sum_concat_then_mlp_gnn, avg_concat_then_mlp_gnn, sum_then_mlp_gnn, avg_then_mlp_gnn,
                  dm_self_attn_gnn, DMSelfAttentionMLP)
from .graphs import GraphsTuple  # noqa: F401

__version__ = "0.1.0"
