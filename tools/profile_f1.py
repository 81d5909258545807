"""f1 (dm_self_attn at the run_grevnet.py defaults) density pass + backward between cudaProfilerStart/Stop."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
rng = np.random.default_rng(0)
n_node = rng.integers(10, 41, size=2048)
s, r = G.utils.senders_receivers(n_node)
nodes = rng.standard_normal((int(n_node.sum()), 2)).astype(np.float32)
g = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
attn = dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False, kq_dim_division=False)
params = O.make_params(13, 12, 2, 256, 5, block="dm_attn", act="relu", attn=attn, last_layer_scale=0.05)
net = H.make_grevnet(params, 256, 5, device="cuda")
dg = H.to_device_graph(g)
n = nodes.shape[0]
for _ in range(2):
    z, _ = net.f64(dg)
    net.backward_from_z(dg, z.nodes, 1.0 / n)
torch.cuda.synchronize()
torch.cuda.profiler.start()
z, _ = net.f64(dg)
net.backward_from_z(dg, z.nodes, 1.0 / n)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled f1", n)
