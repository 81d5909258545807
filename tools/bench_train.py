"""Training-step timing (row f2): tc3x forward + tensor-core reversible backward + Adam at the bench workload."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import graph_normalizing_flows_b200 as G
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
host = bench.make_batch(B, 12345)
net = H.make_grevnet(bench.make_oracle_params(), 256, 5, device="cuda", math="tc3x")
g = host.to("cuda")
opt = torch.optim.Adam([net.params], lr=1e-4, betas=(0.9, 0.9))
for _ in range(2):
    out, _ = net.loss_and_grad(g, per_node=True); opt.step()
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 5
for _ in range(n):
    out, grads = net.loss_and_grad(g, per_node=True); opt.step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
N = g.nodes.shape[0]
print(json.dumps({"graphs": B, "nodes": N, "train_step_ms": dt * 1e3, "node_updates_per_s_train": N * 12 / dt,
                  "loss_per_node": float(out["loss_per_node"]), "grad_norm": float(grads.norm())}))
