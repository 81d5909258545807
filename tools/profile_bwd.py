"""One tensor-core backward (12 reversed half steps) at the bench workload between cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_bwd.csv python tools/profile_bwd.py [graphs] [math]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
math = sys.argv[2] if len(sys.argv) > 2 else "tc3x"
host = bench.make_batch(B, 12345)
net = H.make_grevnet(bench.make_oracle_params(), 256, 5, device="cuda", math="tc3x")
g = host.to("cuda")
z, _ = net.f64(g)
n = g.nodes.shape[0]
grads = torch.zeros_like(net.params.detach())
for _ in range(2):
    net.backward_from_z(g, z.nodes, 1.0 / n, grads=grads, math=math)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.backward_from_z(g, z.nodes, 1.0 / n, grads=grads, math=math)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", B, math, n)
