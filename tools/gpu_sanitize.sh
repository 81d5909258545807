#!/bin/bash
# compute-sanitizer passes over the hot path (small shapes so the tools finish in a minute or two)
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
rng = np.random.default_rng(0)
g = H.random_batch(rng, 20, 5, 40, D=14, isolated=True)
params = O.make_params(1, 2, 14, 256, 4, last_layer_scale=0.05)
dg = H.to_device_graph(g)
for math in ("tc3x", "bf16", "fp32"):
    net = H.make_grevnet(params, 256, 4, math=math)
    out = G.loss.log_prob(net, dg, return_z=True)
    x = net(out["z"], inverse=False)
    print(math, float(out["log_prob_xs"]), float((x.nodes - dg.nodes).abs().max()))
net = H.make_grevnet(params, 256, 4)
for bm in ("tc3x", "bf16", "fp32"):            # tensor-core backward (k_bwd_chain + k_dw_tc), hi-only dW, FFMA path
    out, grads = net.loss_and_grad(dg, backward_math=bm)
    print("backward", bm, float(grads.norm()))
blocks, off = G.loss.pred_adj(dg)
torch.cuda.synchronize()
print("grad norm", float(grads.norm()), "adj", float(blocks.sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -6 gpurun_out/sanitizer_$tool.log
done
