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
# round 2 kernels: whole flow in one cooperative launch, inject mode + tensor-core linears + block-staged attention,
# staged gather+segment kernel (fast path and fallback), one-launch re-pack (runs inside make_grevnet)
for persist in ("0", "1"):
    os.environ["GNF_PERSIST"] = persist
    net = H.make_grevnet(params, 256, 4, math="tc3x")
    out = G.loss.log_prob(net, dg, return_z=True)
    x = net(out["z"], inverse=False)
    print("persist", persist, float(out["log_prob_xs"]), float((x.nodes - dg.nodes).abs().max()))
os.environ.pop("GNF_PERSIST")
n_node = rng.integers(4, 30, size=12)
s, r = G.utils.senders_receivers(n_node)
nodes = rng.standard_normal((int(n_node.sum()), 2)).astype(np.float32)
gfc = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
pa = O.make_params(13, 2, 2, 256, 4, block="dm_attn", act="relu", last_layer_scale=0.1,
                   attn=dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False, kq_dim_division=False))
neta = H.make_grevnet(pa, 256, 4)
dfc = H.to_device_graph(gfc)
for math in ("tc3x", "tc3x_bf16", "fp32"):
    neta.math = math
    out = G.loss.log_prob(neta, dfc, return_z=True)
    x = neta(out["z"], inverse=False)
    print("f1", math, float(out["log_prob_xs"]), float((x.nodes - dfc.nodes).abs().max()))
g24 = H.random_batch(rng, 10, 5, 40, D=24)
net24 = H.make_grevnet(O.make_params(2, 2, 24, 128, 3, last_layer_scale=0.05), 128, 3)
print("inject mp", net24.math, float(G.loss.log_prob(net24, H.to_device_graph(g24))["log_prob_xs"]))
# backward of the inject flows: k_linear_tc (K = 256 in two passes), k_bwd_chain<BINJ>, k_dw_tc with a 32/96-column
# input image, block-staged attention backward (both passes) + the thread-per-head kernels behind it
for bm in ("tc3x", "bf16", "fp32"):
    _, ga = neta.loss_and_grad(dfc, backward_math=bm)
    _, g24g = net24.loss_and_grad(H.to_device_graph(g24), backward_math=bm)
    print("inject backward", bm, float(ga.norm()), float(g24g.norm()))
st = G.graphs.structure_of(dg)
for h in (7, 8, 20):
    xs = torch.randn(dg.nodes.shape[0], h, device="cuda")
    print("seg", h, float(G.gnn.gather_segment_reduce(xs, st, "mean").sum()))
ed = torch.randn(st.n_edges, 7, device="cuda")
print("segsum edges", float(G.gnn.unsorted_segment_sum(ed, dg.receivers, dg.nodes.shape[0]).sum()))
torch.cuda.synchronize()
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -6 gpurun_out/sanitizer_$tool.log
done
