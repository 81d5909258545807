#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (small shapes): k_gemm_tc (ragged K / N, several column blocks,
# rows below one tile), the layered flows that call it, k_dm_attn_warp, k_dw_skinny, the chained batch-norm entries
set -x
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
os.environ["GNF_POISON_WORKSPACE"] = "1"
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
for (m, k, n) in [(300, 164, 2048), (70, 2048, 100), (129, 16, 16), (1, 40, 300), (600, 100, 104)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(k, n, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    c = torch.empty(m, n, device=dev)
    wsb = lib.gnf_debug_linear_tc_workspace(k, n); ws = _lib.workspace(wsb, dev)
    for math in ("tc3x", "bf16"):
        _lib.check(lib.gnf_debug_linear_tc(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), m, k, n, 0, _lib.MATH[math], _lib.ptr(c),
                                           _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), "gnf_debug_linear_tc")
    torch.cuda.synchronize()
    print("gemm", m, k, n, float((c.double() - torch.nn.functional.leaky_relu(a.double() @ w.double() + b.double(), 0.2)).abs().max()))
rng = np.random.default_rng(0)
# layered flow: message passing with L = 64, and attention with residual + wide keys / values (k_dm_attn_warp)
g = H.random_batch(rng, 10, 5, 30, D=6)
net = H.make_grevnet(O.make_params(2, 2, 6, 64, 3, last_layer_scale=0.05), 64, 3)
dg = H.to_device_graph(g)
out = G.loss.log_prob(net, dg, return_z=True)
x = net(out["z"], inverse=False)
print("layered mp", net.math, float(out["log_prob_xs"]), float((x.nodes - dg.nodes).abs().max()))
n_node = np.array([40, 3, 17], np.int32)
s, r = G.utils.senders_receivers(n_node)
nodes = rng.standard_normal((int(n_node.sum()), 8)).astype(np.float32)
gfc = O.GraphsTuple(nodes, None, np.asarray(r, np.int32), np.asarray(s, np.int32), None, n_node, (n_node.astype(np.int64) ** 2).astype(np.int32))
attn = dict(num_heads=2, kq_dim=33, v_dim=40, out_dim=24, concat=True, residual=True, kq_dim_division=True)
pa = O.make_params(5, 1, 8, 96, 3, block="dm_attn", act="leaky_relu", attn=attn, last_layer_scale=0.1)
neta = H.make_grevnet(pa, 96, 3)
dfc = H.to_device_graph(gfc)
out = G.loss.log_prob(neta, dfc, return_z=True)
x = neta(out["z"], inverse=False)
_, ga = neta.loss_and_grad(dfc)
print("layered attn", neta.math, float(out["log_prob_xs"]), float((x.nodes - dfc.nodes).abs().max()), float(ga.norm()))
# k_dw_skinny: attention projections with D/2 = 1 (run_grevnet.py defaults), backward on the tensor cores and fp32
n_node = rng.integers(4, 30, size=8)
s, r = G.utils.senders_receivers(n_node)
nodes = rng.standard_normal((int(n_node.sum()), 2)).astype(np.float32)
g2 = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
p2 = O.make_params(13, 1, 2, 256, 3, block="dm_attn", act="relu", last_layer_scale=0.1,
                   attn=dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False, kq_dim_division=False))
net2 = H.make_grevnet(p2, 256, 3)
for bm in ("tc3x", "fp32"):
    _, g2g = net2.loss_and_grad(H.to_device_graph(g2), backward_math=bm)
    print("skinny dW", bm, float(g2g.norm()))
# chained batch-norm entries, both directions
gb = H.random_batch(rng, 12, 5, 30, D=14)
netb = H.make_grevnet(O.make_params(1, 2, 14, 128, 3, last_layer_scale=0.05), 128, 3)
netb.use_batch_norm = True
db = H.to_device_graph(gb)
out = G.loss.log_prob(netb, db, return_z=True)
x = netb(out["z"], inverse=False)
print("bn chain", float(out["log_prob_xs"]), float(x.nodes.abs().max()))
torch.cuda.synchronize()
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san2.py > gpurun_out/sanitizer_new_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_new_$tool.log
done
