"""Timeline of CTA 0 of one fused coupling kernel (gnf_debug_set_trace)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib
import bench

math = sys.argv[1] if len(sys.argv) > 1 else "tc3x"
host = bench.make_batch(4096, 12345)
net = H.make_grevnet(bench.make_oracle_params(), 256, 5, device="cuda", math=math)
g = host.to("cuda")
out = G.loss.log_prob(net, g)          # warm
torch.cuda.synchronize()
lib = _lib.load()
buf = torch.zeros(10 * 2048, dtype=torch.int64, device="cuda")
lib.gnf_debug_set_trace(_lib.ptr(buf))
handle = net._flow.ensure(net.params.detach())
st = G.graphs.structure_of(g)
n = g.nodes.shape[0]
m = _lib.MATH[math]
wsb = lib.gnf_grevnet_workspace(handle, n, m); ws = _lib.workspace(wsb, "cuda")
x0 = torch.zeros(n, 8, device="cuda"); x1 = torch.zeros(n, 8, device="cuda")
x0[:, :7] = g.nodes[:, :7]; x1[:, :7] = g.nodes[:, 7:]
# forward half step 0 only: use coupling_step inverse=0 => 2 kernels; trace holds the LAST kernel
_lib.check(lib.gnf_coupling_step(handle, 0, 0, _lib.ptr(x0), _lib.ptr(x1), n, st.n_edges, _lib.ptr(st.rowptr),
                                 _lib.ptr(st.csr_senders), None, m, _lib.ptr(ws), wsb, _lib.stream_ptr()))
torch.cuda.synchronize()
lib.gnf_debug_set_trace(None)
t = buf.cpu().numpy().astype(np.uint64).reshape(10, 2048)
ev = []
for role in range(10):
    cnt = int(t[role, 0])
    for v in t[role, 1:1 + cnt]:
        v = int(v)
        ev.append((v & 0xFFFFFFFFFF, role, v >> 56, (v >> 48) & 255, (v >> 40) & 255))
ev.sort()
t0 = ev[0][0]
names = {10: "MMA tile start (h_full)", 11: "MMA L0 issued", 12: "MMA chunk ready->issue", 13: "MMA last ready",
         14: "MMA last issued", 20: "EPI acc_full", 21: "EPI half converted", 22: "EPI last acc", 23: "EPI done",
         30: "GATHER start", 31: "GATHER computed", 32: "GATHER h_empty ok"}
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 260
prev = {}
for (c, role, e, ml, pk) in ev[:lim]:
    dt = c - prev.get(role, c); prev[role] = c
    print(f"{c - t0:9d} (+{dt:6d}) role{role} {names.get(e, e):26s} m={ml >> 4} l={ml & 15} ph={pk >> 2} kc={pk & 3}")
