"""A/B of the two launch schemes of the fused flow (2T launches vs one cooperative launch) inside one process:
per-step CUDA-event times, L2 flushed between steps.  usage: ab_persist.py <workload> [steps] [timing 0|1]"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import bench
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "community_medium"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
timing = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = bench.WORKLOADS[name]
gpg = wl.get("graphs_per_gpu") or wl.get("total_graphs")
host = bench.make_global_batch(wl, 1, gpg)
g = host.to("cuda")
net = H.make_grevnet(bench.make_oracle_params(wl), bench.L, bench.K, device="cuda", math=wl["math"])
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
G.graphs.structure_of(g)
for mode in ("0", "1", "0", "1"):
    os.environ["GNF_PERSIST"] = mode
    for _ in range(5):
        G.loss.mvn_log_prob_sum(*(lambda zl: (zl[0].nodes, zl[1]))(net.f64(g)))
    torch.cuda.synchronize()
    if timing:
        lib.gnf_debug_kernel_timing(1)
    a = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 255)
        a[i].record()
        G.loss.mvn_log_prob_sum(*(lambda zl: (zl[0].nodes, zl[1]))(net.f64(g)))
        b[i].record()
    torch.cuda.synchronize()
    kt, kc = ctypes.c_double(0.0), ctypes.c_int64(0)
    if timing:
        lib.gnf_debug_kernel_time(ctypes.byref(kt), ctypes.byref(kc))
        lib.gnf_debug_kernel_timing(0)
    ms = sorted(x.elapsed_time(y) for x, y in zip(a, b))
    print(json.dumps({"workload": name, "persist": mode, "timing": timing, "steps": steps, "n_nodes": int(g.nodes.shape[0]),
                      "ms_min": ms[0], "ms_median": ms[len(ms) // 2], "ms_mean": sum(ms) / len(ms), "ms_max": ms[-1],
                      "kernel_ms_sum_per_step": kt.value / steps if timing else None, "kernel_launches": kc.value}))
