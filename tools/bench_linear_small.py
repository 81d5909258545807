"""The stand-alone Linears of the f1 workload (N = 51 956 rows: q / k / v / output projections, layer 0, the backward's
W0^T) through k_gemm_tc (gnf_debug_linear_tc); run under `ncu --metrics gpu__time_duration.sum` and compare with the
k_linear_tc rows of profiles/r2_launches_f1_fwd_bwd.csv."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_normalizing_flows_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
m = 51956
for (k, n) in [(8, 80), (8, 12), (80, 80), (88, 256), (256, 88)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(k, n, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    c = torch.empty(m, n, device=dev)
    wsb = lib.gnf_debug_linear_tc_workspace(k, n)
    ws = _lib.workspace(wsb, dev)
    for _ in range(3):
        _lib.check(lib.gnf_debug_linear_tc(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), m, k, n, 2, _lib.MATH["tc3x"], _lib.ptr(c),
                                           _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), "gnf_debug_linear_tc")
    torch.cuda.synchronize()
    print(k, n, "max err", float((c.double() - (a.double() @ w.double() + b.double())).abs().max()))
