"""k_gemm_tc on its own (gnf_debug_linear_tc) at the embedding flow's layer shapes; run under
`ncu --metrics gpu__time_duration.sum` for per-launch times.  GNF_GEMM_VARIANT selects the timing experiments of
csrc/gemm_tc.cu (1: alternate accumulators, 2: converters idle, 4: one row tile per slab)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_normalizing_flows_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 6873
maths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tc3x"]
for (k, n) in [(164, 2048), (2048, 2048), (2048, 100)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(k, n, device=dev) / k ** 0.5; b = torch.randn(n, device=dev)
    c = torch.empty(m, n, device=dev)
    wsb = lib.gnf_debug_linear_tc_workspace(k, n)
    ws = _lib.workspace(wsb, dev)
    for math in maths:
        for _ in range(2):
            _lib.check(lib.gnf_debug_linear_tc(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), m, k, n, 1, _lib.MATH[math], _lib.ptr(c),
                                               _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), "gnf_debug_linear_tc")
    torch.cuda.synchronize()
    ref = torch.relu(a.double() @ w.double() + b.double())
    print(k, n, "max err", float((c.double() - ref).abs().max()))
