"""Standalone gather+segment-sum at the HBM-roofline size (for an ncu --set full capture)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import _lib
from graph_normalizing_flows_b200.graphs import concat_structures
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
structs, feats = bench.make_block(bench.WORKLOADS["community_medium"], B, bench.SEED)
big = concat_structures(structs, nodes=feats)
h = 7
gb = big.replace(nodes=np.ascontiguousarray(big.nodes[:, :h])).to("cuda")
st = G.graphs.BatchStructure(gb.senders, gb.receivers, gb.nodes.shape[0])
out = torch.empty_like(gb.nodes)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.profiler.start()
for i in range(4):
    flush.fill_(i)
    _lib.check(lib.gnf_gather_segment_sum(_lib.ptr(gb.nodes), h, _lib.ptr(st.rowptr), _lib.ptr(st.csr_senders),
                                          gb.nodes.shape[0], 0, _lib.ptr(out), _lib.stream_ptr()))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
n, e = gb.nodes.shape[0], st.n_edges
print("N", n, "E", e, "alg bytes", e * (8 + 4 * h) + 4 * n * h, "compulsory (CSR form)", e * 4 + (n + 1) * 4 + 2 * 4 * n * h)
