"""Per-(MLP, layer) error of the backward paths vs fp64 autograd of the oracle restatement."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
from oracle import gnf_oracle_torch as OT
D, T, L, K, NG = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (14, 2, 128, 4, 9)
rng = np.random.default_rng(21)
g = H.random_batch(rng, NG, 4, 25, D=D)
params = O.make_params(8, T, D, L, K, agg="sum", block="concat", act="leaky_relu", last_layer_scale=0.2)
n = g.nodes.shape[0]
loss_ref, ref = OT.loss_and_grads(g.nodes, g.senders, g.receivers, params, 1.0 / n)
net = H.make_grevnet(params, L, K, device="cuda")
dg = H.to_device_graph(g, "cuda")
z, _ = net.f64(dg)
gmax = np.abs(ref).max()
dims = [(D, L)] + [(L, L)] * (K - 2) + [(L, D // 2)]
per = sum(i * o + o for i, o in dims)
for math in ("tc3x", "tc3x_bf16", "tc2x"):
    got = net.backward_from_z(dg, z.nodes, 1.0 / n, math=math).double().cpu().numpy()
    print(math, "global max err / max|g| = %.2e" % (np.abs(got - ref).max() / gmax))
    rows = []
    for m in range(4 * T):
        off = m * per
        cells = []
        for (i, o) in dims:
            for sz in (i * o, o):
                a, b = got[off:off + sz], ref[off:off + sz]
                cells.append("%.0e/%.0e" % (np.abs(a - b).max() / (np.abs(b).max() + 1e-300), np.abs(a - b).max() / gmax))
                off += sz
        rows.append("  mlp%02d " % m + " ".join(cells))
    print("\n".join(rows))
