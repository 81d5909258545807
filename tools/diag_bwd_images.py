"""Decode the act/delta images of the LAST reversed half step (step 0, half 0) and compare with a torch fp64
recomputation of that half step.  T is forced to 1 so the half step's inputs are known."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
from graph_normalizing_flows_b200 import _lib
D, L, K, NG, T = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (14, 128, 4, 9, 2)
Hh = D // 2
rng = np.random.default_rng(21)
g = H.random_batch(rng, NG, 4, 25, D=D)
params = O.make_params(8, T, D, L, K, agg="sum", block="concat", act="leaky_relu", last_layer_scale=0.2)
n = g.nodes.shape[0]
net = H.make_grevnet(params, L, K, device="cuda")
net._keep_backward_workspace = True
dg = H.to_device_graph(g, "cuda")
z, _ = net.f64(dg)
grads = net.backward_from_z(dg, z.nodes, 1.0 / n, math="tc3x_bf16")
torch.cuda.synchronize()
ws = net._last_backward_workspace.cpu().numpy()
lay = (C.c_int64 * 8)()
lib = _lib.load()
_lib.check(lib.gnf_debug_bwd_layout(net._flow.ensure(net.params.detach()), n, lay))
lay = list(lay)
tiles = (n + 127) // 128

def decode(off_bytes, F, idx):            # idx-th image array of [tiles][2][F*128]
    per = tiles * 2 * F * 128
    raw = ws[off_bytes + idx * per * 2: off_bytes + (idx + 1) * per * 2].view(np.uint16).reshape(tiles, 2, F * 128)
    f32 = (raw.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    val = f32[:, 0] + f32[:, 1]
    out = np.zeros((tiles * 128, F))
    nn, ff = np.meshgrid(np.arange(128), np.arange(F), indexing="ij")
    pos = (nn >> 3) * (F * 8) + (ff >> 3) * 64 + (nn & 7) * 8 + (ff & 7)
    for t in range(tiles):
        out[t * 128:(t + 1) * 128] = val[t][pos]
    return out[:n]

x = g.nodes.astype(np.float64)
x0 = x[:, :Hh]
agg = np.zeros_like(x0)
np.add.at(agg, g.receivers, x0[g.senders])
h = np.concatenate([x0, agg], 1)
h_img = decode(lay[2], 16, 0)
print("h image err", np.abs(h_img[:, :2 * Hh] - h).max())
for m, which in enumerate(("s", "t")):
    mlp = params[which][0][0]
    Ws = [np.asarray(w, np.float64) for (w, b) in mlp]
    bs = [np.asarray(b, np.float64) for (w, b) in mlp]
    acts = []
    a = h
    for l in range(K - 1):
        pre = a @ Ws[l] + bs[l]
        a = np.maximum(pre, 0.2 * pre)
        acts.append((pre, a))
    gtop = decode(lay[3], 16, m)[:, :Hh]
    delta = gtop
    deltas = {}
    for l in range(K - 1, 0, -1):
        delta = (delta @ Ws[l].T) * np.where(acts[l - 1][0] > 0, 1.0, 0.2)
        deltas[l - 1] = delta
    for l in range(K - 1):
        a_img = decode(lay[0], L, m * (K - 1) + l)
        d_img = decode(lay[1], L, m * (K - 1) + l)
        ea = np.abs(a_img - acts[l][1])
        ed = np.abs(d_img - deltas[l])
        dm = np.abs(deltas[l]).max()
        bad = np.argwhere(ed > 1e-3 * dm)
        print(f"{which} layer {l}: act err {ea.max():.2e} (max {np.abs(acts[l][1]).max():.2e})  delta err {ed.max():.2e} (max {dm:.2e})  bad {len(bad)}")
        # weight/bias gradient of layer l+1 (hidden) and layer 0 from the reference quantities vs the grads vector
        dims = [(D, L)] + [(L, L)] * (K - 2) + [(L, Hh)]
        per = sum(i * o + o for i, o in dims)
        gv = grads.double().cpu().numpy()[(2 * T if m else 0) * per:][:per]
        off = sum(i * o + o for i, o in dims[:l])
        a_in = h if l == 0 else acts[l - 1][1]
        dW = a_in.T @ deltas[l]
        db = deltas[l].sum(0)
        gW = gv[off:off + dims[l][0] * dims[l][1]].reshape(dims[l])
        gb = gv[off + dims[l][0] * dims[l][1]:][:dims[l][1]]
        print(f"   dW{l} err {np.abs(gW - dW).max() / np.abs(dW).max():.2e}  db{l} err {np.abs(gb - db).max() / np.abs(db).max():.2e}")
        if len(bad):
            rows, cols = bad[:, 0], bad[:, 1]
            print("   bad rows:", np.unique(rows)[:20], " bad cols:", np.unique(cols)[:40])
            r, c = bad[0]
            print("   e.g.", r, c, "got", d_img[r, c], "want", deltas[l][r, c], "pre", acts[l][0][r, c], "ratio", d_img[r, c] / deltas[l][r, c])
