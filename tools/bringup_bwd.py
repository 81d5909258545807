"""Bring-up of the tensor-core backward (backward_tc.cu) on a GPU box.  Every stage runs in its own
subprocess under a timeout so a hung kernel costs seconds, not the box.

    python tools/bringup_bwd.py            # all stages
    python tools/bringup_bwd.py dw 0       # one stage in-process (flags)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage_dw(flags):
    import torch
    from graph_normalizing_flows_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(3)
    res = []
    for (fa, fb, n, parts, splits) in [(256, 256, 128, 2, 1), (256, 256, 128, 1, 1), (256, 16, 300, 2, 2),
                                       (128, 128, 1000, 2, 3), (128, 16, 77, 1, 1), (256, 256, 5000, 2, 7)]:
        a = torch.randn(n, fa, generator=gen).to(dev)
        b = (torch.randn(n, fb, generator=gen) * 1e-3).to(dev)
        out = torch.empty(fa, fb, device=dev)
        tiles = (n + 127) // 128
        wsb = 4 * tiles * 128 * (fa + fb) + 4 * splits * fa * fb + 4096
        ws = _lib.workspace(wsb, dev)
        _lib.check(lib.gnf_debug_dw_gemm(_lib.ptr(a), _lib.ptr(b), n, fa, fb, parts, splits, _lib.ptr(out), _lib.ptr(ws),
                                         wsb, _lib.stream_ptr(dev)), "dw")
        torch.cuda.synchronize()
        ref = a.double().T @ b.double()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        res.append({"fa": fa, "fb": fb, "n": n, "parts": parts, "splits": splits, "rel_err": err})
    print(json.dumps({"stage": "dw", "flags": flags, "results": res}))


def stage_bwd(flags, L=128, K=4, D=14, T=2, n_graphs=9):
    import numpy as np
    import torch
    import helpers as H
    from oracle import gnf_oracle as O
    from graph_normalizing_flows_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(21)
    g = H.random_batch(rng, n_graphs, 4, 25, D=D)
    params = O.make_params(8, T, D, L, K, agg="sum", block="concat", act="leaky_relu", last_layer_scale=0.2)
    net = H.make_grevnet(params, L, K, device="cuda", math="tc3x")
    dg = H.to_device_graph(g, "cuda")
    z, _ = net.f64(dg)
    n = g.nodes.shape[0]
    out = {}
    for math in ("fp32", "tc3x", "bf16"):
        grads, x = net.backward_from_z(dg, z.nodes, 1.0 / n, return_x=True, math=math)
        torch.cuda.synchronize()
        out[math] = (grads.cpu().numpy().astype(np.float64), x.cpu().numpy())
    ref = out["fp32"][0]
    rep = {"stage": "bwd", "flags": flags, "L": L, "K": K, "n": int(n)}
    per = net._flow.param_count // (4 * T)
    for math in ("tc3x", "bf16"):
        got = out[math][0]
        rep[math] = {"max_rel": float(np.abs(got - ref).max() / np.abs(ref).max()),
                     "cos": float(got @ ref / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-300)),
                     "x_err": float(np.abs(out[math][1] - g.nodes).max()), "finite": bool(np.isfinite(got).all())}
        # per layer of MLP 0 (which=s, half 0, step 0): where does it go wrong?
        off = 0
        lay = []
        dims = [(D, L)] + [(L, L)] * (K - 2) + [(L, D // 2)]
        for (i, o) in dims:
            for nm, sz in (("W", i * o), ("b", o)):
                a, b = got[off:off + sz], ref[off:off + sz]
                lay.append(f"{nm}{len(lay)//2}:{np.abs(a-b).max()/(np.abs(b).max()+1e-30):.1e}")
                off += sz
        rep[math]["mlp0"] = lay
    print(json.dumps(rep))


def stage_time(B):
    import torch
    import helpers as H
    import bench
    host = bench.make_batch(B, 12345)
    net = H.make_grevnet(bench.make_oracle_params(), 256, 5, device="cuda", math="tc3x")
    g = host.to("cuda")
    z, _ = net.f64(g)
    n = g.nodes.shape[0]
    rep = {"stage": "time", "graphs": B, "nodes": int(n)}
    for math in ("tc3x", "bf16", "fp32"):
        grads = torch.zeros_like(net.params.detach())
        for _ in range(2):
            net.backward_from_z(g, z.nodes, 1.0 / n, grads=grads, math=math)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 3
        for _ in range(reps):
            net.backward_from_z(g, z.nodes, 1.0 / n, grads=grads, math=math)
        e1.record()
        torch.cuda.synchronize()
        rep[math + "_ms"] = e0.elapsed_time(e1) / reps
    print(json.dumps(rep))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        st = sys.argv[1]
        if st == "dw":
            stage_dw(int(sys.argv[2]))
        elif st == "bwd":
            stage_bwd(int(sys.argv[2]), *[int(v) for v in sys.argv[3:]])
        elif st == "time":
            stage_time(int(sys.argv[2]))
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup_bwd.log"), "a")
    stages = [["dw", "0"], ["bwd", "0", "128", "4", "14", "2", "9"], ["bwd", "0", "256", "5", "14", "1", "40"],
              ["bwd", "0", "128", "3", "6", "2", "9"], ["time", "512"], ["time", "4096"]]
    for s in stages:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + s, capture_output=True, text=True, timeout=100)
            msg = f"{s}: rc={r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-1500:]}\n"
        except subprocess.TimeoutExpired:
            msg = f"{s}: TIMEOUT\n"
        print(msg)
        log.write(msg)
        log.flush()
