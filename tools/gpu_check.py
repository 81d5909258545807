"""First-contact diagnostic for the GPU box: runs each kernel family against the oracle and
prints error magnitudes.  Each math mode runs in its own subprocess with a timeout so a
deadlocked kernel cannot eat the GPU budget."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(math, L, K, D, T, n_graphs, block, agg):
    import numpy as np
    import torch
    import helpers as H
    from oracle import gnf_oracle as O
    import graph_normalizing_flows_b200 as G
    rng = np.random.default_rng(7)
    g = H.random_batch(rng, n_graphs, 5, 40, D=D)
    params = O.make_params(3, T, D, L, K, agg=agg, block=block, last_layer_scale=0.05)
    t0 = time.time()
    z_ref, ldj_ref = O.grevnet_f(g.nodes, g.senders, g.receivers, params)
    lp_ref = O.log_prob(z_ref, ldj_ref, g.n_node)
    p64 = O.cast_params(params, np.float64)
    z64, ldj64 = O.grevnet_f(g.nodes.astype(np.float64), g.senders, g.receivers, p64)
    lp64 = O.log_prob(z64, ldj64, g.n_node)
    net = H.make_grevnet(params, L, K, math=math)
    dg = H.to_device_graph(g)
    out = G.loss.log_prob(net, dg, return_z=True)
    torch.cuda.synchronize()
    z = out["z"].nodes.cpu().numpy()
    x_back = net(out["z"], inverse=False).nodes.cpu().numpy()
    res = {
        "math": math, "L": L, "K": K, "D": D, "T": T, "N": int(g.nodes.shape[0]), "E": int(len(g.senders)),
        "block": block, "agg": agg,
        "z_maxabs_vs_fp32": float(np.abs(z - z_ref).max()),
        "z_maxabs_vs_fp64": float(np.abs(z - z64).max()),
        "oracle32_z_vs_fp64": float(np.abs(z_ref - z64).max()),
        "ldj": float(out["log_det_jacobian"]), "ldj_ref32": float(ldj_ref), "ldj_ref64": float(ldj64),
        "logprob": float(out["log_prob_xs"]), "logprob_ref32": float(lp_ref["log_prob_xs"]),
        "logprob_ref64": float(lp64["log_prob_xs"]),
        "logprob_rel_vs_fp64": H.rel_err(out["log_prob_xs"], lp64["log_prob_xs"]),
        "oracle32_logprob_rel_vs_fp64": H.rel_err(lp_ref["log_prob_xs"], lp64["log_prob_xs"]),
        "roundtrip_maxabs": float(np.abs(x_back - g.nodes).max()),
        "secs": time.time() - t0,
    }
    print("RESULT " + json.dumps(res), flush=True)


def main():
    cases = []
    for math in ("fp32", "tc3x", "tc3x_bf16", "bf16"):
        cases.append((math, 256, 5, 14, 2, 24, "concat", "sum"))
    cases += [("fp32", 64, 3, 4, 2, 6, "agg_then", "mean"), ("tc3x", 128, 4, 14, 2, 24, "concat", "sum"),
              ("tc3x", 256, 2, 2, 1, 24, "agg_then", "mean"), ("tc3x", 256, 5, 14, 6, 300, "concat", "sum")]
    for c in cases:
        cmd = [sys.executable, __file__, "--child"] + [str(x) for x in c]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
            print(f"== {c} rc={r.returncode} {time.time()-t0:.1f}s")
            for ln in tail:
                print("   ", ln)
        except subprocess.TimeoutExpired:
            print(f"== {c} TIMEOUT (kernel hang?)")
        sys.stdout.flush()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        a = sys.argv[2:]
        child(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), a[6], a[7])
    else:
        main()
