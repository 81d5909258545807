"""Accuracy of every math mode on a mid-size community_medium batch (oracle fp64 as truth)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
import bench
host = bench.make_batch(96, 12345)
params = bench.make_oracle_params()
z64, ldj64 = O.grevnet_f(host.nodes.astype(np.float64), host.senders, host.receivers, O.cast_params(params, np.float64))
lp64 = O.log_prob(z64, ldj64, host.n_node)
z32, ldj32 = O.grevnet_f(host.nodes, host.senders, host.receivers, params)
lp32 = O.log_prob(z32, ldj32, host.n_node)
print(json.dumps({"mode": "oracle_fp32", "N": int(host.nodes.shape[0]), "z_maxabs": float(np.abs(z32 - z64).max()),
                  "logprob_rel": H.rel_err(lp32["log_prob_xs"], lp64["log_prob_xs"]),
                  "ldj_abs": abs(float(ldj32) - float(ldj64))}))
dg = host.to("cuda")
for math in ("fp32", "tc3x", "tc3x_bf16", "tc2x", "bf16"):
    net = H.make_grevnet(params, 256, 5, device="cuda", math=math)
    out = G.loss.log_prob(net, dg, return_z=True)
    xb = net(out["z"], inverse=False).nodes
    z = out["z"].nodes.cpu().numpy()
    print(json.dumps({"mode": math, "z_maxabs": float(np.abs(z - z64).max()), "z_rms": float(np.sqrt(((z - z64) ** 2).mean())),
                      "logprob_rel": H.rel_err(out["log_prob_xs"], lp64["log_prob_xs"]),
                      "ldj_abs": abs(float(out["log_det_jacobian"]) - float(ldj64)),
                      "roundtrip_maxabs": float((xb - dg.nodes).abs().max())}))
