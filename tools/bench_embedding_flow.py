"""Density pass of the embedding flow at the defaults of train_grevnet_with_data.py:41-47,58-59,104-117 (node_embedding_dim
200, latent_dim 2048, 3 MLP layers, 10 coupling layers, dm_self_attn with kq = v = out = 64 and 1 head, batch of 32 fully
connected graphs of grid size, 100..361 nodes): the layered tensor-core path (k_gemm_tc + k_linear_tc) against the fp32
FFMA kernels on the same flow and batch.  CUDA events; one JSON line.  `profile` as argv[1]: one tc3x pass between
cudaProfilerStart/Stop for an ncu launch list."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    T, D, L, K = 10, 200, 2048, 3
    rng = np.random.default_rng(0)
    side = rng.integers(10, 20, size=(32, 2))
    n_node = (side[:, 0] * side[:, 1]).astype(np.int64)
    s, r = G.utils.senders_receivers(n_node)
    n = int(n_node.sum())
    nodes = (0.5 * rng.standard_normal((n, D))).astype(np.float32)
    g = O.GraphsTuple(nodes, None, np.asarray(r, np.int32), np.asarray(s, np.int32), None, n_node.astype(np.int32),
                      (n_node ** 2).astype(np.int32))
    attn = dict(num_heads=1, kq_dim=64, v_dim=64, out_dim=64, concat=True, residual=False, kq_dim_division=True,
                layer_norm=False)
    params = O.make_params(12345, T, D, L, K, block="dm_attn", act="relu", attn=attn, last_layer_scale=0.02)
    dg = H.to_device_graph(g)
    G.graphs.structure_of(dg)
    res = {"row": "f1/f4 embedding flow", "graphs": 32, "nodes": n, "edges": len(s), "T": T, "D": D, "L": L, "K": K}
    flops = n * 2 * T * 2 * 2.0 * ((D // 2 + 64) * L + (K - 2) * L * L + L * (D // 2))     # MLP layers only
    outs = {}
    for math in ("tc3x", "fp32"):
        net = H.make_grevnet(params, L, K, device="cuda", math=math)
        if len(sys.argv) > 1 and sys.argv[1] == "profile":
            if math != "tc3x":
                continue
            G.loss.log_prob(net, dg)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            G.loss.log_prob(net, dg)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            print("profiled embedding flow", n)
            return
        ms = timeit(lambda: G.loss.log_prob(net, dg), reps=3, warm=1)
        if math == "tc3x":          # reversible backward: recompute on the tensor cores, dX / dW GEMMs on the fp32 kernels
            G.graphs.transposed_structure_of(dg)
            z, _ = net.f64(dg)
            res["backward_ms"] = timeit(lambda: net.backward_from_z(dg, z.nodes, 1.0 / n), reps=2, warm=1)
        outs[math] = float(G.loss.log_prob(net, dg)["log_prob_xs"])
        res[f"density_pass_ms_{math}"] = ms
        res[f"node_updates_per_s_{math}"] = n * 2 * T / (ms * 1e-3)
        res[f"mlp_algorithmic_tflops_{math}"] = flops / (ms * 1e-3) / 1e12
        del net
    res["log_prob_rel_tc3x_vs_fp32"] = abs(outs["tc3x"] - outs["fp32"]) / abs(outs["fp32"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
