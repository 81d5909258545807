"""Measurements for the SURVEY §8(f) rows on one B200 (CUDA events; JSON lines):
  f1  dm_self_attn GNN at the run_grevnet.py defaults (D=2, 8 heads, kq=v=10, out 80, L=256, K=5, T=12 coupling layers)
      on fully connected synthetic graphs: density pass and training-step evaluation, vs the torch-CPU restatement
  f3  decode tail pred_adj on the community_medium batch
  a9  batch-norm density pass / training step on the bench workload
"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
import bench


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


def f1(n_graphs=2048, T=12, math=None):
    rng = np.random.default_rng(0)
    n_node = rng.integers(10, 41, size=n_graphs)
    s, r = G.utils.senders_receivers(n_node)
    nodes = rng.standard_normal((int(n_node.sum()), 2)).astype(np.float32)
    g = O.GraphsTuple(nodes, None, r, s, None, n_node.astype(np.int32), (n_node ** 2).astype(np.int32))
    attn = dict(num_heads=8, kq_dim=10, v_dim=10, out_dim=80, concat=True, residual=False, kq_dim_division=False)
    params = O.make_params(13, T, 2, 256, 5, block="dm_attn", act="relu", attn=attn, last_layer_scale=0.05)
    net = H.make_grevnet(params, 256, 5, device="cuda", math=math)
    dg = H.to_device_graph(g)
    n, e = nodes.shape[0], len(s)
    G.graphs.structure_of(dg); G.graphs.transposed_structure_of(dg)
    fwd = timeit(lambda: G.loss.log_prob(net, dg))
    z, _ = net.f64(dg)
    bwd = timeit(lambda: net.backward_from_z(dg, z.nodes, 1.0 / n), reps=3, warm=1)
    # CPU restatement on a bounded sample
    from oracle import gnf_oracle_torch as OT
    k = 128
    n_s = int(n_node[:k].sum()); e_s = int((n_node[:k] ** 2).sum())
    p = OT.params_to_torch(params)
    xs, ss, rs = torch.from_numpy(nodes[:n_s]), torch.from_numpy(s[:e_s]).long(), torch.from_numpy(r[:e_s]).long()
    torch.set_num_threads(os.cpu_count() or 1)
    OT.grevnet_f(xs, ss, rs, p)
    t0 = time.perf_counter(); OT.grevnet_f(xs, ss, rs, p); cpu = time.perf_counter() - t0
    print(json.dumps({"row": "f1", "graphs": n_graphs, "nodes": n, "edges": e, "T": T, "math": net.math,
                      "density_pass_ms": fwd, "node_updates_per_s": n * 2 * T / (fwd * 1e-3),
                      "backward_ms": bwd, "cpu_node_updates_per_s": n_s * 2 * T / cpu, "cpu_cores": torch.get_num_threads(),
                      "cpu_sample": f"first {k} graphs ({n_s} nodes)"}))


def f3_and_bn():
    wl = bench.WORKLOADS["community_medium"]
    host = bench.make_global_batch(wl, 1, 4096)
    dg = host.to("cuda")
    n = dg.nodes.shape[0]
    t = timeit(lambda: G.loss.pred_adj(dg))
    n2 = int((host.n_node.astype(np.int64) ** 2).sum())
    print(json.dumps({"row": "f3", "what": "pred_adj(scaled_hacky_sigmoid_l2), per-graph blocks", "nodes": n, "adj_entries": n2,
                      "ms": t, "GB_per_s_written": n2 * 4 / (t * 1e-3) / 1e9}))
    net = H.make_grevnet(bench.make_oracle_params(wl), 256, 5, device="cuda", math="tc3x")
    net.use_batch_norm = True
    fwd = timeit(lambda: G.loss.log_prob(net, dg))
    step = timeit(lambda: net.loss_and_grad(dg), reps=5, warm=3)
    print(json.dumps({"row": "a9/f4", "what": "use_batch_norm=True on the bench workload (B=4096)", "nodes": n,
                      "density_pass_ms": fwd, "train_step_eval_ms": step}))


if __name__ == "__main__":
    f1(math=sys.argv[1] if len(sys.argv) > 1 else None)
    if len(sys.argv) <= 2:
        f3_and_bn()
