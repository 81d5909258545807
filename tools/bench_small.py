"""Small-batch latency (BASELINE configs[3]: protein, B=256 graphs): eager launches vs one CUDA-graph replay."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from oracle import gnf_oracle as O
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200 import graph_data as GD
npz = np.load(os.path.join(ROOT, "tests", "golden", "graphs_protein_4_128.npz"))
ds = GD.GraphDataset(None, 14, structures=GD.structures_from_fixture(npz))
params = O.make_params(12345, 6, 14, 256, 5, last_layer_scale=0.05)
net = H.make_grevnet(params, 256, 5, device="cuda", math="tc3x")
for B in (64, 256, 1024):
    g = ds.draw_batch(B, np.random.default_rng(1)).to("cuda")
    n = g.nodes.shape[0]
    def timeit(fn, reps=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    eager = timeit(lambda: G.loss.mvn_log_prob_sum(*(lambda zl: (zl[0].nodes, zl[1]))(net.f64(g))))
    runner = G.loss.GraphedLogProb(net, g)
    graphed = timeit(lambda: runner())
    print(json.dumps({"family": "protein_4_128", "graphs": B, "nodes": n, "eager_us": eager * 1e6, "cuda_graph_us": graphed * 1e6,
                      "node_updates_per_s_graphed": n * 12 / graphed}))
    # use_batch_norm=True (the scripts' default): the whole chain in one library call vs the host mirror driving it half
    # step by half step (what sharded runs do)
    bn = H.make_grevnet(params, 256, 5, device="cuda", math="tc3x")
    bn.use_batch_norm = True
    row = {"family": "protein_4_128", "graphs": B, "nodes": n, "what": "use_batch_norm density pass"}
    for chain in (True, False):
        bn.bn_chain = chain
        row["chained_us" if chain else "per_half_step_us"] = timeit(lambda: G.loss.log_prob(bn, g)) * 1e6
    print(json.dumps(row))
