"""Steady-state power / clock trace: >= 3 s of back-to-back density passes at the bench workload,
NVML sampled every 10 ms (SM clock, power, throttle reasons), next to the same for a cuBLAS bf16 8192^3 matmul loop.
Writes gpurun_out/power_trace.json: per phase the time series summary and the achieved step time."""
import json, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import bench
import graph_normalizing_flows_b200 as G
from graph_normalizing_flows_b200.graphs import concat_structures
import pynvml

pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.halt = [], threading.Event()

    def run(self):
        while not self.halt.is_set():
            self.rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                              pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0,
                              pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)))
            time.sleep(0.01)


def phase(name, fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = Sampler(); s.start()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); n += 10
    dt = time.perf_counter() - t0
    s.halt.set(); s.join(timeout=1)
    rows = [r for r in s.rows if r[0] - t0 > 0.5]            # steady state: drop the first 0.5 s
    clk = np.array([r[1] for r in rows]); pw = np.array([r[2] for r in rows])
    reasons = 0
    for r in rows:
        reasons |= r[3]
    return {"phase": name, "seconds": dt, "calls": n, "ms_per_call": dt / n * 1e3, "samples": len(rows),
            "sm_mhz_median": float(np.median(clk)), "sm_mhz_min": float(clk.min()), "sm_mhz_max": float(clk.max()),
            "power_w_median": float(np.median(pw)), "power_w_max": float(pw.max()), "power_w_min": float(pw.min()),
            "throttle_reason_mask_or": hex(reasons),
            "power_limit_w": pynvml.nvmlDeviceGetEnforcedPowerLimit(hnd) / 1000.0,
            "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM)}


wl = bench.WORKLOADS["community_medium"]
structs, feats = bench.make_block(wl, 4096, bench.SEED)
g = concat_structures(structs, nodes=feats).to("cuda")
out = []
for math in ("tc3x", "bf16"):
    net = H.make_grevnet(bench.make_oracle_params(wl), 256, 5, device="cuda", math=math)
    out.append(phase(f"GRevNet.f + log-prob, community_medium B=4096, {math}",
                     lambda: G.loss.mvn_log_prob_sum(*(lambda zl: (zl[0].nodes, zl[1]))(net.f64(g))), 4.0))
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
r = phase("cuBLAS bf16 8192^3 matmul", lambda: torch.matmul(a, b), 4.0)
r["tflops"] = 2 * 8192 ** 3 / (r["ms_per_call"] * 1e-3) / 1e12
out.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "power_trace.json"), "w"), indent=1)
for r in out:
    print(json.dumps(r))
