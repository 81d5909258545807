#!/usr/bin/env python
"""bench.py -- GRevNet fwd+logdet node-updates/sec (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU restatement of the reference path, same config
    python bench.py --workload {community_medium,grid_t12_bf16,protein_b256,citeseer,mixed}

A "step" is one pass of the hot path over one batch: GRevNet.f (2T fused half-coupling kernels) + log-prob
assembly.  Default workload = BASELINE configs[1]: community_medium_4_128 graphs, B graphs per GPU drawn with
replacement (rng 12345 + r for the r-th 4096-graph block), D=14, T=6, sum_concat_then_mlp, L=256, K=5, leaky_relu,
weights Glorot/trunc-normal seed 12345 with the last layer x0.05 (keeps the reference's unclamped exp(s) finite).
node_updates = N * T * 2.  The other workloads are BASELINE configs[2..4].

Multi-GPU (torchrun): every rank builds the same GLOBAL batch, `GraphShardedGRevNet.local_shard` assigns whole
graphs to ranks (greedy LPT on the cost model), each rank runs the fused kernels on its shard and ONE NCCL
all-reduce of the fp64 4-vector per step assembles the batch log-likelihood (`log_prob_async`: the collective
runs on a side stream).  The line carries per-rank ms, per-rank nodes and the all-reduce time.

value : device-resident throughput (shard + CSR already in HBM), CUDA events per step, L2 flushed between steps,
        max over ranks.
e2e   : same metric through the public API from pinned HOST buffers: H2D of the packed shard (nodes, senders,
        receivers, n_node, n_edge), index validation + CSR build, f, log-prob, all-reduce, D2H of the 4 scalars
        (z stays on the device, as in the reference's sess.run fetch of the scalars) -- every step, host wall
        clock.  The next step's batch is staged on a side stream (graphs.BatchPrefetcher) while this one computes.
parity: at N=1 the torch-CPU restatement of the reference runs on the EXACT timed batch and weights;
        log-prob / z / log-det differences are on the line (and that run is the cpu_baseline).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GRevNet fwd+logdet node-updates/sec"
UNIT = "node-updates/s"
D, L, K = 14, 256, 5
SEED = 12345
ALL_FAMILIES = ["community_medium_4_128", "grid_4_128", "protein_4_128", "citeseer_4_128", "caveman_4_128"]

# BASELINE.json configs[1..4] made concrete
WORKLOADS = {
    "community_medium": dict(families=["community_medium_4_128"], T=6, math="tc3x", last_scale=0.05, graphs_per_gpu=4096,
                             scaling="weak",
                             desc="community_medium_4_128 graphs, 6-step GRevNet, node_dim=14, sum_concat_then_mlp L=256 K=5 leaky_relu"),
    "grid_t12_bf16": dict(families=["grid_4_128"], T=12, math="bf16", last_scale=0.02, graphs_per_gpu=512, scaling="weak",
                          desc="grid_4_128 graphs, 12-step GRevNet, bf16 single-pass fused coupling+message kernel, node_dim=14, L=256 K=5"),
    "protein_b256": dict(families=["protein_4_128"], T=6, math="tc3x", last_scale=0.05, total_graphs=256, scaling="strong",
                         desc="protein_4_128 graphs, batch=256 graphs TOTAL, graph-sharded (cost-balanced LPT), 6-step GRevNet, node_dim=14"),
    "citeseer": dict(families=["citeseer_4_128"], T=6, math="tc3x", last_scale=0.05, total_graphs=605, scaling="strong",
                     no_resample=True, train=True,
                     desc="citeseer_4_128, the whole train split (605 graphs) as one batch, graph-sharded, 6-step GRevNet + grad all-reduce"),
    "mixed": dict(families=ALL_FAMILIES, T=6, math="tc3x", last_scale=0.05, graphs_per_gpu=1280, scaling="weak", train=True,
                  desc="mixed-family batch (equal draws from the five 4_128 families, shuffled), 6-step GRevNet, NCCL log-prob and gradient all-reduce"),
}


def flops_per_node_update():
    return 2 * 2 * (D * L + (K - 2) * L * L + L * (D // 2))      # 807 936  


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(path, kernel=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` summary
    (profiles/, written by tools/summarize_ncu.py); None if the file is absent."""
    try:
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for line in open(path):
            c = line.strip().split(",")
            if c[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                vals = [float(v) for v in c[2:]]
                tot += unit[c[1]] * sum(vals) / len(vals)
        return tot or None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
_FIXTURES = {}


def family_structs(family):
    from graph_normalizing_flows_b200 import graph_data as GD
    if family not in _FIXTURES:
        npz = np.load(os.path.join(ROOT, "tests", "golden", f"graphs_{family}.npz"))
        _FIXTURES[family] = GD.structures_from_fixture(npz)
    return _FIXTURES[family]


def make_block(wl, n_graphs, seed):
    """One block of `n_graphs` graphs: (structs, nodes).  Single family: idx then features from one Generator
    (identical to GraphDataset.draw_batch, so the N=1 default batch is round 1's).  Several families: equal draws
    from each, shuffled."""
    rng = np.random.default_rng(seed)
    fams = wl["families"]
    if wl.get("no_resample"):
        structs = list(family_structs(fams[0]))[:n_graphs]
    elif len(fams) == 1:
        pool = family_structs(fams[0])
        structs = [pool[i] for i in rng.integers(0, len(pool), size=n_graphs)]
    else:
        structs = []
        per = n_graphs // len(fams)
        for fam in fams:
            pool = family_structs(fam)
            structs += [pool[i] for i in rng.integers(0, len(pool), size=per)]
        structs = [structs[i] for i in rng.permutation(len(structs))]
    n = sum(s[0] for s in structs)
    return structs, rng.standard_normal((n, D)).astype(np.float32)


def make_global_batch(wl, world, graphs_per_gpu):
    """The job's global batch on the host.  weak scaling: `world` blocks of graphs_per_gpu graphs (block r seeded
    SEED + r); strong scaling: one block of wl['total_graphs'] graphs whatever the world size."""
    from graph_normalizing_flows_b200.graphs import concat_structures
    if wl["scaling"] == "strong":
        blocks = [make_block(wl, wl["total_graphs"], SEED)]
    else:
        blocks = [make_block(wl, graphs_per_gpu, SEED + r) for r in range(world)]
    structs = [s for b in blocks for s in b[0]]
    return concat_structures(structs, nodes=np.concatenate([b[1] for b in blocks], axis=0))


def make_batch(n_graphs, seed, workload="community_medium"):
    """One block of `n_graphs` graphs of a workload as a host GraphsTuple (tools/)."""
    from graph_normalizing_flows_b200.graphs import concat_structures
    structs, feats = make_block(WORKLOADS[workload], n_graphs, seed)
    return concat_structures(structs, nodes=feats)


def make_oracle_params(wl=None):
    from oracle import gnf_oracle as O
    wl = wl or WORKLOADS["community_medium"]
    return O.make_params(SEED, wl["T"], D, L, K, agg="sum", block="concat", act="leaky_relu",
                         bias_init_stddev=0.1, last_layer_scale=wl["last_scale"])


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.period = period
        self.index, self.samples, self.power, self.reasons, self.max_mhz = index, [], [], set(), None
        self._halt = threading.Event()
        self.armed = False            # NVML's first queries are slow and can hold up launches: the thread starts before
                                      # the warm-up and only RECORDS between arm() and stop()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None or os.environ.get("GNF_BENCH_NO_SAMPLER"):
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                # clocks and throttle reasons only: nvmlDeviceGetPowerUsage was seen to hold the driver for 100-200 ms
                # now and then, which stalls the launching thread (a 10-step tc2x run lost 200 ms to it); power is
                # logged by tools/power_trace.py instead
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.armed:
                    self.samples.append(clk)
                    for bit, nm in names.items():
                        if mask & bit:
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def arm(self):
        self.armed = True

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": (max(self.power) if self.power else None)}


# ------------------------------------------------------------------------------------------------
def cpu_fp64_truth(wl, host):
    """One float64 pass of the same restatement on the same batch: the yardstick both fp32 arms are measured against
    in the parity object (untimed)."""
    import torch
    from oracle import gnf_oracle_torch as OT
    torch.set_num_threads(os.cpu_count() or 1)
    p = OT.params_to_torch(make_oracle_params(wl), torch.float64)
    nodes = torch.from_numpy(host.nodes).double()
    s, r = torch.from_numpy(host.senders).long(), torch.from_numpy(host.receivers).long()
    z, ldj = OT.grevnet_f(nodes, s, r, p)
    return {"z": z.numpy(), "ldj": float(ldj), "log_prob_xs": float(OT.log_prob_xs(z, ldj))}


def cpu_reference_run(wl, host, steps, warmup, keep_outputs=False):
    """The reference's CPU path (torch-CPU restatement, all host threads) on `host` (a GraphsTuple of numpy arrays)."""
    import torch
    from oracle import gnf_oracle_torch as OT
    torch.set_num_threads(os.cpu_count() or 1)
    p = OT.params_to_torch(make_oracle_params(wl))
    nodes = torch.from_numpy(host.nodes)
    s, r = torch.from_numpy(host.senders).long(), torch.from_numpy(host.receivers).long()
    n = nodes.shape[0]
    times = []
    z = ldj = lp = None
    parts64 = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        last = keep_outputs and i == warmup + steps - 1
        z, ldj = OT.grevnet_f(nodes, s, r, p, ldj64_out=parts64 if last else None)
        lp = float(OT.log_prob_xs(z, ldj))
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    out = {"value": n * wl["T"] * 2 / sec, "sec_per_step": sec, "n_nodes": n, "n_edges": int(len(host.senders)),
           "cores": torch.get_num_threads(), "log_prob_xs": lp, "ldj": float(ldj),
           "sample": f"the rank-0 batch of the b200 arm itself: {len(host.n_node)} graphs, N={n} nodes, E={len(host.senders)} "
                     f"edges; one GRevNet.f + log-prob per step, {steps} timed steps after {warmup} warm-up, torch-CPU fp32, "
                     f"{torch.get_num_threads()} threads"}
    if keep_outputs:
        out["z"] = z.numpy()
        out["ldj_f64_sums"] = float(np.sum(parts64))              # same s, every reduce_sum accumulated in float64
    return out


def base_config(wl, name, world, graphs_per_gpu, n_nodes, n_edges):
    """Identical in both arms (the driver compares them)."""
    return {"workload": wl["desc"], "workload_name": name,
            "graphs_per_gpu": graphs_per_gpu if wl["scaling"] == "weak" else None,
            "total_graphs": wl.get("total_graphs") if wl["scaling"] == "strong" else graphs_per_gpu * world,
            "T": wl["T"], "D": D, "L": L, "K": K, "parallelism": f"graph-sharded dp{world} (LPT over whole graphs)",
            "last_layer_scale": wl["last_scale"], "n_nodes_rank0": n_nodes, "n_edges_rank0": n_edges,
            "l2": "b200 arm: flushed between timed steps (256 MB write); reference arm: CPU, not applicable",
            "node_update": "one node through one half coupling (aggregate + s-MLP + t-MLP + affine)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="community_medium", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs-per-gpu", type=int, default=None, help="weak-scaling workloads: graphs per GPU")
    ap.add_argument("--math", default=None, choices=["tc3x", "fp32", "bf16", "tc3x_bf16", "tc2x"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the row-f2 backward timing")
    ap.add_argument("--no-seg", action="store_true", help="skip the scatter-reduce HBM roofline leg")
    ap.add_argument("--seg-graphs", type=int, default=65536, help="batch for the scatter-reduce HBM roofline")
    ap.add_argument("--profile", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return                                                       # rank 0 alone runs the CPU arm
    wl = WORKLOADS[args.workload]
    T = wl["T"]
    math = args.math or wl["math"]
    gpg = args.graphs_per_gpu or wl.get("graphs_per_gpu") or wl.get("total_graphs")
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    F = flops_per_node_update()

    from graph_normalizing_flows_b200 import sharding as SH
    global_host = make_global_batch(wl, world, gpg)
    parts = SH.partition_graphs(global_host.n_node, global_host.n_edge, world)

    # ---------------------------------------------------------------- reference arm (CPU) ---
    if args.impl == "reference":
        host = SH.shard_graphs_tuple(global_host, parts[0])          # rank 0's shard: the b200 arm's rank-0 batch
        r = cpu_reference_run(wl, host, args.steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": r["sec_per_step"] * 1e3,
                "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": base_config(wl, args.workload, world, gpg, r["n_nodes"], r["n_edges"]),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "log_prob_xs": r["log_prob_xs"]}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------------- B200 arm -------
    import torch
    import torch.distributed as dist
    import graph_normalizing_flows_b200 as G
    from graph_normalizing_flows_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE JSON line: whatever libraries write to fd 1 (NCCL prints its version banner there)
    # is sent to stderr; the line itself goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = load_peaks()

    net = H.make_grevnet(make_oracle_params(wl), L, K, device=dev, math=math)
    sharded = SH.GraphShardedGRevNet(net, peer_memory=False if os.environ.get("GNF_NO_PEER") else "auto")
    assert sharded.world_size == world and sharded.rank == rank
    sharded.broadcast_parameters(0)
    host = sharded.local_shard(global_host)                       # this rank's whole graphs (same LPT on every rank)
    n_nodes, n_edges = int(host.nodes.shape[0]), int(len(host.senders))
    n_global = int(global_host.nodes.shape[0])
    node_updates_global = n_global * T * 2
    graph = host.to(dev)
    torch.cuda.synchronize()
    G.graphs.structure_of(graph)                                  # CSR build, reported separately below

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def run_steps(k, flush=True, events=None):
        """k steps through GraphShardedGRevNet.log_prob_async; the collective of step i is joined after step i+1's
        kernels are enqueued (the last one before returning, i.e. inside the timed region)."""
        pending, handles = None, []
        for i in range(k):
            if flush:
                flush_buf.fill_(i & 0xFF)                         # flush L2 (256 MB > 126 MB)
            if events:
                events[0][i].record()
            cur = sharded.log_prob_async(graph)
            if pending is not None:
                pending.wait()
            pending = cur
            handles.append(cur)
            if i == k - 1:
                pending.wait()                                    # every all-reduce is inside the timed region
            if events:
                events[1][i].record()
        return handles

    import ctypes
    # NVML queries hold the driver for milliseconds now and then: harmless when the device has steps queued (default
    # workload, 3.7 ms steps), visible when the host is barely ahead (protein B=256, 0.57 ms steps) -> sample those less often
    sampler = ClockSampler(local_rank, period=0.02 if args.workload == "community_medium" else 0.25)
    sampler.start()
    lib.gnf_debug_kernel_timing(1)                                # device timestamps per fused launch (no host events)
    run_steps(warmup)
    torch.cuda.synchronize()
    net.check_numerics()
    ev = ([torch.cuda.Event(enable_timing=True) for _ in range(args.steps)],
          [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)])
    k_total, k_count = ctypes.c_double(0.0), ctypes.c_int64(0)
    lib.gnf_debug_kernel_timing(1)                                # reset the slots: only the timed launches are kept
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib.gnf_launch_count(1)
    sampler.arm()
    if args.profile:
        torch.cuda.profiler.start()
    handles = run_steps(args.steps, events=ev)
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.stop()
    lib.gnf_debug_kernel_time(ctypes.byref(k_total), ctypes.byref(k_count))
    lib.gnf_debug_kernel_timing(0)
    launches = int(lib.gnf_launch_count(1))
    if world > 1:
        dist.barrier()
    ms = [s.elapsed_time(e) for s, e in zip(*ev)]
    my_ms = float(sum(ms)) / args.steps
    ms_sorted = sorted(ms)
    step_stats = {"min": ms_sorted[0], "median": ms_sorted[len(ms) // 2], "max": ms_sorted[-1], "first": ms[0],
                  "slowest_steps": sorted(range(len(ms)), key=lambda i: -ms[i])[:3]}
    ar_ms = float(np.mean([h.all_reduce_ms() for h in handles]))
    stats = torch.tensor([my_ms, float(n_nodes), float(n_edges), ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        allst = torch.stack(allst).cpu().numpy()
    else:
        allst = stats.cpu().numpy()[None]
    ms_per_step = float(allst[:, 0].max())                         # max over ranks
    value = node_updates_global / (ms_per_step * 1e-3)
    scal = handles[-1].wait()
    log_prob_xs = float(scal["log_prob_xs"].item())
    net.check_numerics()

    # ---- dominant kernel: one fused half-coupling launch (k_coupling_tc) ----------------------
    roof = None
    if math != "fp32" and k_count.value > 0:
        # the kernel's duration INSIDE the headline loop: every launch wrote {first CTA past its dependency wait, last
        # CTA done} from the device's %globaltimer into its own slot (gnf_debug_kernel_timing) -- no host events between
        # launches, programmatic dependent launch overlaps as in production, so 2T x ms_per_launch <= ms_per_step
        k_ms = k_total.value / k_count.value
        per_step = k_count.value / args.steps                      # 2T launches per step, or 1 (persistent launch)
        share = (per_step * k_ms) / my_ms
        flops = n_nodes * F * (2 * T / per_step)                   # a persistent launch runs all 2T half steps
        ach = flops / (k_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        ncu_file = os.path.join(ROOT, "profiles", "r2_ncu_full_k_coupling_tc.csv")
        if not os.path.exists(ncu_file):
            ncu_file = os.path.join(ROOT, "profiles", "r1_ncu_full_k_coupling_tc.csv")
        roof = {"bound": "tensor", "kernel": "k_coupling_tc (one fused half coupling step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": ncu_traffic(ncu_file) if (math == "tc3x" and args.workload == "community_medium") else None,
                "traffic_note": f"DRAM bytes per launch (ncu --set full, profiles/{os.path.basename(ncu_file)}); the "
                                "kernel is tensor-bound and lives in L2/smem/TMEM",
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                "ms_per_launch": k_ms,
                "ms_per_launch_how": "mean over the launches of the HEADLINE loop of (last CTA done - first CTA past "
                                     "griddepcontrol.wait), device %globaltimer written by the kernel itself",
                "launches_timed": int(k_count.value), "launches_per_step": per_step,
                "algorithmic_flops_per_launch": flops,
                "executed_mma_flops_per_algorithmic_flop": {"tc3x": 3, "tc3x_bf16": 3, "tc2x": 2}.get(math, 1),
                "share_of_step": share}

    # ---- scatter-reduce sub-op against the HBM roofline (standalone gather+segment-sum) --------
    seg = None
    if rank == 0 and not args.no_seg and args.workload == "community_medium":
        from graph_normalizing_flows_b200.graphs import concat_structures
        structs, feats = make_block(WORKLOADS["community_medium"], args.seg_graphs, SEED)
        big = concat_structures(structs, nodes=feats)
        nb, eb, h = int(big.nodes.shape[0]), int(len(big.senders)), D // 2
        gb = big.replace(nodes=np.ascontiguousarray(big.nodes[:, :h])).to(dev)
        torch.cuda.synchronize()
        for _ in range(2):                                        # second build: allocator and module loading warm
            t0 = time.perf_counter()
            stb = G.graphs.BatchStructure(gb.senders, gb.receivers, nb)
            torch.cuda.synchronize()
            csr_ms = (time.perf_counter() - t0) * 1e3
        outb = torch.empty_like(gb.nodes)
        stream = _lib.stream_ptr(dev)

        def seg_once():
            _lib.check(lib.gnf_gather_segment_sum(_lib.ptr(gb.nodes), h, _lib.ptr(stb.rowptr), _lib.ptr(stb.csr_senders),
                                                  nb, 0, _lib.ptr(outb), stream))
        for _ in range(3):
            seg_once()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        reps = 10
        for i in range(reps):
            flush_buf.fill_(i)
            e0.record()
            seg_once()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        seg_ms = tot / reps
        bytes_alg = eb * (8 + 4 * h) + 4 * nb * h                  # SURVEY §8d "THE figure"
        bytes_csr = eb * 4 + (nb + 1) * 4 + 2 * 4 * nb * h         # what the CSR form must move at least once
        ach = bytes_alg / (seg_ms * 1e-3) / 1e9
        traffic = ncu_traffic(os.path.join(ROOT, "profiles", "r2_ncu_full_k_gather_segment.csv"))
        seg = {"bound": "hbm", "kernel": "k_gather_segment (gather + segment-sum, standalone)",
               "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
               "traffic": traffic, "ms_per_launch": seg_ms, "algorithmic_bytes_per_launch": bytes_alg,
               "compulsory_bytes_csr_form": bytes_csr,
               "achieved_on_compulsory_bytes": bytes_csr / (seg_ms * 1e-3) / 1e9,
               "frac_on_compulsory_bytes": bytes_csr / (seg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
               "achieved_on_measured_dram_bytes": (traffic / (seg_ms * 1e-3) / 1e9) if traffic else None,
               "n_nodes": nb, "n_edges": eb, "csr_build_ms": csr_ms,
               "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})"}
        del gb, outb, stb

    # ---- row f2: reversible backward / training-step evaluation (forward + gradients) -----------
    train = None
    if math != "fp32" and not args.no_train and (wl.get("train") or (rank == 0 and world == 1)):
        reps = 5
        if wl.get("train") and world > 1:
            # sharded training-step evaluation: density pass, 4-vector all-reduce, reversible backward, gradient
            # all-reduce (NCCL) -- GraphShardedGRevNet.loss_and_grad
            for _ in range(2):
                sharded.loss_and_grad(graph, per_node=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(reps):
                out_t, grads = sharded.loss_and_grad(graph, per_node=True)
            e1.record()
            torch.cuda.synchronize()
            step_ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
            # the gradient all-reduce on its own (19.5 MB at T=6), ranks aligned by a barrier
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            g0.record()
            for _ in range(reps):
                dist.all_reduce(grads, op=dist.ReduceOp.SUM)
            g1.record()
            torch.cuda.synchronize()
            train = {"sharded_loss_and_grad_ms": float(step_ms.item()), "grad_all_reduce_ms": g0.elapsed_time(g1) / reps,
                     "grad_bytes": int(grads.numel() * 4), "loss_per_node": float(out_t["loss_per_node"].item()),
                     "node_updates_per_s_fwd_bwd": node_updates_global / (float(step_ms.item()) * 1e-3), "math": math,
                     "what": "GraphShardedGRevNet.loss_and_grad: f + 4-vector all-reduce + reversible backward + gradient "
                             "all-reduce(SUM) over NCCL, max over ranks, mean of %d" % reps}
        elif rank == 0:
            grads = torch.zeros_like(net.params.detach())
            z, _ = net.f64(graph)
            for _ in range(2):
                net.backward_from_z(graph, z.nodes, 1.0 / n_nodes, grads=grads)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                net.backward_from_z(graph, z.nodes, 1.0 / n_nodes, grads=grads)
            e1.record()
            torch.cuda.synchronize()
            bwd_ms = e0.elapsed_time(e1) / reps
            # executed-algorithm FLOPs of the backward: recompute (1x) + dX chain (1x) + dW (1x) of the forward's
            train = {"backward_ms": bwd_ms, "forward_ms": ms_per_step, "math": math,
                     "node_updates_per_s_fwd_bwd": n_nodes * T * 2 / ((bwd_ms + ms_per_step) * 1e-3),
                     "backward_algorithmic_tflops": 3 * n_nodes * 2 * T * F / (bwd_ms * 1e-3) / 1e12,
                     "what": "gnf_grevnet_backward (reversible: recompute + dX chain + dW GEMM per half step), "
                             "device-resident, CUDA events, mean of %d" % reps}
            del grads

    # ---- end to end through the public API from pinned host memory ---------------------------
    pinned = G.GraphsTuple(*[torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if v is not None else None
                             for v in host])
    h2d = sum(v.numel() * v.element_size() for v in pinned if v is not None)
    prefetch = G.graphs.BatchPrefetcher(dev)

    def e2e_loop(k):
        """k steps; every step's shard is copied from pinned host memory (H2D of nodes, senders, receivers,
        n_node, n_edge), validated and CSR-indexed inside the timed region -- staged one step ahead on a
        side stream, as a training input pipeline does -- and every step's 4 scalars are copied back to pinned host
        memory and READ by the host (one step later, so the device always has the next step queued; all k results
        have been read when the loop returns)."""
        ticket = prefetch.submit(pinned)
        host_vec = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)]
        inflight, last = None, None
        for i in range(k):
            g = prefetch.wait(ticket)
            pend = sharded.log_prob_async(g)                      # enqueue this step's kernels + all-reduce first ...
            if i + 1 < k:
                ticket = prefetch.submit(pinned)                  # ... then stage the next batch while they run
            pend.wait()
            host_vec[i & 1].copy_(pend.vec, non_blocking=True)    # D2H of this step's 4 scalars
            ev_done = torch.cuda.Event()
            ev_done.record()
            if inflight is not None:                              # host reads step i-1 while step i runs
                inflight[0].synchronize()
                last = inflight[1].clone()
            inflight = (ev_done, host_vec[i & 1])
        inflight[0].synchronize()
        return inflight[1].clone()

    e2e_loop(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    vec = e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_sec = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
    e2e_value = node_updates_global / float(e2e_sec.item())
    clocks = sampler.stop()
    net.check_numerics()

    # ---- CPU baseline + parity on the exact timed batch (rank 0, N=1 only) ---------------------
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(wl, host, 2, 1, keep_outputs=True)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        z_dev, ldj_dev = net.f64(graph)
        zg = z_dev.nodes.cpu().numpy()
        scale = float(np.abs(r["z"]).max())
        parity = {"oracle": "oracle/gnf_oracle_torch.py (torch-CPU fp32 restatement of gnn.py:304-341 + run_grevnet.py:292-296) "
                            "on the exact timed batch and weights",
                  "log_prob_xs_b200": log_prob_xs, "log_prob_xs_cpu": r["log_prob_xs"],
                  "log_prob_rel": abs(log_prob_xs - r["log_prob_xs"]) / abs(r["log_prob_xs"]),
                  "z_max_abs": float(np.abs(zg - r["z"]).max()), "z_max_abs_over_max_abs_z": float(np.abs(zg - r["z"]).max()) / scale,
                  "ldj_abs": abs(float(ldj_dev.item()) - r["ldj_f64_sums"]),
                  "ldj_rel": abs(float(ldj_dev.item()) - r["ldj_f64_sums"]) / abs(r["ldj_f64_sums"]),
                  "ldj_note": "against the CPU arm's own s with its reduce_sums accumulated in float64; the CPU arm's fp32 "
                              "reduce_sum (what TF does) is itself off by ldj_cpu_fp32_sum_rel",
                  "ldj_cpu_fp32_sum_rel": abs(r["ldj"] - r["ldj_f64_sums"]) / abs(r["ldj_f64_sums"]),
                  "vs_fp64": None,
                  "tolerance": "1e-5 relative on log-prob (north star)" if math != "bf16" else
                               "bf16 single pass: 1e-3 relative on log-prob (stated in tests/test_gpu_parity.py::test_config2_grid_12_step_bf16)"}

    if parity is not None and args.workload in ("community_medium", "protein_b256"):
        t64 = cpu_fp64_truth(wl, host)                               # ~10 s at B=4096; untimed
        rel = lambda a, b: abs(a - b) / abs(b)
        parity["vs_fp64"] = {
            "what": "both fp32 arms against ONE float64 pass of the restatement on the same batch (the common yardstick)",
            "log_prob_rel_b200": rel(log_prob_xs, t64["log_prob_xs"]), "log_prob_rel_cpu_fp32": rel(r["log_prob_xs"], t64["log_prob_xs"]),
            "ldj_rel_b200": rel(float(ldj_dev.item()), t64["ldj"]), "ldj_rel_cpu_fp32": rel(r["ldj"], t64["ldj"]),
            "z_max_abs_b200": float(np.abs(zg - t64["z"]).max()), "z_max_abs_cpu_fp32": float(np.abs(r["z"] - t64["z"]).max())}
    if rank == 0:
        config = base_config(wl, args.workload, world, gpg, int(allst[0, 1]), int(allst[0, 2]))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"],
                "vs_baseline": None,
                "dtype": {"tc3x": "f16x2-split/f32-acc", "tc3x_bf16": "bf16x2-split/f32-acc",
                          "tc2x": "f16 act x f16x2-split weights/f32-acc", "bf16": "bf16", "fp32": "f32"}[math],
                "data": "synthetic", "config": config, "math": math, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 32,
                        "returns": "the 4 log-prob scalars (fp64) per step; z stays on the device (sharded), as the "
                                   "reference's sess.run fetches scalars"},
                "gpu_launches": launches, "roofline": roof, "roofline_segment_sum": seg, "cpu_baseline": cpu,
                "parity": parity, "train_step": train, "log_prob_xs": log_prob_xs, "ms_per_step_stats_rank0": step_stats,
                "per_rank": {"ms_per_step": [float(v) for v in allst[:, 0]], "n_nodes": [int(v) for v in allst[:, 1]],
                             "n_edges": [int(v) for v in allst[:, 2]], "all_reduce_ms": [float(v) for v in allst[:, 3]],
                             "n_nodes_global": n_global,
                             "collective": ("nvlink peer memory (own one-warp kernel, cudaIpc)" if sharded.peer is not None
                                            else "nccl" if world > 1 else "none"),
                             "note": "all_reduce_ms = device time from reaching the all-reduce on the side stream to its "
                                     "completion (includes waiting for the slowest rank)"},
                "effective_tflops": value * F / 1e12}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
