"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the committed summaries under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        tot[r[ki]] += v
        cnt[r[ki]] += 1
    s = sum(tot.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,avg_us,share_pct\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"\"{k[:110]}\",{cnt[k]},{v/1e3:.1f},{v/cnt[k]/1e3:.2f},{100*v/s:.2f}\n")
    print(open(out).read())


def raw(rep, out, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(rows) - 2)) + "\n")
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k) for k in keys):
                f.write(f"{h},{units[i]}," + ",".join(r[i] for r in rows[2:]) + "\n")
    print(open(out).read())


KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor"]

if __name__ == "__main__":
    launches("gpurun_out/launches.csv", f"profiles/{tag}_launches_bench_tc3x.csv")
    raw("gpurun_out/prof_tc.ncu-rep", f"profiles/{tag}_ncu_full_k_coupling_tc.csv", KEYS)
