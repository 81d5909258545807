"""ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel summary (kernel, launches, total, average, share)."""
import csv, sys
from collections import defaultdict
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
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
with open(dst, "w") as f:
    f.write("kernel,launches,total_us,avg_us,share_pct\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"\"{k[:110]}\",{cnt[k]},{v/1e3:.1f},{v/cnt[k]/1e3:.2f},{100*v/s:.2f}\n")
print(open(dst).read()[:1800], "total_us", s / 1e3)
