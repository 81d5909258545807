"""cuobjdump -sass of the in-tree libgnf_b200.so -> per-kernel instruction histogram (profiles/<tag>_sass_histogram.csv).
Runs without a GPU.  The columns that prove the tcgen05 / TMEM / bulk-copy path: UTCHMMA (tcgen05.mma), UTCBAR
(tcgen05.commit), LDTM / STTM (tcgen05.ld / .st), UBLKCP (cp.async.bulk), UTMALDG (cp.async.bulk.tensor through a
tensor map), SYNCS (mbarrier)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = os.path.join(ROOT, "graph_normalizing_flows_b200", "libgnf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "FFMA", "FADD",
        "DADD", "MUFU", "ATOMG", "RED", "SHFL", "BAR", "ELECT"]
rows, cur, hist = [], None, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, hist))
        cur, hist = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[m.group(1).split(".")[0]] += 1
        hist["_total"] += 1
if cur:
    rows.append((cur, hist))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
out = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.csv")
with open(out, "w") as f:
    f.write("kernel,instructions," + ",".join(COLS) + "\n")
    for (name, h), d in sorted(zip(rows, dem), key=lambda t: -t[0][1]["_total"]):
        short = re.sub(r"gnf::\(anonymous namespace\)::", "", d)
        short = re.sub(r"\(.*", "", short)[:90]
        f.write(f"\"{short}\",{h['_total']}," + ",".join(str(h[c]) for c in COLS) + "\n")
tot = collections.Counter()
for _, h in rows:
    tot.update(h)
print(out, "kernels:", len(rows), {c: tot[c] for c in COLS[:7]})
