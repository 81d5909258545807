#!/bin/bash
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_gemm_tc --log-file gpurun_out/linear_small.csv python tools/bench_linear_small.py > gpurun_out/linear_small.log 2>&1
cat gpurun_out/linear_small.log | tail -5
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/linear_small.csv')) if len(r)>10]
hdr=rows[0]; vi=hdr.index("Metric Value")
print([f"{float(r[vi].replace(',',''))/1e3:.1f}" for r in rows[1:]])
PY
