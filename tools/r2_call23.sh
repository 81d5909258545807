#!/bin/bash
# k_gemm_tc: staggered K order + per-warp arrivals; unit tests, then per-launch times under the timing variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "layered_tensor_core_linear or embedding_flow" -o timeout=60 > gpurun_out/r2_pytest_call23a.log 2>&1; rc=$?
tail -5 gpurun_out/r2_pytest_call23a.log
if [ $rc -ne 0 ]; then echo "unit tests failed (rc=$rc)"; exit 1; fi
for v in 0 32 8 16 2; do
  GNF_GEMM_VARIANT=$v timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_gemm_tc --log-file gpurun_out/gemm_var$v.csv python tools/bench_gemm_tc.py 6873 tc3x,bf16 > gpurun_out/gemm_var$v.log 2>&1
done
python - <<'PY'
import csv
for v in (0,32,8,16,2):
    rows=[r for r in csv.reader(open(f'gpurun_out/gemm_var{v}.csv')) if len(r)>10]
    hdr=rows[0]; ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
    vals=[(('3x' if '<3' in r[ki] else '1x'), float(r[vi].replace(',',''))/1e3) for r in rows[1:]]
    print(v, [f"{a}:{b:.0f}" for a,b in vals])
PY
