#!/bin/bash
# k_gemm_tc with the TMA tensor-map A path: unit tests on both paths, per-launch times of both, embedding-flow timing
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x -k "layered_tensor_core_linear or embedding_flow or golden_parity" -o timeout=60 > gpurun_out/r2_pytest_call28a.log 2>&1; rc=$?
tail -4 gpurun_out/r2_pytest_call28a.log
if [ $rc -ne 0 ]; then echo "TMA path failed (rc=$rc)"; grep -E "Error|error|assert" gpurun_out/r2_pytest_call28a.log | head -20; fi
GNF_GEMM_TMA=0 timeout 200 python -m pytest tests -q -m gpu -x -k "layered_tensor_core_linear" -o timeout=60 > gpurun_out/r2_pytest_call28b.log 2>&1; echo "LDG path rc=$?"; tail -2 gpurun_out/r2_pytest_call28b.log
for t in 1 0; do
  GNF_GEMM_TMA=$t timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_gemm_tc --log-file gpurun_out/gemm_tma$t.csv python tools/bench_gemm_tc.py 6873 tc3x > gpurun_out/gemm_tma$t.log 2>&1
  tail -3 gpurun_out/gemm_tma$t.log
done
python - <<'PY'
import csv
for v in (1,0):
    rows=[r for r in csv.reader(open(f'gpurun_out/gemm_tma{v}.csv')) if len(r)>10]
    hdr=rows[0]; ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
    print("tma",v,[f"{float(r[vi].replace(',',''))/1e3:.0f}" for r in rows[1:]])
PY
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
