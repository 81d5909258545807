#!/bin/bash
# k_gemm_tc with 128-row items for narrow layers: unit tests, embedding-flow timing + launch list, ncu --set full (2048x2048 layer)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "layered_tensor_core_linear or embedding_flow or golden_parity or warp_kernel" -o timeout=60 > gpurun_out/r2_pytest_call27a.log 2>&1; rc=$?
tail -4 gpurun_out/r2_pytest_call27a.log
if [ $rc -ne 0 ]; then echo "unit tests failed (rc=$rc)"; exit 1; fi
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_embedding_flow.csv python tools/bench_embedding_flow.py profile > gpurun_out/ncu_emb.log 2>&1; tail -1 gpurun_out/ncu_emb.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 2 -c 1 -f -o gpurun_out/ncu_full_k_gemm_tc python tools/bench_gemm_tc.py 6873 tc3x > gpurun_out/ncu_gemm_full.log 2>&1; tail -1 gpurun_out/ncu_gemm_full.log
ncu -i gpurun_out/ncu_full_k_gemm_tc.ncu-rep --page raw --csv > gpurun_out/ncu_full_k_gemm_tc_raw.csv 2>/dev/null; wc -c gpurun_out/ncu_full_k_gemm_tc_raw.csv
