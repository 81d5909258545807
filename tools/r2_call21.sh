#!/bin/bash
# warp attention kernel + deeper prefetch in k_gemm_tc: new tests, embedding-flow timing + launch list, whole suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "warp_kernel or embedding_flow or layered_tensor" -o timeout=60 > gpurun_out/r2_pytest_call21a.log 2>&1; rc=$?
tail -25 gpurun_out/r2_pytest_call21a.log
if [ $rc -ne 0 ]; then echo "new tests failed (rc=$rc)"; exit 1; fi
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_embedding_flow.csv python tools/bench_embedding_flow.py profile > gpurun_out/ncu_emb.log 2>&1; tail -2 gpurun_out/ncu_emb.log
timeout 700 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_call21.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_call21.log | head -40
