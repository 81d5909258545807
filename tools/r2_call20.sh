#!/bin/bash
# after routing every remaining shape through k_gemm_tc: whole suite, embedding-flow timing + launch list
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_call20.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_call20.log | head -40
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_embedding_flow.csv python tools/bench_embedding_flow.py profile > gpurun_out/ncu_emb.log 2>&1; tail -2 gpurun_out/ncu_emb.log
