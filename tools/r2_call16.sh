#!/bin/bash
# after the chain-boundary fix: the new paths first (short timeouts), then the whole suite, timing, launch list
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x -k "inject or wide_graphs or f1_dm" -o timeout=120 > gpurun_out/r2_pytest_call16a.log 2>&1; rc=$?
tail -15 gpurun_out/r2_pytest_call16a.log
if [ $rc -ne 0 ]; then echo "new-path tests failed (rc=$rc): stopping here"; exit 1; fi
timeout 300 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1_b.jsonl
GNF_ATTN_MINB=2 timeout 300 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1_minb2.jsonl
timeout 600 python -m pytest tests -q -m gpu -o timeout=120 2>&1 | tail -15 > gpurun_out/r2_pytest_call16.log; tail -15 gpurun_out/r2_pytest_call16.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1_b.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -2 gpurun_out/ncu_f1.log
SAN_TOOLS="memcheck" timeout 500 bash tools/gpu_sanitize.sh 2>&1 | grep -E "SUMMARY|rror" | head
