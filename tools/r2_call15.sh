#!/bin/bash
# f1 / inject backward on the tensor cores + block-staged attention backward: tests, sanitizers, timing, launch list
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_pytest_call15.log; tail -40 gpurun_out/r2_pytest_call15.log
SAN_TOOLS="memcheck racecheck" bash tools/gpu_sanitize.sh 2>&1 | grep -E "SUMMARY|error|Error" | head
timeout 600 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1_b.jsonl
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1_b.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -2 gpurun_out/ncu_f1.log
