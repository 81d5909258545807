#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "f1_dm or inject" 2>&1 | tail -6
timeout 600 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1.jsonl
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -2 gpurun_out/ncu_f1.log
