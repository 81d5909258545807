#!/bin/bash
# whole GPU suite with poisoned workspaces (no -x: enumerate every failure), then f1 timing, default bench, f1 launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_call18.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_call18.log | head -40
timeout 200 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1_b.jsonl
timeout 200 python bench.py 2>gpurun_out/bench_err.log | tail -1 | tee gpurun_out/r2_bench_n1.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1_fwd_bwd.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -2 gpurun_out/ncu_f1.log
