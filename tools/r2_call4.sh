#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "inject or f1_dm" 2>&1 | tail -25 > gpurun_out/r2_pytest_inject.log; cat gpurun_out/r2_pytest_inject.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/bench_next_rows.py 2>&1 | tail -4 | tee gpurun_out/r2_next_rows.jsonl
timeout 600 python tools/bench_next_rows.py fp32 skip 2>&1 | tail -2 | tee -a gpurun_out/r2_next_rows.jsonl
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -2 gpurun_out/ncu_f1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_b.err; cat gpurun_out/r2_bench_n1.json | cut -c1-600; tail -3 gpurun_out/r2_b.err
