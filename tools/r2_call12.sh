#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-seg > gpurun_out/r2_bench_n1_c12.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1_c12.json; tail -3 gpurun_out/r2_b.err
timeout 600 python bench.py --math tc2x --steps 10 --warmup 3 --no-cpu-baseline --no-seg > gpurun_out/r2_bench_n1_tc2x.json 2> gpurun_out/r2_b.err; cut -c1-200 gpurun_out/r2_bench_n1_tc2x.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_bwd.csv python tools/profile_bwd.py 4096 tc3x > gpurun_out/ncu_launch_bwd.log 2>&1
timeout 300 python tools/bench_train.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_train.jsonl
timeout 300 python tools/bringup_bwd.py time 4096 2>&1 | tail -4 | tee gpurun_out/r2_bwd_time.log
