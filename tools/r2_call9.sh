#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
SAN_TOOLS="synccheck" bash tools/gpu_sanitize.sh 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_b.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-seg > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1_b.json
timeout 300 python tools/ab_persist.py community_medium 30 1 2>&1 | tail -4 | tee gpurun_out/r2_ab_inproc2.jsonl
timeout 300 python tools/ab_persist.py protein_b256 100 1 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc2.jsonl
timeout 300 python tools/ab_persist.py grid_t12_bf16 30 1 2>&1 | tail -2 | tee -a gpurun_out/r2_ab_inproc2.jsonl
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_coupling_tc -c 2 \
    -f -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-seg --profile > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-seg --profile > gpurun_out/ncu_launch.log 2>&1
