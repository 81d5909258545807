#!/bin/bash
# round 2, first box visit: parity suite, bench (all workloads at N=1), reference arm, ncu of the new segment kernel,
# launch list, power trace
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
for w in grid_t12_bf16 protein_b256 citeseer mixed; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err; cat gpurun_out/r2_bench_${w}_n1.json; tail -3 gpurun_out/r2_bench_${w}_n1.err
done
timeout 600 python bench.py --workload grid_t12_bf16 --graphs-per-gpu 8192 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_grid_b8192_n1.json 2> gpurun_out/r2_bench_grid_b8192_n1.err; cat gpurun_out/r2_bench_grid_b8192_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gather_segment -c 2 \
    -f -o gpurun_out/prof_seg python tools/seg_only.py > gpurun_out/ncu_seg.log 2>&1; tail -3 gpurun_out/ncu_seg.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --seg-graphs 2048 --profile > gpurun_out/ncu_launch.log 2>&1
timeout 300 python tools/power_trace.py 2>&1 | tail -5
timeout 300 python tools/bench_small.py 2>&1 | tail -4 | tee gpurun_out/r2_bench_small.jsonl
ls -la gpurun_out
