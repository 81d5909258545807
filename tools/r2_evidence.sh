#!/bin/bash
# round 2 evidence run on one B200: parity suite, bench (all workloads at N=1), reference arm, ncu launch lists and
# --set full captures (fused kernel, gather+segment kernel, tensor-core linear, attention), small-batch / training /
# next-row measurements, accuracy per math mode
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_b.err
for w in grid_t12_bf16 protein_b256 citeseer mixed; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_b.err; cut -c1-200 gpurun_out/r2_bench_${w}_n1.json; tail -3 gpurun_out/r2_b.err
done
timeout 600 python bench.py --workload grid_t12_bf16 --graphs-per-gpu 8192 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_grid_b8192_n1.json 2> gpurun_out/r2_b.err; cut -c1-200 gpurun_out/r2_bench_grid_b8192_n1.json
for m in tc3x_bf16 tc2x bf16 fp32; do
  timeout 600 python bench.py --math $m --steps 10 --warmup 3 --no-cpu-baseline --no-seg > gpurun_out/r2_bench_n1_$m.json 2> gpurun_out/r2_b.err; cut -c1-200 gpurun_out/r2_bench_n1_$m.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_cpu.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_reference_cpu.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-seg --profile > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_coupling_tc -c 2 \
    -f -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-seg --profile > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gather_segment -c 2 \
    -f -o gpurun_out/prof_seg python tools/seg_only.py > gpurun_out/ncu_seg.log 2>&1; tail -2 gpurun_out/ncu_seg.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"k_linear_tc|k_dm_attn_block" -c 6 \
    -f -o gpurun_out/prof_f1 python tools/profile_f1.py > gpurun_out/ncu_f1_full.log 2>&1; tail -2 gpurun_out/ncu_f1_full.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_bwd.csv python tools/profile_bwd.py 4096 tc3x > gpurun_out/ncu_launch_bwd.log 2>&1
timeout 300 python tools/bench_small.py 2>&1 | tail -3 | tee gpurun_out/r2_bench_small.jsonl
timeout 300 python tools/bench_train.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_train.jsonl
timeout 600 python tools/bench_next_rows.py 2>&1 | tail -3 | tee gpurun_out/r2_next_rows.jsonl
timeout 300 python tools/accuracy_report.py 2>&1 | tail -6 | tee gpurun_out/r2_accuracy_modes.jsonl
timeout 300 python tools/power_trace.py 2>&1 | tail -3
ls gpurun_out | wc -l
