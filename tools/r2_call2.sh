#!/bin/bash
# round 2, second box visit: persistent launch bring-up + A/B, new segment kernel, full parity suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "persistent" 2>&1 | tail -15 > gpurun_out/r2_pytest_persist.log; cat gpurun_out/r2_pytest_persist.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
for p in 0 1; do
  GNF_PERSIST=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2_ab_persist${p}_default.json 2> gpurun_out/r2_ab.err; cat gpurun_out/r2_ab_persist${p}_default.json; tail -3 gpurun_out/r2_ab.err
  GNF_PERSIST=$p timeout 600 python bench.py --workload protein_b256 --steps 50 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2_ab_persist${p}_protein.json 2> gpurun_out/r2_ab.err; cat gpurun_out/r2_ab_persist${p}_protein.json; tail -3 gpurun_out/r2_ab.err
  GNF_PERSIST=$p timeout 600 python bench.py --workload grid_t12_bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2_ab_persist${p}_grid.json 2> gpurun_out/r2_ab.err; cat gpurun_out/r2_ab_persist${p}_grid.json; tail -3 gpurun_out/r2_ab.err
  GNF_PERSIST=$p timeout 300 python tools/bench_small.py 2>&1 | tail -4 | tee gpurun_out/r2_bench_small_persist${p}.jsonl
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gather_segment -c 2 \
    -f -o gpurun_out/prof_seg python tools/seg_only.py > gpurun_out/ncu_seg.log 2>&1; tail -3 gpurun_out/ncu_seg.log
ls -la gpurun_out | head -50
