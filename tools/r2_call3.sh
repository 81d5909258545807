#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ab_persist.py community_medium 30 0 2>&1 | tail -4 | tee gpurun_out/r2_ab_inproc.jsonl
timeout 300 python tools/ab_persist.py community_medium 30 1 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc.jsonl
timeout 300 python tools/ab_persist.py protein_b256 100 0 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc.jsonl
timeout 300 python tools/ab_persist.py protein_b256 100 1 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc.jsonl
timeout 300 python tools/ab_persist.py grid_t12_bf16 30 0 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc.jsonl
timeout 300 python tools/ab_persist.py mixed 30 0 2>&1 | tail -4 | tee -a gpurun_out/r2_ab_inproc.jsonl
GNF_PERSIST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_n1_p0.json 2> gpurun_out/r2_b.err; cat gpurun_out/r2_bench_n1_p0.json; tail -3 gpurun_out/r2_b.err
GNF_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r2_bench_n1_p1.json 2> gpurun_out/r2_b.err; cat gpurun_out/r2_bench_n1_p1.json; tail -3 gpurun_out/r2_b.err
timeout 600 python bench.py --workload protein_b256 --steps 50 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r2_bench_protein_auto.json 2> gpurun_out/r2_b.err; cat gpurun_out/r2_bench_protein_auto.json; tail -3 gpurun_out/r2_b.err
ls gpurun_out | head -80
