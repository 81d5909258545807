#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
SAN_TOOLS="racecheck synccheck" bash tools/gpu_sanitize.sh 2>&1 | grep -E "SUMMARY|error" | head
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-seg > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1_b.json; tail -3 gpurun_out/r2_b.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-seg > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1_c.json
timeout 300 python tools/ab_persist.py community_medium 30 1 2>&1 | tail -4 | tee gpurun_out/r2_ab_inproc3.jsonl
timeout 600 python bench.py --workload protein_b256 --steps 100 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r2_bench_protein_b256_n1.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_protein_b256_n1.json
