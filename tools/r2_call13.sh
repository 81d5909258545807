#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_b.err; cut -c1-400 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_b.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-seg > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_n1_b.json
timeout 600 python bench.py --workload protein_b256 --steps 100 --warmup 5 --no-train > gpurun_out/r2_bench_protein_b256_n1.json 2> gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_bench_protein_b256_n1.json
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
