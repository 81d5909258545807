#!/bin/bash
# full GPU check of the round: parity suite, sanitizers, default bench + reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash tools/gpu_sanitize.sh
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
