#!/bin/bash
set -x
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh 2>&1 | tail -30
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
