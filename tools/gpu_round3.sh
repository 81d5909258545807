#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --math tc2x --no-cpu-baseline --seg-graphs 2048 > gpurun_out/bench_tc2x.json 2> gpurun_out/bench_tc2x.err; cat gpurun_out/bench_tc2x.json; tail -2 gpurun_out/bench_tc2x.err
python tools/accuracy_report.py > gpurun_out/accuracy.log 2>&1; tail -12 gpurun_out/accuracy.log
