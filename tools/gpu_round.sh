#!/bin/bash
# One GPU-box visit: parity suite, bench line, ncu launch list, ncu full capture of the top kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --seg-graphs 2048 --profile > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_coupling_tc -c 2 \
    -f -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --seg-graphs 2048 --profile > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
