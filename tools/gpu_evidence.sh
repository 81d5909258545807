#!/bin/bash
# Evidence run of the round: parity suite, bench (+ reference arm), ncu launch lists (forward step and backward),
# ncu --set full of the three tensor-core kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --seg-graphs 2048 --profile > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_bwd.csv python tools/profile_bwd.py 4096 tc3x > gpurun_out/ncu_launch_bwd.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_coupling_tc -c 2 \
    -f -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --seg-graphs 2048 --profile > gpurun_out/ncu_tc.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_bwd_chain -c 1 \
    -f -o gpurun_out/prof_bwd_chain python tools/profile_bwd.py 4096 tc3x > gpurun_out/ncu_bwd_chain.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_dw_tc -c 1 \
    -f -o gpurun_out/prof_dw python tools/profile_bwd.py 4096 tc3x > gpurun_out/ncu_dw.log 2>&1
tail -2 gpurun_out/ncu_tc.log gpurun_out/ncu_bwd_chain.log gpurun_out/ncu_dw.log
ls -la gpurun_out
