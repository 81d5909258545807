#!/bin/bash
# evidence on the final tree: whole GPU suite, default bench line, launch lists of the f1 and embedding-flow workloads
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu.log | head -20
timeout 200 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r2_bench_n1.json; cut -c1-400 gpurun_out/r2_bench_n1.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_f1_fwd_bwd.csv python tools/profile_f1.py > gpurun_out/ncu_f1.log 2>&1; tail -1 gpurun_out/ncu_f1.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_embedding_flow.csv python tools/bench_embedding_flow.py profile > gpurun_out/ncu_emb.log 2>&1; tail -1 gpurun_out/ncu_emb.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -5
