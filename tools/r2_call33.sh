#!/bin/bash
# ncu --set full of the final k_gemm_tc (TMA path, 2048x2048 layer) and of k_dm_attn_warp; embedding-flow line with its backward
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 2 -c 1 -f -o gpurun_out/ncu_full_k_gemm_tc python tools/bench_gemm_tc.py 6873 tc3x > gpurun_out/ncu_gemm_full.log 2>&1; tail -1 gpurun_out/ncu_gemm_full.log
ncu -i gpurun_out/ncu_full_k_gemm_tc.ncu-rep --page raw --csv > gpurun_out/ncu_full_k_gemm_tc_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_dm_attn_warp -c 1 -f -o gpurun_out/ncu_full_k_dm_attn_warp python tools/bench_embedding_flow.py profile > gpurun_out/ncu_attn_full.log 2>&1; tail -1 gpurun_out/ncu_attn_full.log
ncu -i gpurun_out/ncu_full_k_dm_attn_warp.ncu-rep --page raw --csv > gpurun_out/ncu_full_k_dm_attn_warp_raw.csv 2>/dev/null
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
rm -f gpurun_out/*.ncu-rep
