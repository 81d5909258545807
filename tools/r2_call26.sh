#!/bin/bash
# A/B of the forward attention kernel variants on the f1 workload (density pass only), then ncu --set full of k_gemm_tc
mkdir -p gpurun_out
for v in "GNF_ATTN_SPECIAL=1 GNF_ATTN_MINB=3" "GNF_ATTN_SPECIAL=1 GNF_ATTN_MINB=2" "GNF_ATTN_SPECIAL=0"; do
  echo "== $v"
  env $v timeout 200 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['density_pass_ms'], d['backward_ms'])"
done 2>&1 | tee gpurun_out/r2_ab_attention_variants.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 1 -o gpurun_out/ncu_full_k_gemm_tc python tools/bench_gemm_tc.py 6873 tc3x > gpurun_out/ncu_gemm_full.log 2>&1; tail -2 gpurun_out/ncu_gemm_full.log
ncu -i gpurun_out/ncu_full_k_gemm_tc.ncu-rep --page raw --csv > gpurun_out/ncu_full_k_gemm_tc_raw.csv 2>/dev/null; wc -c gpurun_out/ncu_full_k_gemm_tc_raw.csv
