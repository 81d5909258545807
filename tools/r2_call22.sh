#!/bin/bash
# where do k_gemm_tc's cycles go: per-launch times of the three layer shapes under the timing variants
mkdir -p gpurun_out
for v in 0 1 2 4 6; do
  GNF_GEMM_VARIANT=$v timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_gemm_tc --log-file gpurun_out/gemm_var$v.csv python tools/bench_gemm_tc.py 6873 tc3x,bf16 > gpurun_out/gemm_var$v.log 2>&1
  echo "variant $v"; grep k_gemm_tc gpurun_out/gemm_var$v.csv | awk -F'","' '{print $5, $NF}' | cut -c1-20,60-
done
