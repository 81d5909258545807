#!/bin/bash
# layered flows: forward recompute of the backward in k_gemm_tc / k_linear_tc -- whole suite, embedding-flow line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu.log | head -20
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
