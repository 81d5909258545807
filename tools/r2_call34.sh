#!/bin/bash
# warp-per-(node, head) attention backward kernels: tests, embedding-flow line (with its backward)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "warp_kernel or embedding_flow or f1_dm or wide_graphs" -o timeout=100 > gpurun_out/r2_pytest_call34a.log 2>&1; echo rc=$?
tail -3 gpurun_out/r2_pytest_call34a.log
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
