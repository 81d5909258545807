#!/bin/bash
# warp-per-output fold after k_dw_skinny: attention backward tests, f1 timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "f1_dm or wide_graphs or warp_kernel or run_grevnet_port or embedding_flow" -o timeout=100 > gpurun_out/r2_pytest_call30a.log 2>&1; echo rc=$?
tail -3 gpurun_out/r2_pytest_call30a.log
timeout 200 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1.jsonl
