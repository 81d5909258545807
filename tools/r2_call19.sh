#!/bin/bash
# k_gemm_tc bring-up: unit test first (short timeout), then the flows that now route through it, then the whole suite
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "layered_tensor_core_linear" -o timeout=60 > gpurun_out/r2_pytest_call19a.log 2>&1; rc=$?
tail -25 gpurun_out/r2_pytest_call19a.log
if [ $rc -ne 0 ]; then echo "unit test failed (rc=$rc)"; exit 1; fi
timeout 300 python -m pytest tests -q -m gpu -x -k "embedding_flow or f1_dm or wide" -o timeout=100 > gpurun_out/r2_pytest_call19b.log 2>&1; rc=$?
tail -25 gpurun_out/r2_pytest_call19b.log
if [ $rc -ne 0 ]; then echo "flow tests failed (rc=$rc)"; exit 1; fi
timeout 600 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_call19.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_call19.log | head -40
