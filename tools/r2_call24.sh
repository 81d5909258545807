#!/bin/bash
# skinny dW kernel + chained batch-norm entries: new tests, whole suite, small-batch + f1 timing
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "batch_norm or f1_dm or a9 or f4 or run_grevnet_port" -o timeout=100 > gpurun_out/r2_pytest_call24a.log 2>&1; rc=$?
tail -15 gpurun_out/r2_pytest_call24a.log
if [ $rc -ne 0 ]; then echo "new tests failed (rc=$rc)"; exit 1; fi
timeout 700 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_call24.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_call24.log | head -40
timeout 200 python tools/bench_small.py 2>&1 | grep "^{" | tee gpurun_out/r2_bench_small.jsonl
timeout 200 python tools/bench_next_rows.py tc3x 2>&1 | grep "^{" | tee gpurun_out/r2_next_rows.jsonl
