#!/bin/bash
# One box visit that regenerates the late-round-2 evidence under gpurun_out/ (copy what should be judged into profiles/):
#   whole GPU suite (poisoned workspaces), default bench line, f1 and embedding-flow numbers + launch lists,
#   k_gemm_tc per-launch times (TMA and LDG A paths), ncu --set full of k_gemm_tc, sanitizers over the new kernels.
# Needs ~6 GPU-minutes on one B200.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu -o timeout=100 > gpurun_out/r2_pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu.log | head -20
timeout 200 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r2_bench_n1.json; cut -c1-300 gpurun_out/r2_bench_n1.json
timeout 200 python tools/bench_next_rows.py tc3x 2>&1 | grep "^{" | tee gpurun_out/r2_next_rows.jsonl
timeout 200 python tools/bench_small.py 2>&1 | grep "^{" | tee gpurun_out/r2_bench_small.jsonl
timeout 300 python tools/bench_embedding_flow.py 2>&1 | tail -1 | tee gpurun_out/r2_embedding_flow.jsonl
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --profile-from-start off --log-file gpurun_out/launches_f1.csv python tools/profile_f1.py > /dev/null 2>&1
timeout 300 $NCU --profile-from-start off --log-file gpurun_out/launches_embedding.csv python tools/bench_embedding_flow.py profile > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_f1.csv gpurun_out/r2_launches_f1_fwd_bwd.csv | head -12
python tools/summarize_launches.py gpurun_out/launches_embedding.csv gpurun_out/r2_launches_embedding_flow.csv | head -8
for t in 1 0; do
  GNF_GEMM_TMA=$t timeout 120 $NCU -k regex:k_gemm_tc --log-file gpurun_out/gemm_tma$t.csv python tools/bench_gemm_tc.py 6873 tc3x > /dev/null 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 2 -c 1 -f -o gpurun_out/ncu_full_k_gemm_tc python tools/bench_gemm_tc.py 6873 tc3x > /dev/null 2>&1
ncu -i gpurun_out/ncu_full_k_gemm_tc.ncu-rep --page raw --csv > gpurun_out/ncu_full_k_gemm_tc_raw.csv 2>/dev/null; rm -f gpurun_out/*.ncu-rep
bash tools/gpu_sanitize_new.sh 2>&1 | grep -E "SUMMARY"
