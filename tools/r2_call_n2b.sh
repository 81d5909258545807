#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -8
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -s -m gpu -k "nccl" 2>&1 | tail -15 > gpurun_out/r2_pytest_nccl2.log; cat gpurun_out/r2_pytest_nccl2.log
run() { # name, extra args, env
  env $3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $2 > gpurun_out/r2_bench_$1_n2.json 2> gpurun_out/r2_bench_$1_n2.err
  cat gpurun_out/r2_bench_$1_n2.json | cut -c1-300; tail -3 gpurun_out/r2_bench_$1_n2.err
}
run default_peer "--no-train" "GNF_X=1"
run default_nccl "--no-train" "GNF_NO_PEER=1"
run default_peer2 "--no-train" "GNF_X=1"
run default_nccl2 "--no-train" "GNF_NO_PEER=1"
run protein_b256 "--workload protein_b256 --steps 50" "GNF_X=1"
run mixed "--workload mixed" "GNF_X=1"
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r2_bench_n1_same_box.json 2>/dev/null; cat gpurun_out/r2_bench_n1_same_box.json | cut -c1-300
timeout 600 python tools/bench_next_rows.py tc3x skip 2>&1 | tail -1 | tee gpurun_out/r2_next_rows_f1.jsonl
