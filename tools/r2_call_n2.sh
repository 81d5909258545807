#!/bin/bash
# 2 GPUs: NCCL world_size-2 test of GraphShardedGRevNet, bench at N=2 (default weak scaling, protein strong scaling, mixed + grad all-reduce)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "nccl" 2>&1 | tail -15 > gpurun_out/r2_pytest_nccl2.log; cat gpurun_out/r2_pytest_nccl2.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $2 > gpurun_out/r2_bench_$1_n2.json 2> gpurun_out/r2_bench_$1_n2.err
  cat gpurun_out/r2_bench_$1_n2.json; tail -3 gpurun_out/r2_bench_$1_n2.err
}
run default ""
run protein_b256 "--workload protein_b256 --steps 50"
run mixed "--workload mixed"
run citeseer "--workload citeseer"
timeout 600 python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_n1_same_box.json 2>/dev/null; cat gpurun_out/r2_bench_n1_same_box.json | cut -c1-400
