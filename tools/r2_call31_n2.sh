#!/bin/bash
# 2 GPUs on the final tree: NCCL world-2 test of GraphShardedGRevNet (incl. gradient and batch-norm all-reduces), default bench at N=2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "nccl" -o timeout=200 2>&1 | tail -5 > gpurun_out/r2_pytest_nccl2.log; cat gpurun_out/r2_pytest_nccl2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_default_n2_final.json 2> gpurun_out/r2_bench_default_n2_final.err
cut -c1-300 gpurun_out/r2_bench_default_n2_final.json; tail -2 gpurun_out/r2_bench_default_n2_final.err
