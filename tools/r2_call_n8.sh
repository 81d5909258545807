#!/bin/bash
# 8 GPUs, one box: scaling table of the default workload (peer-memory collective, NCCL A/B at N=8), BASELINE configs[3]
# (protein B=256 over 4 GPUs) and configs[4] (citeseer + mixed over 8 GPUs with the gradient all-reduce)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -12 > gpurun_out/r2_topo.txt
run() { # tag, nproc, args, env
  env $4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --warmup 5 $3 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err
  cut -c1-260 gpurun_out/r2_bench_$1.json; tail -2 gpurun_out/r2_bench_$1.err
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-seg --no-cpu-baseline > gpurun_out/r2_bench_default_n1_box8.json 2>/dev/null; cut -c1-260 gpurun_out/r2_bench_default_n1_box8.json
run default_n8 8 "--steps 20 --no-train --no-seg" "A=1"
run default_n8_nccl 8 "--steps 20 --no-train --no-seg" "GNF_NO_PEER=1"
run default_n4 4 "--steps 20 --no-train --no-seg" "A=1"
run default_n2 2 "--steps 20 --no-train --no-seg" "A=1"
run protein_n4 4 "--workload protein_b256 --steps 200 --no-train" "A=1"
run protein_n4_nosampler 4 "--workload protein_b256 --steps 200 --no-train" "GNF_BENCH_NO_SAMPLER=1"
run protein_n4_nccl 4 "--workload protein_b256 --steps 200 --no-train" "GNF_NO_PEER=1"
run mixed_n8 8 "--workload mixed --steps 20" "A=1"
run citeseer_n8 8 "--workload citeseer --steps 50" "A=1"
run grid_n8 8 "--workload grid_t12_bf16 --steps 20 --no-train" "A=1"
ls gpurun_out | wc -l
