#!/bin/bash
# 2-GPU NCCL check of the bench, plus the other math modes on one GPU.
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
for m in bf16 tc3x_bf16 fp32; do
  python bench.py --steps 5 --warmup 3 --math $m --no-cpu-baseline --seg-graphs 2048 > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; cat gpurun_out/bench_$m.json; tail -2 gpurun_out/bench_$m.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
