#!/bin/bash
# 2-GPU validation of the sharded bench (one process per GPU, NCCL weight broadcast).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
cat gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_${N}gpu.json
