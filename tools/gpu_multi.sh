#!/bin/bash
# N-GPU validation of the sharded bench and of the region run (one process per GPU, NCCL).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
cut -c1-700 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
   tools/bench_region.py --verify > gpurun_out/region_${N}gpu.json 2> gpurun_out/region_${N}gpu.err; echo "region rc=$?"
cat gpurun_out/region_${N}gpu.json; tail -2 gpurun_out/region_${N}gpu.err
