#!/bin/bash
# region config through bench.py at N = $1 GPUs (default 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --config region --steps 1 --warmup 1 > gpurun_out/region_n1.json 2> gpurun_out/region_n1.err; echo "rc=$?"
  tail -3 gpurun_out/region_n1.err; cat gpurun_out/region_n1.json
  timeout 600 python tools/bench_region.py --rows 24 --cols 40 --verify > gpurun_out/region_verify.json 2>&1; tail -1 gpurun_out/region_verify.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --config region --steps 1 --warmup 1 > gpurun_out/region_n$N.json 2> gpurun_out/region_n$N.err; echo "rc=$?"
  tail -3 gpurun_out/region_n$N.err; tail -1 gpurun_out/region_n$N.json
fi
