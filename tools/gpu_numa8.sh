#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
lscpu | grep -i "numa\|socket\|model name" ; ls /sys/devices/system/node | head; free -g | head -2
for il in 1 0; do
STC_BENCH_INTERLEAVE=$il timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$il bench.py --gpus $N --steps 10 --warmup 3 --no-tile-chain > gpurun_out/bench_numa${il}_n$N.json 2> gpurun_out/bench_numa${il}_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_numa${il}_n$N.json").read().strip().splitlines()[-1])
print("interleave=$il", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_gbs_per_rank"), d["e2e"].get("pinned_pages"), "e2e_f32", d["e2e_f32"]["value"])
PY
done
