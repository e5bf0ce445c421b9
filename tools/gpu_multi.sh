#!/bin/bash
# multi-GPU bench lines on one box: gpurun --gpus N -- tools/gpu_multi.sh N [workloads...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
WL=${@:-city_16k}
for W in $WL; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 2 --warmup 1 --workload $W > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err
  tail -n 3 gpurun_out/bench_${W}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${W}_n$N.json") if l.startswith("{")][-1])
    print("$W n=$N", {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "voxels", "nodes", "batches")}, "e2e", d["e2e"]["seconds_per_step"] if d.get("e2e") else None)
except Exception as e:
    print("no json for $W:", e)
PY
done
