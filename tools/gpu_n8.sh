#!/bin/bash
# N = 8 bench line with per-rank phase times; optional second run under an environment override (A/B on the same box)
TAG=${1:-n8}; ALT=$2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 > gpurun_out/bench_${TAG}_$1.json 2> gpurun_out/bench_${TAG}_$1.err
  echo "rc=$? $1 $2"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_${TAG}_$1.err | tail -5
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}_$1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ssvdag_sha256"][:16], d["parity"]["ok"])
for k, v in d["per_rank"].items(): print(" ", k, v)
PY
}
run default "SVB_X=0"
[ -n "$ALT" ] && run alt "$ALT"
