#!/bin/bash
# multi-GPU checks: the svbuilder --devices test, the simulated-rank / NCCL tests, then bench at N = all visible GPUs (and N = 2)
TAG=${1:-m2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
[ "$2" = "benchonly" ] || timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "svbuilder_cli or sharded or merge_collision" 2>&1 | grep -v "^\[vx-stats\]" | tail -15 | tee gpurun_out/pytest_multi_${TAG}.log
for N in ${3:-2 $NG}; do
  [ "$N" -gt "$NG" ] && continue
  [ "$N" = "2" ] && [ "$NG" = "2" ] && [ -f gpurun_out/bench_${TAG}_n2.json ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
  echo "rc=$? N=$N"; tail -c 800 gpurun_out/bench_${TAG}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print($N, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ssvdag_sha256"][:16], d["parity"]["ok"])
    print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print($N, "failed", e)
PY
done
