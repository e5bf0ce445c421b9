#!/bin/bash
# Round 2, call D: full GPU suite with tracebacks kept, then bench at N = 1 (and N = 2 when two GPUs are visible)
TAG=${1:-r2d}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^\[vx-stats\]" > gpurun_out/pytest_gpu_${TAG}.log
tail -60 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -c 300 gpurun_out/bench_${TAG}_n1.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
  tail -c 1500 gpurun_out/bench_${TAG}_n2.err
fi
python - <<PY
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ssvdag_sha256"][:16], d["parity"]["ok"])
    except Exception as e:
        print(n, "failed", e)
PY
