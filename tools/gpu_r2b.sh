#!/bin/bash
# Round 2, call B (2 GPUs): GPU tests incl. the multi-GPU tool, bench at N=1 and N=2 (GPU-side encoders, stream-ordered exchange)
TAG=${1:-r2b}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -c 600 gpurun_out/bench_${TAG}_n1.err
timeout 300 python tools/e2e_breakdown.py city_16k 5 2>&1 | tail -1 | tee gpurun_out/e2e_breakdown_${TAG}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
tail -c 600 gpurun_out/bench_${TAG}_n2.err
python - <<PY
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ssvdag_sha256"][:16], d["parity"]["ok"])
    except Exception as e:
        print(n, "failed", e)
PY
