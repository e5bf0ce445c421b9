#!/bin/bash
# end-of-round records at N = 1: the default bench line (as the driver runs it), smoke(), every workload once, the launch list
TAG=${1:-r2q}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}_city16k_n1.json 2> gpurun_out/bench_${TAG}_city16k_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${TAG}_city16k_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_workloads.sh $TAG
bash tools/gpu_launch_list.sh > gpurun_out/ll_${TAG}.log 2>&1; tail -3 gpurun_out/launches_city16k_summary.md
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}_city16k_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e_ms", round(d["e2e"]["seconds_per_step"] * 1e3, 1), "parity", d["parity"]["ok"], "roofline", d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
print({k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
PY
