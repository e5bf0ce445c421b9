#!/bin/bash
# one GPU call: the whole GPU suite, then memcheck of the chain-centre box-mesh build (k_slow_leaves<false,...>)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[vx-stats\]" | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_chain_city.py 2>&1 | grep -v vx-stats | tail -6 | tee gpurun_out/san_chain.log
