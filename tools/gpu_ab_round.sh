#!/bin/bash
# one GPU call: in-process A/B of the environment toggles, then the small parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_ab_inproc.py "old:SVB_SLOW_LEAVES=0" "sl4:" "sl3:SVB_VX_OCC_SL=3" "sl5:SVB_VX_OCC_SL=5" "old2:SVB_SLOW_LEAVES=0" "sl4b:" 2> gpurun_out/ab_stats.err | tee gpurun_out/ab_inproc.log | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_parity.log
