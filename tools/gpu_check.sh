#!/bin/bash
# one GPU call: the whole GPU suite, then A/B of the mixed-scene fusion on the composite crop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^\[vx-stats\]" | tail -8 | tee gpurun_out/pytest_gpu.log
AB_GOLD=size_composite_crop4k.json AB_REPS=5 timeout 300 python tools/gpu_ab_inproc.py "mixed_off:SVB_SLOW_LEAVES_MIXED=0" "mixed_on:" "mixed_off2:SVB_SLOW_LEAVES_MIXED=0" "mixed_on2:" 2>/dev/null | tee gpurun_out/ab_mixed.log
