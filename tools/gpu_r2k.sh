#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_REPS=3 python tools/gpu_ab_inproc.py "tma:" "notma:SVB_CHILDREN_TMA=0" "tma2:" 2>&1 | tail -4
AB_GOLD=size_terrain4k.json AB_REPS=5 python tools/gpu_ab_inproc.py "tma:" "notma:SVB_CHILDREN_TMA=0" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "^\[vx-stats\]" > gpurun_out/pytest_gpu_r2k.log
tail -12 gpurun_out/pytest_gpu_r2k.log
