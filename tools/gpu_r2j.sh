#!/bin/bash
TAG=${1:-r2j}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_REPS=3 python tools/gpu_ab_inproc.py "fuse3:" "nofuse3:SVB_FUSE3=0" "fuse3b:" 2>&1 | tail -4
AB_GOLD=size_composite_crop4k.json AB_REPS=3 python tools/gpu_ab_inproc.py "fuse3:" "nofuse3:SVB_FUSE3=0" 2>&1 | tail -3
AB_GOLD=bigsize_city8k.json AB_REPS=3 python tools/gpu_ab_inproc.py "fuse3:" "nofuse3:SVB_FUSE3=0" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "^\[vx-stats\]" > gpurun_out/pytest_gpu_${TAG}.log
tail -15 gpurun_out/pytest_gpu_${TAG}.log
