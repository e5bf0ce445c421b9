#!/bin/bash
TAG=${1:-r2i}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^\[vx-stats\]" > gpurun_out/pytest_gpu_${TAG}.log
tail -25 gpurun_out/pytest_gpu_${TAG}.log
AB_REPS=3 python tools/gpu_ab_inproc.py "default:" 2>&1 | tail -2
AB_GOLD=size_terrain4k.json AB_REPS=5 python tools/gpu_ab_inproc.py "default:" 2>&1 | tail -2
AB_GOLD=size_spongeball1k.json AB_REPS=5 python tools/gpu_ab_inproc.py "default:" 2>&1 | tail -2
AB_GOLD=size_composite_crop4k.json AB_REPS=3 python tools/gpu_ab_inproc.py "default:" 2>&1 | tail -2
