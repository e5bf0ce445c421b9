#!/bin/bash
# reference hashes at 4096^3 / 8192^3 (tests/golden/{midsize_city4k,bigsize_city8k}.json): the in-process tool prints
# ref-ok / REF-MISMATCH.  gpurun -- tools/gpu_ref_check.sh [lots levels step]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_LOTS=${1:-64} AB_LEVELS=${2:-12} AB_STEP=${3:-3} AB_REPS=1 timeout 60 python tools/gpu_ab_inproc.py "default:" 2>&1 | tee gpurun_out/ref_check_${1:-64}.log | tail -3
