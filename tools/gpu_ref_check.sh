#!/bin/bash
# reference hashes of a tests/golden/*size*.json (unmodified reference svbuilder): the in-process tool prints ref-ok / REF-MISMATCH.
#   gpurun -- tools/gpu_ref_check.sh midsize_city4k.json | bigsize_city8k.json | size_terrain4k.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_GOLD=${1:-midsize_city4k.json} AB_REPS=1 AB_WARMUP=${2:-1} timeout 60 python tools/gpu_ab_inproc.py "default:" 2>&1 | tee gpurun_out/ref_check_${1%.json}.log | tail -3
