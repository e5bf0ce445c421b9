#!/bin/bash
# reference hashes at 4096^3 (tests/golden/midsize_city4k.json): the in-process tool (prints ref-ok / REF-MISMATCH) + its test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_LOTS=64 AB_LEVELS=12 AB_STEP=3 AB_REPS=1 timeout 40 python tools/gpu_ab_inproc.py "default:" "plain:SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0,SVB_STAR_STORE=0,SVB_K64_ONEPASS=0,SVB_DEDUP_LAZY=0,SVB_LEAF_LAZY=0,SVB_INNER_MARKED=0,SVB_LEAF_NOTSTAR=0,SVB_SCAN_WIDE=0" 2>&1 | tee gpurun_out/ref_check.log | tail -4
timeout 40 python -m pytest tests/test_gpu_fullsize.py -q -k 4096 2>&1 | tail -3 | tee -a gpurun_out/ref_check.log
