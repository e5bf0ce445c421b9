#!/bin/bash
# one GPU call: in-process A/B of the environment toggles, then the small parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Z="SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0,SVB_STAR_STORE=0,SVB_K64_ONEPASS=0,SVB_DEDUP_LAZY=0,SVB_LEAF_LAZY=0,SVB_INNER_MARKED=0"
timeout 600 python tools/gpu_ab_inproc.py "base:$Z" "unmarked:SVB_INNER_MARKED=0" "marked:" "marked_npt4:SVB_K64_ONEPASS=4" "marked_npt1:SVB_K64_ONEPASS=1" "all2:" 2>&1 | tee gpurun_out/ab_inproc.log | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_parity.log
