#!/bin/bash
# one GPU call: in-process A/B of the pipelined emit / children kernels and the permute key builder, then the small parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Z="SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0"
timeout 600 python tools/gpu_ab_inproc.py "base:$Z" "emit8:SVB_EMIT_PIPE=8,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0" "emit6:SVB_EMIT_PIPE=6,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0" \
  "children:SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=1,SVB_K64_PERM=0" "perm:SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=1" "all8:" "all6:SVB_EMIT_PIPE=6" "base2:$Z" 2>&1 | tee gpurun_out/ab_inproc.log | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_parity.log
