#!/bin/bash
# one GPU call: in-process A/B of the environment toggles, then the small parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Z="SVB_EMIT_PIPE=0,SVB_CHILDREN_PIPE=0,SVB_K64_PERM=0,SVB_STAR_STORE=0,SVB_K64_ONEPASS=0,SVB_DEDUP_LAZY=0,SVB_LEAF_LAZY=0,SVB_INNER_MARKED=0"
SVB_VX_STATS=1 timeout 600 python tools/gpu_ab_inproc.py "base:$Z" "tracked:SVB_LEAF_NOTSTAR=0" "elided:" "tracked2:SVB_LEAF_NOTSTAR=0" "elided2:" 2> gpurun_out/ab_stats.err | tee gpurun_out/ab_inproc.log | tail -12
grep -c "voxelizing again" gpurun_out/ab_stats.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_parity.log
