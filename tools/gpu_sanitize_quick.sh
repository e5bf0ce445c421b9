#!/bin/bash
# compute-sanitizer memcheck + racecheck over small multi-batch builds, bounded to ~220 s (gpurun -- tools/gpu_sanitize_quick.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
for mesh, kw, L, s in [("city", dict(lots=8), 9, 2), ("terrain", dict(n=48), 8, 1), ("soup", dict(n=300, seed=3), 8, 2)]:   # box mesh (k_slow_leaves, k_flat_leaves3), general triangles, mixed
    tris = pkg.meshgen.make_mesh(mesh, **kw)
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(6 << 20)
    st = t.build(L, s)
    t.to_sdag()
    print(mesh, st["nTotalVoxels"], st["nNodesDAG"], st["nBatches"])
PY
SVB_VX_STATS=1 timeout 110 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 10 python /tmp/san.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/san_memcheck.log
grep -v "vx-stats\] tiles" gpurun_out/san_memcheck.log | tail -8
timeout 110 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 5 python /tmp/san.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/san_racecheck.log
tail -5 gpurun_out/san_racecheck.log
