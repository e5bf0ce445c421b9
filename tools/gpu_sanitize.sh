#!/bin/bash
# compute-sanitizer memcheck over small builds that exercise every kernel path (gpurun -- tools/gpu_sanitize.sh)
cd "$(dirname "$0")/.."
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
for mesh, kw, L, s in [("city", dict(lots=8), 9, 2), ("soup", dict(n=400, seed=7), 8, 1), ("terrain", dict(n=48), 8, 0)]:
    tris = pkg.meshgen.make_mesh(mesh, **kw)
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(6 << 20)
    st = t.build(L, s)
    t.to_sdag()
    img = pkg.encoders.encode(t, "ssvdag")
    u = pkg.GeomOctree(tris); u.build(L, s); u.cross_merge()
    print(mesh, st["nTotalVoxels"], st["nNodesDAG"], st["nBatches"], len(img))
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python /tmp/san.py 2>&1 | tail -12
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 3 python /tmp/san.py 2>&1 | grep -v 'Host Frame: .*python\|Host Frame: _Py\|Host Frame: Py\|ffi\|ctypes' | head -60
echo "exit: ${PIPESTATUS[0]}"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 5 python /tmp/san.py 2>&1 | tail -4
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 3 --print-limit 5 python /tmp/san.py 2>&1 | tail -3
