#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cat > /tmp/q.py <<'PY'
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
for it in range(3):
    st = t.build(14, 4, bbox=bbox)
print("ms total %.1f vox %.1f dedup %.1f dag %d" % (st["msTotal"], st["msVoxelize"], st["msDedup"], st["nNodesDAG"]))
PY
echo default; python /tmp/q.py
echo no-tstar-experiment; SVB_EXP_NOTSTAR=1 python /tmp/q.py
