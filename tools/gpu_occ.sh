#!/bin/bash
# voxelizer occupancy sweep of the slow-stream classify kernel (SVB_VX_OCC = CTAs of 256 threads per SM)
cd "$(dirname "$0")/.."
cat > /tmp/q.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
for it in range(3):
    st = t.build(14, 4, bbox=bbox)
print("ms total %.1f vox %.1f dedup %.1f" % (st["msTotal"], st["msVoxelize"], st["msDedup"]))
PY
for o in 3 4 5 6; do echo "occ $o"; SVB_VX_OCC=$o python /tmp/q.py; done
