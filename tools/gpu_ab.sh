#!/bin/bash
# A/B of environment toggles on the 16K^3 city build: gpurun -- tools/gpu_ab.sh "VAR=a" "VAR=b" ...
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
print("ms total %.1f vox %.1f dedup %.1f dag %d pairs %d" % (st["msTotal"], st["msVoxelize"], st["msDedup"], st["nNodesDAG"], st["nPairsTotal"]))
PY
for cfg in "$@"; do echo "== $cfg"; env $cfg python /tmp/q.py; done
