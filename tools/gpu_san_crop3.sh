#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
sed -i 's/^if os.environ.get("DIRTY")/if False and os.environ.get("DIRTY")/' /tmp/san4.py 2>/dev/null
cat > /tmp/san5.py <<'PY'
import sys, json, hashlib, os
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
gj = json.load(open("tests/golden/size_composite_crop4k.json"))
tris = pkg.meshgen.make_mesh(gj["mesh"], **gj["kw"])
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
st = t.build(gj["levels"], gj["step"], bbox=bbox)
sv = hashlib.sha256(pkg.encoders.encode(t, "svdag")).hexdigest()
print("batches", st["nBatches"], "svdag ok", sv == gj["files"]["svdag"]["sha256"], flush=True)
PY
SVB_CHILDREN_TMA=0 timeout 500 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 8 python /tmp/san5.py > gpurun_out/san_crop_initcheck_notma.log 2>&1
echo "initcheck exit $?"
grep -n "Uninitialized\|    at \|svdag ok\|ERROR SUMMARY" gpurun_out/san_crop_initcheck_notma.log | sed 's/(.*//' | head -40
