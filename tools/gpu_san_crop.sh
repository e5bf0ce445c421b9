#!/bin/bash
# compute-sanitizer (initcheck, memcheck, racecheck) over the cropped-composite build at 4096^3 (multi-batch, general + flat triangles)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san3.py <<'PY'
import sys, json, hashlib
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
print("batches", st["nBatches"], "svdag ok", sv == gj["files"]["svdag"]["sha256"])
PY
for tool in initcheck memcheck racecheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 6 python /tmp/san3.py > gpurun_out/san_crop_$tool.log 2>&1
  echo "$tool exit $?"; grep -v "^$" gpurun_out/san_crop_$tool.log | tail -12
done
