#!/usr/bin/env python
"""Where the end-to-end time of one build goes (host buffers in, .ssvdag image out), phase by phase, wall clock with a device
synchronisation after every phase.  Usage: python tools/e2e_breakdown.py [workload] [reps]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "city_16k"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
import torch  # noqa: E402
pkg = bench.load_pkg()
tris = pkg.meshgen.make_mesh(wl["mesh"], **wl["kw"])
T = tris.shape[0]
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
pinned = torch.from_numpy(tris.reshape(-1)).pin_memory()
o = pkg.GeomOctree(device=0)
rows = []
for it in range(reps + 2):
    t0 = time.perf_counter()
    o.set_triangles_ptr(pinned.data_ptr(), T)
    t1 = time.perf_counter()
    st = o.build(wl["levels"], wl["step"], bbox=bbox)
    t2 = time.perf_counter()
    sd = o.to_sdag()
    t3 = time.perf_counter()
    img = pkg.encoders.encode(o, "ssvdag")
    t4 = time.perf_counter()
    if it >= 2:
        rows.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0, st["msTotal"] * 1e-3, sd["msSdag"] * 1e-3])
a = np.array(rows).mean(axis=0) * 1e3
out = dict(zip(["h2d_ms", "build_ms", "to_sdag_ms", "encode_ms", "total_ms", "build_device_ms", "sdag_device_ms"], [round(float(x), 3) for x in a]))
out.update(triangles=int(T), h2d_GBps=round(T * 36 / (a[0] * 1e-3) / 1e9, 2), ssvdag_bytes=len(img), levels=o.level_sizes())
print(json.dumps(out))
