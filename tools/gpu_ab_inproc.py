"""In-process A/B of the per-call environment toggles on the 16K^3 city build (the mesh is generated and uploaded once).
    gpurun -- python tools/gpu_ab_inproc.py "NAME:VAR=a,VAR2=b" ...
Every configuration is built REPS times; prints build / voxelize / dedup ms and checks that the result (node counts and
the SSVDAG bytes) is identical across configurations and equal to the reference hashes in tests/golden/fullsize_city16k.json."""
import hashlib
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np

import __graft_entry__ as g

REPS = int(os.environ.get("AB_REPS", "2"))
pkg = g._pkg()
if os.environ.get("AB_GOLD"):   # a tests/golden/*size*.json file: its mesh, levels, step (and its hashes below)
    _g = json.loads((ROOT / "tests" / "golden" / os.environ["AB_GOLD"]).read_text())
    tris = pkg.meshgen.make_mesh(_g.get("mesh", "city"), **_g.get("kw", {"lots": _g.get("lots", 256)}))
    levels, step = _g["levels"], _g["step"]
else:
    tris = pkg.meshgen.city(int(os.environ.get("AB_LOTS", "256")))
    levels, step = int(os.environ.get("AB_LEVELS", "14")), int(os.environ.get("AB_STEP", "4"))
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
for _ in range(int(os.environ.get("AB_WARMUP", "2"))):
    t.build(levels, step, bbox=bbox)
gold = None
for gp in sorted((ROOT / "tests" / "golden").glob("*size_*.json")):   # reference hashes (tests/golden/make_fullsize.py)
    gj = json.loads(gp.read_text())
    if (gj.get("levels", 14), gj.get("step", 4)) == (levels, step) and gj.get("triangles", 11006740) == len(tris):
        gold = gj
first = None
out = {}
for spec in sys.argv[1:] or ["default:"]:
    name, _, kv = spec.partition(":")
    sets = dict(x.split("=", 1) for x in kv.split(",") if x)
    old = {k: os.environ.get(k) for k in sets}
    os.environ.update(sets)
    ms = []
    for _ in range(REPS):
        st = t.build(levels, step, bbox=bbox)
        ms.append((st["msTotal"], st["msVoxelize"], st["msDedup"]))
    sig = (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"])
    sv = hashlib.sha256(pkg.encoders.encode(t, "svdag")).hexdigest()
    t.to_sdag()
    ss = hashlib.sha256(pkg.encoders.encode(t, "ssvdag")).hexdigest()
    ok = "same" if first in (None, (sig, sv, ss)) else "DIFFERENT"
    first = first or (sig, sv, ss)
    if gold:
        ok += " ref-ok" if (sv == gold["files"]["svdag"]["sha256"] and ss == gold["files"]["ssvdag"]["sha256"] and sig == (gold["Voxels"], gold["SVO Nodes"], gold["DAG Nodes"])) else " REF-MISMATCH"
    best = min(ms)
    out[name] = {"ms_total": best[0], "ms_voxelize": best[1], "ms_dedup": best[2], "result": ok}
    print(f"{name:28s} total {best[0]:8.1f}  vox {best[1]:8.1f}  dedup {best[2]:7.1f}   {ok}", flush=True)
    for k, o in old.items():
        if o is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = o
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "ab_inproc.json").write_text(json.dumps(out, indent=1))
