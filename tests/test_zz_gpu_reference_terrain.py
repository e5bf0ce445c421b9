"""BASELINE.json configs[1] at full size -- procedural heightfield terrain, 2.09 M general (non-flat) triangles at 4096^3,
levels 12 step 3 -- against the SHA-256 of the files the UNMODIFIED reference svbuilder wrote for the same input
(tests/golden/size_terrain4k.json, minted by tests/golden/make_fullsize.py terrain).  Kept in a file that sorts last: it
is the one reference comparison of the general-triangle classify path at a size no CPU oracle run accompanies."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "size_terrain4k.json"


@pytest.mark.skipif(not GOLD.exists(), reason="tests/golden/size_terrain4k.json not minted")
def test_terrain_4096_files_equal_reference(pkg, meshgen):
    g = json.loads(GOLD.read_text())
    tris = meshgen.make_mesh(g["mesh"], **g["kw"])
    assert len(tris) == g["triangles"]
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(tris)
    st = t.build(g["levels"], g["step"], bbox=bbox)
    assert (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"]) == (g["Voxels"], g["SVO Nodes"], g["DAG Nodes"])
    files = {"svdag": pkg.encoders.encode(t, "svdag"), "esvdag": pkg.encoders.encode(t, "esvdag")}
    sd = t.to_sdag()
    assert sd["nNodesSDAG"] == g["SDAG Nodes"]
    files["ussvdag"] = pkg.encoders.encode(t, "ussvdag")
    files["ssvdag"] = pkg.encoders.encode(t, "ssvdag")
    for k, want in g["files"].items():
        if k in files:
            assert len(files[k]) == want["bytes"], k
            assert hashlib.sha256(files[k]).hexdigest() == want["sha256"], f"{k}: bytes differ from the reference's file"
