"""BASELINE.json configs[1] and configs[0] at full size -- the procedural heightfield terrain, 2.09 M general (non-flat)
triangles at 4096^3, levels 12 step 3, and the sphere + Menger sponge at 1024^3, levels 10 step 1 -- against the SHA-256
of the files the UNMODIFIED reference svbuilder wrote for the same inputs (tests/golden/size_*.json, minted by
tests/golden/make_fullsize.py terrain | spongeball; the CPU restatement reproduces both as well,
tests/golden/check_oracle_midsize.py)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDS = sorted((Path(__file__).resolve().parent / "golden").glob("size_*.json"))   # size_terrain4k.json, size_spongeball1k.json


@pytest.mark.parametrize("gold", GOLDS, ids=lambda p: p.stem)
def test_baseline_configs_files_equal_reference(pkg, meshgen, gold):
    g = json.loads(gold.read_text())
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


MID = Path(__file__).resolve().parent / "golden" / "midsize_city4k.json"


@pytest.mark.skipif(not MID.exists() or "cross" not in json.loads(MID.read_text()), reason="no CSVDAG reference hashes")
def test_cross_level_merge_equals_reference_at_4096(pkg, meshgen):
    """BASELINE.json configs[3] (city -> CSVDAG, `svbuilder ... -c`) at 64x64 lots / 4096^3: the -multi.svdag file and the
    number of nodes the cross-level merge eliminates against the unmodified reference (96 s on 8 cores; the CPU
    restatement reproduces the same file, 885 s)."""
    g = json.loads(MID.read_text())
    tris = meshgen.make_mesh("city", lots=g["lots"])
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(tris)
    t.build(g["levels"], g["step"], bbox=bbox)
    assert hashlib.sha256(pkg.encoders.encode(t, "svdag")).hexdigest() == g["files"]["svdag"]["sha256"]
    st = t.cross_merge()
    assert st["nCrossLevelMerged"] == g["cross"]["nodes_eliminated"]
    multi = pkg.encoders.encode(t, "svdag")
    assert len(multi) == g["cross"]["multi.svdag"]["bytes"]
    assert hashlib.sha256(multi).hexdigest() == g["cross"]["multi.svdag"]["sha256"], "-multi.svdag differs from the reference's file"


CROP = Path(__file__).resolve().parent / "golden" / "size_composite_crop4k.json"


@pytest.mark.skipif(not CROP.exists(), reason="tests/golden/size_composite_crop4k.json not minted")
def test_composite_octant_sharded_build_equals_reference(pkg, meshgen):
    """the 65536^3 composite is validated by (a) multi-GPU == single-GPU identity and (b) CPU parity on a
    cropped octant.  Here both at once on the cropped octant (0.77 M triangles: 17 % general terrain triangles under
    axis-aligned city boxes, 4096^3): the svb_shard_* protocol over 4 ranks (simulated on one device; the NCCL path runs the
    same calls, tests/test_gpu_parity.py / bench.py) must give the files the unmodified reference wrote."""
    from test_gpu_parity import _simulate_ranks
    g = json.loads(CROP.read_text())
    tris = meshgen.make_mesh(g["mesh"], **g["kw"])
    octs, stats = _simulate_ranks(pkg, tris, g["levels"], g["step"], 4)
    for st in stats:
        assert (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"]) == (g["Voxels"], g["SVO Nodes"], g["DAG Nodes"])
    t = octs[-1]
    assert hashlib.sha256(pkg.encoders.encode(t, "svdag")).hexdigest() == g["files"]["svdag"]["sha256"]
    assert t.to_sdag()["nNodesSDAG"] == g["SDAG Nodes"]
    assert hashlib.sha256(pkg.encoders.encode(t, "ssvdag")).hexdigest() == g["files"]["ssvdag"]["sha256"]


@pytest.mark.skipif(not CROP.exists(), reason="tests/golden/size_composite_crop4k.json not minted")
@pytest.mark.parametrize("budget_mb", [192, 384, 768, 1536], ids=lambda b: f"{b}MB")
def test_composite_octant_equals_reference_under_any_batch_plan(pkg, meshgen, budget_mb):
    """The result must not depend on how the sub-octrees are cut into tile batches.  Regression test of a round-1 bug found in
    round 2: the lazy leaf pass decided "this voxel mask is frozen" on the LIVE table, which CTAs of the same launch were
    already updating -- with batches that bring new voxel masks after the first one (terrain under boxes, several million leaf
    nodes per batch so that the grid does not fit the machine at once) late CTAs skipped nodes that held a mask's smallest
    order key, and the node order of the file changed with the batch plan (i.e. with the free device memory)."""
    g = json.loads(CROP.read_text())
    tris = meshgen.make_mesh(g["mesh"], **g["kw"])
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(budget_mb << 20)
    st = t.build(g["levels"], g["step"], bbox=bbox)
    assert st["nBatches"] >= 3
    assert (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"]) == (g["Voxels"], g["SVO Nodes"], g["DAG Nodes"])
    assert hashlib.sha256(pkg.encoders.encode(t, "svdag")).hexdigest() == g["files"]["svdag"]["sha256"]
    t.to_sdag()
    assert hashlib.sha256(pkg.encoders.encode(t, "ssvdag")).hexdigest() == g["files"]["ssvdag"]["sha256"]
