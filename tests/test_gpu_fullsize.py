"""Full-size checks at BASELINE.json's headline configuration (procedural city, 11.0 M triangles, 16384^3, levels 14
step 4): (1) the encoded files against the hashes of what the UNMODIFIED reference svbuilder wrote for the same
input (tests/golden/fullsize_city16k.json, minted once with tests/golden/make_fullsize.py -- a 1-hour CPU job), and
(2) size-independent properties: voxel-count conservation through the DAG / SDAG, pointer sanity, format agreement
of the ray-cast depth images."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "fullsize_city16k.json"
NULL = 0xFFFFFFFE


@pytest.fixture(scope="module")
def built(pkg, meshgen):
    tris = meshgen.make_mesh("city", lots=256)
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(tris)
    st = t.build(14, 4, bbox=bbox)
    dag_levels = t.levels_host()
    files = {"svdag": pkg.encoders.encode(t, "svdag"), "esvdag": pkg.encoders.encode(t, "esvdag")}
    sd = t.to_sdag()
    sdag_levels = t.levels_host()
    files["ussvdag"] = pkg.encoders.encode(t, "ussvdag")
    files["ssvdag"] = pkg.encoders.encode(t, "ssvdag")
    return dict(tris=tris, st=st, sd=sd, dag=dag_levels, sdag=sdag_levels, files=files)


def _voxels_through(levels):
    """Voxels represented by the root, counting shared subtrees once per reference (bottom-up multiplicities)."""
    cnt = np.array([bin(int(m)).count("1") for m in range(256)], dtype=np.uint64)[levels[-1]["mask"]]
    for lv in reversed(levels[:-1]):
        ch = lv["child"].astype(np.int64)
        ok = ch != NULL
        assert (ch[ok] < len(cnt)).all(), "child pointer out of range"
        assert np.array_equal(ok, ((lv["mask"][:, None] >> np.arange(8)) & 1).astype(bool)), "child mask and pointers disagree"
        c = np.zeros(ch.shape, dtype=np.uint64)
        c[ok] = cnt[ch[ok]]
        cnt = c.sum(axis=1)
    return int(cnt[0])


def test_voxel_count_is_conserved(built):
    assert _voxels_through(built["dag"]) == built["st"]["nTotalVoxels"]
    assert _voxels_through(built["sdag"]) == built["st"]["nTotalVoxels"]      # mirroring never changes a count
    assert built["st"]["nNodesDAG"] == 1 + sum(len(l["mask"]) for l in built["dag"][1:])
    assert built["sd"]["nNodesSDAG"] == sum(len(l["mask"]) for l in built["sdag"][1:])


def test_every_level_is_duplicate_free(built):
    """toDAG's defining property (geom_octree.cpp:462-548): no two nodes of a level share (mask, children)."""
    for l, lv in enumerate(built["dag"][1:], 1):
        key = np.concatenate([lv["mask"][:, None].astype(np.uint32), lv["child"]], axis=1)
        assert len(np.unique(key, axis=0)) == len(key), f"level {l} has duplicate nodes"


def test_depth_images_agree_across_formats(pkg, built):
    v = built["tris"].reshape(-1, 3)
    lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
    c, d = (lo + hi) / 2, float(np.linalg.norm(hi - lo))
    vi = pkg.camera.look_at_inv(c + np.array([0.55, 0.45, 0.5]) * d, c)
    pi = pkg.camera.perspective_inv(45.0, 1.0)
    f = built["files"]
    ref = pkg.raycast_depth(f["svdag"], "svdag", vi, pi, 512, 512, 20000)
    assert (ref[..., 0] > 0).mean() > 0.3 and ref[..., 1].max() == 13
    assert np.array_equal(pkg.raycast_depth(f["ussvdag"], "ussvdag", vi, pi, 512, 512, 20000), ref)
    s4 = pkg.raycast_depth(f["ssvdag"], "ssvdag", vi, pi, 512, 512, 20000)
    assert np.array_equal(s4[..., 0] > 0, ref[..., 0] > 0)


def test_cross_level_merge_at_full_size(pkg, built):
    """BASELINE.json configs[3]: CSVDAG of the same city at 16384^3 -- the merged file must render exactly like the
    plain one (geometry preserved), point only to existing nodes, and be smaller."""
    v = built["tris"].reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(built["tris"])
    st = t.build(14, 4, bbox=bbox)
    cm = t.cross_merge()
    assert cm["nCrossLevelMerged"] > 0 and cm["nNodesDAG"] == st["nNodesDAG"] - cm["nCrossLevelMerged"]
    lv = t.levels_host()
    sizes = [len(l["mask"]) for l in lv]
    for l, L in enumerate(lv[:-1]):
        ok = L["child"] != NULL
        tl = L["childLevel"][ok].astype(np.int64)
        assert ((tl >= 1) & (tl <= l + 1)).all(), f"level {l}: child level out of range"
        assert (L["child"][ok].astype(np.int64) < np.asarray(sizes)[tl]).all(), f"level {l}: dangling cross-level pointer"
    multi = pkg.encoders.encode(t, "svdag")
    assert len(multi) < len(built["files"]["svdag"])
    lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
    c, d = (lo + hi) / 2, float(np.linalg.norm(hi - lo))
    vi = pkg.camera.look_at_inv(c + np.array([0.55, 0.45, 0.5]) * d, c)
    pi = pkg.camera.perspective_inv(45.0, 1.0)
    a = pkg.raycast_depth(built["files"]["svdag"], "svdag", vi, pi, 512, 512, 20000)
    b = pkg.raycast_depth(multi, "svdag", vi, pi, 512, 512, 20000)
    assert np.array_equal(a, b)


@pytest.mark.skipif(not GOLD.exists(), reason="tests/golden/fullsize_city16k.json not minted")
def test_files_equal_reference_at_full_size(built):
    g = json.loads(GOLD.read_text())
    assert built["st"]["nTotalVoxels"] == g["Voxels"]
    assert built["st"]["nNodesSVO"] == g["SVO Nodes"]
    assert built["st"]["nNodesDAG"] == g["DAG Nodes"]
    assert built["sd"]["nNodesSDAG"] == g["SDAG Nodes"]
    for k, want in g["files"].items():
        got = built["files"][k]
        assert len(got) == want["bytes"], k
        assert hashlib.sha256(got).hexdigest() == want["sha256"], f"{k}: bytes differ from the reference's file"


SIZED = [p for p in sorted((Path(__file__).resolve().parent / "golden").glob("*size_city*.json")) if p.name != GOLD.name]


@pytest.mark.parametrize("gold", SIZED, ids=lambda p: p.stem)
def test_files_equal_reference_at_4096(pkg, meshgen, gold):
    """The same city generator at 64x64 lots / 4096^3 (levels 12, step 3; 0.69 M triangles) and 128x128 lots / 8192^3
    (levels 13, step 3; 2.75 M triangles) -- every lot spans 64 voxels as in the 16K^3 workload -- against the hashes of
    what the unmodified reference svbuilder wrote (tests/golden/make_fullsize.py midsize | bigsize)."""
    g = json.loads(gold.read_text())
    tris = meshgen.make_mesh("city", lots=g["lots"])
    assert len(tris) == g["triangles"]
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(1 << 30)
    st = t.build(g["levels"], g["step"], bbox=bbox)
    assert (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"]) == (g["Voxels"], g["SVO Nodes"], g["DAG Nodes"])
    files = {"svdag": pkg.encoders.encode(t, "svdag"), "esvdag": pkg.encoders.encode(t, "esvdag")}
    sd = t.to_sdag()
    assert sd["nNodesSDAG"] == g["SDAG Nodes"]
    files["ussvdag"] = pkg.encoders.encode(t, "ussvdag")
    files["ssvdag"] = pkg.encoders.encode(t, "ssvdag")
    for k, want in g["files"].items():
        if k not in files:
            continue
        assert len(files[k]) == want["bytes"], k
        assert hashlib.sha256(files[k]).hexdigest() == want["sha256"], f"{k}: bytes differ from the reference's file"
