"""GPU parity tests: the CUDA path through the C ABI (libsvb.so) against the CPU oracle and the
golden files minted from the unmodified reference binary.  Bit-exact: node arrays, counts, bytes."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case

pytestmark = pytest.mark.gpu


def _assert_levels_equal(got, want, what, fields=("mask", "child")):
    assert len(got) == len(want), f"{what}: level count {len(got)} != {len(want)}"
    for l, (a, b) in enumerate(zip(got, want)):
        assert len(a["mask"]) == len(b["mask"]), f"{what}: level {l} has {len(a['mask'])} nodes, oracle {len(b['mask'])}"
        for f in fields:
            if not np.array_equal(a[f], b[f]):
                bad = np.nonzero((a[f].reshape(len(a["mask"]), -1) != b[f].reshape(len(b["mask"]), -1)).any(axis=1))[0]
                raise AssertionError(f"{what}: level {l} field {f} differs at {len(bad)} nodes, first {bad[:5]}: "
                                     f"got {a[f][bad[0]]} want {b[f][bad[0]]}")


def _oracle_levels(o):
    return [o.level(l) for l in range(o.levels)]


@pytest.mark.parametrize("path", [p for p in GOLDEN if not p.stem.endswith("_c")], ids=lambda p: p.stem)
def test_golden_files_bit_exact(pkg, path):
    """mesh -> SVDAG/ESVDAG/USSVDAG/SSVDAG on the GPU == the files the reference binary wrote."""
    g = golden_case(path)
    t = pkg.GeomOctree(g["tris"])
    st = t.build(g["levels"], g["step"])
    vox, svo, dag, sdag = (int(x) for x in g["stats"])
    assert (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"]) == (vox, svo, dag)
    sizes = t.level_sizes()
    for lev in range(1, g["levels"]):
        if g["reduced"][lev].sum():
            assert sizes[lev] == int(g["reduced"][lev][1])
    assert pkg.encoders.encode(t, "svdag") == g["files"]["svdag"]
    assert pkg.encoders.encode(t, "esvdag") == g["files"]["esvdag"]
    st = t.to_sdag()
    assert st["nNodesSDAG"] == sdag
    assert pkg.encoders.encode(t, "ussvdag") == g["files"]["ussvdag"]
    assert pkg.encoders.encode(t, "ssvdag") == g["files"]["ssvdag"]


CASES = [
    # mesh, kwargs, levels, step
    ("sphere", dict(n_lat=64, n_lon=128), 8, 0),
    ("sphere", dict(n_lat=64, n_lon=128), 9, 2),
    ("sphere", dict(n_lat=64, n_lon=128), 9, 3),
    ("city", dict(lots=8), 8, 0),
    ("city", dict(lots=8), 9, 2),
    ("terrain", dict(n=64), 8, 1),
    ("terrain", dict(n=64), 9, 3),
    ("sphere_menger", dict(n_lat=32, n_lon=64, sponge_level=2), 8, 0),
    ("sphere_menger", dict(n_lat=32, n_lon=64, sponge_level=2), 9, 2),
    ("soup", dict(n=400, seed=7), 8, 0),     # degenerate triangles (points, segments, slivers), lattice ties
    ("soup", dict(n=400, seed=7), 8, 2),
    ("soup", dict(n=1200, seed=3), 7, 1),
]


@pytest.mark.parametrize("mesh,kw,levels,step", CASES, ids=[f"{m}-L{l}-s{s}" for m, _, l, s in CASES])
def test_dag_and_sdag_match_oracle(pkg, orc, meshgen, mesh, kw, levels, step):
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    t = pkg.GeomOctree(tris)
    st = t.build(levels, step)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG", "nNodesLastLevSVO", "nNodesLastLevDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG")
    if step == 0:
        assert t.level_sizes_svo()[1:] == o.level_sizes_before_dag()[1:]
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")
    o.to_sdag()
    st = t.to_sdag()
    assert st["nNodesSDAG"] == o.stat("nNodesSDAG")
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "SDAG", fields=("mask", "child", "mirror", "inv"))
    assert pkg.encoders.encode(t, "ussvdag") == o.encode("ussvdag")
    assert pkg.encoders.encode(t, "ssvdag") == o.encode("ssvdag")


def test_batch_splitting_gives_identical_result(pkg, meshgen):
    """A tiny batch budget forces many tile batches; node order must not depend on the batching."""
    tris = meshgen.make_mesh("sphere", n_lat=64, n_lon=128)
    a = pkg.GeomOctree(tris)
    a.build(9, 2)
    b = pkg.GeomOctree(tris)
    b.set_batch_budget(8 << 20)
    st = b.build(9, 2)
    assert st["nBatches"] > 1
    _assert_levels_equal(b.levels_host(), a.levels_host(), "batched DAG")


def test_sdag_from_uploaded_dag(pkg, orc, meshgen):
    """Stage entry: a DAG produced elsewhere (here: the oracle) uploaded, reduced to an SDAG on the GPU."""
    tris = meshgen.make_mesh("terrain", n=48)
    o = orc.OracleOctree(tris)
    o.build(8, 0)
    lv = _oracle_levels(o)
    lo, hi = o.scene_bbox()
    t = pkg.GeomOctree()
    t.upload_levels(lv, np.concatenate([lo, hi]).astype(np.float32), 1.0, o.stat("nTotalVoxels"))
    o.to_sdag()
    t.to_sdag()
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "SDAG(uploaded)", fields=("mask", "child", "mirror", "inv"))


def test_wrong_state_and_bad_args(pkg, meshgen):
    tris = meshgen.make_mesh("sphere", n_lat=8, n_lon=16)
    t = pkg.GeomOctree(tris)
    with pytest.raises(pkg.SvbError):
        t.to_sdag()                      # "ERROR! This is not a DAG or SDAG!" (geom_octree.cpp:560-563)
    with pytest.raises(pkg.SvbError):
        t.build(6, 5)                    # step + 1 must be < levels
    t.build(5, 0)
    t.to_sdag()
    with pytest.raises(pkg.SvbError):
        t.to_sdag()                      # already an SDAG
    with pytest.raises(pkg.SvbError):
        pkg.encoders.encode(t, "svdag")  # EncodedSVDAG::encode needs DAG state (encoded_svdag.cpp:109)


def test_empty_scene(pkg):
    t = pkg.GeomOctree(np.zeros((0, 3, 3), np.float32))
    st = t.build(5, 0, bbox=(np.zeros(3), np.ones(3)))
    assert st["nTotalVoxels"] == 0 and st["nNodesDAG"] == 1
    assert t.level_sizes() == [1, 0, 0, 0, 0]


@pytest.mark.parametrize("mesh,kw,levels,step", [
    ("city", dict(lots=16), 10, 3),
    ("terrain", dict(n=128), 10, 2),
    ("sphere_menger", dict(n_lat=48, n_lon=96, sponge_level=2), 9, 0),
    ("terrain", dict(n=1024), 12, 3),        # BASELINE configs[1] at full size: 2.09 M general triangles at 4096^3
    ("composite_crop", dict(n_terrain=257, lots=64, octant=0), 11, 2),   # terrain under axis-aligned boxes
], ids=["city", "terrain", "spongeball", "terrain4k", "composite-crop"])
def test_filtered_classifier_equals_exact_predicate(pkg, meshgen, mesh, kw, levels, step, monkeypatch):
    """The FP64 interval filter in front of the SAT predicate must never change a decision: a build
    with every child decided by the reference-order predicate (SVB_CLASSIFY=exact) gives the same
    octree, node for node.  (city: walls lying exactly in voxel faces -> many exact ties.)"""
    tris = meshgen.make_mesh(mesh, **kw)
    a = pkg.GeomOctree(tris)
    sa = a.build(levels, step)
    monkeypatch.setenv("SVB_CLASSIFY", "exact")
    b = pkg.GeomOctree(tris)
    sb = b.build(levels, step)
    monkeypatch.delenv("SVB_CLASSIFY")
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG", "nPairsTotal"):
        assert sa[k] == sb[k], k
    _assert_levels_equal(a.levels_host(), b.levels_host(), "filtered vs exact")


def _affine(tris, scale, offset):
    t = tris.reshape(-1, 3).astype(np.float64) * np.asarray(scale) + np.asarray(offset)
    return np.ascontiguousarray(t.astype(np.float32).reshape(-1, 9))


@pytest.mark.parametrize("centre", ["auto", "chain"])
@pytest.mark.parametrize("mesh,kw,levels,step,bbox_mode", [
    ("terrain", dict(n=48), 8, 0, "scene"),
    ("city", dict(lots=8), 9, 2, "scene"),
    ("sphere", dict(n_lat=32, n_lon=64), 9, 3, "scene"),
    ("city", dict(lots=8), 9, 2, "double"),
    ("terrain", dict(n=48), 9, 1, "double"),
    ("soup", dict(n=400, seed=7), 8, 2, "double"),       # flat and general triangles in one scene, replayed centres (k_slow_leaves<false> next to the pair path)
], ids=["terrain-s0", "city-s2", "sphere-s3", "city-doublebox", "terrain-doublebox", "soup-doublebox"])
def test_arbitrary_bbox_matches_oracle(pkg, orc, meshgen, mesh, kw, levels, step, bbox_mode, centre, monkeypatch):
    """Scenes whose bbox is NOT the unit cube: node centres are no longer short dyadic numbers, so the
    reference's centre chain (geom_octree.cpp:222-230) and the float narrowing of sub-octree boxes (:177-184,
    :340-344) matter.  "scene": float bbox of a scaled/shifted mesh (the chain is still exact -> closed-form
    centres); "double": a caller-supplied bbox with full 53-bit doubles (chain rounds -> the kernels must replay
    it).  SVB_CENTRE=chain forces the replay everywhere; all variants must equal the oracle bit for bit."""
    tris = _affine(meshgen.make_mesh(mesh, **kw), (3.7, 2.9, 5.3), (11.3, -5.1, 2.9))
    bbox = None
    if bbox_mode == "double":
        v = tris.reshape(-1, 3).astype(np.float64)
        lo, hi = v.min(axis=0), v.max(axis=0)
        bbox = (lo - np.array([0.1, 0.2, 0.3]) / 3.0, hi + np.array([0.7, 0.1, 0.05]) / 7.0)
    if centre == "chain":
        monkeypatch.setenv("SVB_CENTRE", "chain")
    o = orc.OracleOctree(tris)
    o.build(levels, step, bbox=bbox)
    t = pkg.GeomOctree(tris)
    st = t.build(levels, step, bbox=bbox)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG (arbitrary bbox)")
    o.to_sdag()
    t.to_sdag()
    assert pkg.encoders.encode(t, "ssvdag") == o.encode("ssvdag")


@pytest.mark.parametrize("mesh,kw,levels,step", [
    ("city", dict(lots=8), 9, 2),
    ("terrain", dict(n=64), 8, 0),
    ("sphere_menger", dict(n_lat=32, n_lon=64, sponge_level=2), 9, 3),
], ids=["city-s2", "terrain-s0", "spongeball-s3"])
def test_wide_leaf_order_keys(pkg, orc, meshgen, mesh, kw, levels, step, monkeypatch):
    """Leaf-level order keys kept as two words (what a 64K^3 build needs: tile + triangle + path bits > 63):
    forced on small scenes with SVB_WIDE_LEAF=1, the result must not change."""
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    monkeypatch.setenv("SVB_WIDE_LEAF", "1")
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(24 << 20)          # several batches: the (hi, lo) minima must merge across them
    st = t.build(levels, step)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG (wide leaf keys)")
    if step > 0:
        octs, _ = _simulate_ranks(pkg, tris, levels, step, 3)
        for r, oc in enumerate(octs):
            _assert_levels_equal(oc.levels_host(), _oracle_levels(o), f"sharded DAG rank {r} (wide leaf keys)")


def _simulate_ranks(pkg, tris, levels, step, world, merge_seed=0):
    """Run the multi-GPU protocol with `world` contexts on ONE device, doing the all-gathers by hand
    (torch.cat of the per-rank export buffers).  Exercises svb_shard_* end to end without NCCL."""
    import torch
    dev = torch.device("cuda", 0)
    octs = [pkg.GeomOctree(tris) for _ in range(world)]
    for o in octs:
        o.set_merge_seed(merge_seed)
    bbox = octs[0].scene_bbox()
    for r, o in enumerate(octs):
        o.shard_build(levels, step, bbox, r, world)
    first, last, ntiles, _ = octs[0].shard_info()
    counters = [o.shard_info()[3] for o in octs]
    for g in range(last, first - 1, -1):
        cr = [o.shard_level_count(g) for o in octs]
        counts = np.array([c[0] for c in cr], dtype=np.uint64)
        rec = cr[0][1]
        stride = int(max(16, (int(counts.max()) * rec + 15) // 16 * 16))
        bufs = []
        for o in octs:
            b = torch.empty(stride, dtype=torch.uint8, device=dev)    # padding beyond the records is never read
            o.shard_export_level(g, b.data_ptr())
            bufs.append(b)
        for o in octs:
            o.synchronize()                                           # exports are stream-ordered on each context's own stream
        allb = torch.cat(bufs)
        torch.cuda.synchronize()
        for o in octs:
            o.shard_import_level(g, allb.data_ptr(), counts, stride)
        for o in octs:
            o.synchronize()                                           # allb is dropped at the next iteration
    nt = max(ntiles, 1)
    roots = []
    for o in octs:
        b = torch.empty(nt, dtype=torch.int32, device=dev)
        o.shard_export_roots(b.data_ptr())
        o.synchronize()
        roots.append(b)
    allr = torch.cat(roots)
    torch.cuda.synchronize()
    totals = np.sum(np.array(counters, dtype=np.uint64), axis=0)
    stats = []
    for o in octs:
        o.shard_import_roots(allr.data_ptr())
        stats.append(o.shard_finish(totals))
    return octs, stats


@pytest.mark.parametrize("mesh,kw,levels,step,world", [
    ("sphere", dict(n_lat=64, n_lon=128), 9, 2, 2),
    ("city", dict(lots=8), 9, 3, 4),
    ("terrain", dict(n=64), 9, 2, 8),
    ("sphere_menger", dict(n_lat=32, n_lon=64, sponge_level=2), 8, 5, 3),   # 2-level sub-octrees: roots are the K64 level
    ("sphere", dict(n_lat=32, n_lon=64), 7, 5, 2),                           # 1-level sub-octrees: roots are voxel masks
    ("city", dict(lots=8), 9, 2, 8),                                         # SVB_SHARD=octant (see below)
    ("soup", dict(n=400, seed=7), 8, 2, 5),
    # ranks with an EMPTY share (six of the eight top-level octants hold nothing; more ranks than sub-octrees): their
    # order-key widths must still match everyone else's, or they rank the merged tables differently
    ("sphere", dict(n_lat=32, n_lon=64, center=(0.2, 0.2, 0.2), radius=0.15), 8, 2, 8),
    ("sphere", dict(n_lat=32, n_lon=64, center=(0.2, 0.2, 0.2), radius=0.15), 7, 1, 12),
], ids=["sphere-w2", "city-w4", "terrain-w8", "spongeball-w3", "sphere-leafroots-w2", "city-octant-w8", "soup-w5",
        "cornersphere-octant-w8", "cornersphere-w12"])
def test_sharded_protocol_equals_single_gpu_build(pkg, meshgen, mesh, kw, levels, step, world, monkeypatch, request):
    if "octant" in request.node.callspec.id:
        monkeypatch.setenv("SVB_SHARD", "octant")    # the pure top-level-octant split BASELINE.json names (unbalanced for flat scenes)
    tris = meshgen.make_mesh(mesh, **kw)
    ref = pkg.GeomOctree(tris)
    sref = ref.build(levels, step)
    want = ref.levels_host()
    octs, stats = _simulate_ranks(pkg, tris, levels, step, world)
    for r, (o, st) in enumerate(zip(octs, stats)):
        for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG", "nNodesLastLevSVO", "nPairsTotal"):
            assert st[k] == sref[k], (r, k)
        _assert_levels_equal(o.levels_host(), want, f"rank {r} of {world}")
    # and the merged octree goes on through toSDAG + the encoders like any other
    ref.to_sdag()
    octs[-1].to_sdag()
    assert pkg.encoders.encode(octs[-1], "ssvdag") == pkg.encoders.encode(ref, "ssvdag")


@pytest.mark.parametrize("path", [p for p in GOLDEN if p.stem.endswith("_c")], ids=lambda p: p.stem)
def test_cross_level_merge_golden_files(pkg, path):
    """mesh -> SVDAG -> CSVDAG (-c): <base>_<L>.svdag and <base>_<L>-multi.svdag equal the reference's files."""
    g = golden_case(path)
    t = pkg.GeomOctree(g["tris"])
    t.build(g["levels"], g["step"])
    assert pkg.encoders.encode(t, "svdag") == g["files"]["svdag"]
    t.cross_merge()
    assert pkg.encoders.encode(t, "svdag") == g["files"]["multi_svdag"]


@pytest.mark.parametrize("mesh,kw,levels,step", [
    ("sphere", dict(n_lat=64, n_lon=128), 9, 0),
    ("city", dict(lots=8), 9, 2),
    ("terrain", dict(n=64), 8, 1),
    ("sphere_menger", dict(n_lat=32, n_lon=64, sponge_level=2), 9, 2),
], ids=["sphere", "city", "terrain", "spongeball"])
def test_cross_level_merge_matches_oracle(pkg, orc, meshgen, mesh, kw, levels, step):
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    removed = o.cross_merge()
    t = pkg.GeomOctree(tris)
    t.build(levels, step)
    st = t.cross_merge()
    assert st["nCrossLevelMerged"] == removed
    assert st["nNodesDAG"] == o.stat("nNodesDAG")
    got, want = t.levels_host(), _oracle_levels(o)
    _assert_levels_equal(got, want, "CSVDAG", fields=("mask", "child"))
    for l, (a, b) in enumerate(zip(got, want)):
        live = b["child"] != 0xFFFFFFFE
        assert np.array_equal(a["childLevel"][live], b["childLevel"][live]), f"childLevels differ at level {l}"
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")


def test_svbuilder_cli_writes_reference_files(pkg, tmp_path):
    """The C++ drop-in tool: `svbuilder m.obj L s` from an ASCII OBJ -> the four files of the reference, byte for byte."""
    import subprocess
    tool = pkg.lib_path().parent / "svbuilder"
    if not tool.exists():
        pytest.skip("svbuilder host tool not built")
    for path in [p for p in GOLDEN if p.stem in ("sphere_L7_s1", "city_L7_s2", "terrain_L6_s0", "city_L7_s1_c")]:
        g = golden_case(path)
        d = tmp_path / g["name"]
        d.mkdir()
        pkg.meshgen.write_obj(d / "m.obj", g["tris"])
        cmd = [str(tool), str(d / "m.obj"), str(g["levels"]), str(g["step"])] + (["-c"] if g["cross"] else [])
        r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        for ext, data in g["files"].items():
            f = d / (f"m_{g['levels']}-multi.svdag" if ext == "multi_svdag" else f"m_{g['levels']}.{ext}")
            assert f.read_bytes() == data, f"{g['name']}: {f.name} differs from the reference's file"
        assert (d / "m.obj.bincache").exists() and (d / "stats.txt").exists()


def test_svbuilder_cli_from_svdag_input(pkg, tmp_path):
    """`svbuilder m_7.svdag 7 0 [-c]` (main.cpp:96-101): decode a saved DAG, then SSVDAG / cross-level merge on the GPU.
    The reference binary writes exactly the files of the direct build (checked when the goldens were minted), so the
    goldens of the direct build are the expectation."""
    import subprocess
    tool = pkg.lib_path().parent / "svbuilder"
    if not tool.exists():
        pytest.skip("svbuilder host tool not built")
    plain = golden_case([p for p in GOLDEN if p.stem == "city_L7_s2"][0])
    d = tmp_path / "plain"
    d.mkdir()
    (d / "m_7.svdag").write_bytes(plain["files"]["svdag"])
    r = subprocess.run([str(tool), str(d / "m_7.svdag"), "7", "0"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in ("svdag", "esvdag", "ussvdag", "ssvdag"):
        assert (d / f"m_7_7.{ext}").read_bytes() == plain["files"][ext], ext
    cross = golden_case([p for p in GOLDEN if p.stem == "city_L7_s1_c"][0])
    d = tmp_path / "cross"
    d.mkdir()
    (d / "m_7.svdag").write_bytes(cross["files"]["svdag"])
    r = subprocess.run([str(tool), str(d / "m_7.svdag"), "7", "0", "-c"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (d / "m_7_7.svdag").read_bytes() == cross["files"]["svdag"]
    assert (d / "m_7_7-multi.svdag").read_bytes() == cross["files"]["multi_svdag"]


# ------------------------------------------------------------------ round 1e: reduced-work paths against the plain ones
LEGACY = {  # every environment toggle that selects the straightforward variant of a kernel / pass
    "SVB_EMIT_PIPE": "0", "SVB_CHILDREN_PIPE": "0", "SVB_STAR_STORE": "0", "SVB_K64_PERM": "0", "SVB_K64_ONEPASS": "0",
    "SVB_DEDUP_LAZY": "0", "SVB_LEAF_LAZY": "0", "SVB_INNER_MARKED": "0", "SVB_LEAF_NOTSTAR": "0", "SVB_SCAN_WIDE": "0",
    "SVB_K64_NOTSTAR": "0", "SVB_ROOTS_ONCE": "0", "SVB_EMIT_WARP": "0", "SVB_SCAN_MULTI": "0", "SVB_FAST_ILP": "1", "SVB_LEAVES_ILP": "1", "SVB_CHILDREN_TMA": "0", "SVB_SLOW_LEAVES": "0", "SVB_SLOW_LEAVES_MIXED": "0",
}


@pytest.mark.parametrize("mesh,kw,levels,step,budget", [
    ("city", dict(lots=16), 10, 2, 24 << 20),
    ("sphere", dict(n_lat=64, n_lon=128), 9, 2, 8 << 20),
    ("terrain", dict(n=96), 9, 3, 8 << 20),
    ("soup", dict(n=1200, seed=3), 8, 2, 4 << 20),
], ids=["city", "sphere", "terrain", "soup"])
def test_reduced_work_paths_equal_plain_paths(pkg, orc, meshgen, mesh, kw, levels, step, budget, monkeypatch):
    """Single-pass / frozen-entry dedup, lazy leaf pass, leaf level without first touches (direct query for new voxel
    masks), star stores, pipelined emit: many small batches, default paths vs. every toggle off vs. the oracle."""
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    fast = pkg.GeomOctree(tris)
    fast.set_batch_budget(budget)
    st = fast.build(levels, step)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(fast.levels_host(), _oracle_levels(o), "DAG (default paths, many batches)")
    for k, v in LEGACY.items():
        monkeypatch.setenv(k, v)
    plain = pkg.GeomOctree(tris)
    plain.set_batch_budget(budget)
    sp = plain.build(levels, step)
    assert (sp["nTotalVoxels"], sp["nNodesSVO"], sp["nNodesDAG"]) == (st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"])
    assert sp["nBatches"] > 1 and st["nBatches"] > 1      # (the batch plans may differ: the paths hold different amounts of memory)
    _assert_levels_equal(plain.levels_host(), _oracle_levels(o), "DAG (plain paths, many batches)")


@pytest.mark.parametrize("toggle", sorted(LEGACY), ids=lambda s: s[4:].lower())
def test_each_toggle_alone(pkg, meshgen, toggle, monkeypatch):
    """Each reduced-work path switched off on its own (the others stay on): same DAG, same SSVDAG bytes."""
    tris = meshgen.make_mesh("city", lots=16)
    a = pkg.GeomOctree(tris)
    a.set_batch_budget(24 << 20)
    a.build(10, 2)
    want_levels = a.levels_host()
    a.to_sdag()
    want = pkg.encoders.encode(a, "ssvdag")
    monkeypatch.setenv(toggle, LEGACY[toggle])
    b = pkg.GeomOctree(tris)
    b.set_batch_budget(24 << 20)
    b.build(10, 2)
    _assert_levels_equal(b.levels_host(), want_levels, f"DAG with {toggle}={LEGACY[toggle]}")
    b.to_sdag()
    assert pkg.encoders.encode(b, "ssvdag") == want


@pytest.mark.parametrize("centre", ["auto", "chain"])
@pytest.mark.parametrize("occ", ["3", "4", "5", "6"])
def test_slow_stream_fused_leaves_every_variant(pkg, orc, meshgen, occ, centre, monkeypatch, capfd):
    """Box mesh: the slow stream's last two levels are decided in place (k_slow_leaves): every instantiation (CTAs/SM x
    closed-form / replayed centres) equals the oracle on a scaled, shifted city whose hypotenuses produce voxels inside
    the filter's margin (the reference-order predicate decides those: nExactTests > 0), and the last level really has
    no pair list."""
    tris = _affine(meshgen.make_mesh("city", lots=8), (3.7, 2.9, 5.3), (11.3, -5.1, 2.9))
    o = orc.OracleOctree(tris)
    o.build(9, 2)
    monkeypatch.setenv("SVB_VX_OCC_SL", occ)
    monkeypatch.setenv("SVB_VX_STATS", "1")
    if centre == "chain":
        monkeypatch.setenv("SVB_CENTRE", "chain")
    t = pkg.GeomOctree(tris)
    st = t.build(9, 2)
    err = capfd.readouterr().err
    import re
    lv = [re.search(r"level (\d+)/(\d+) nodes \d+ flat pairs (\d+) slow pairs (\d+)", ln) for ln in err.splitlines() if ln.startswith("[vx-stats] tiles")]
    last = [m for m in lv if m and m.group(1) == m.group(2)]
    assert last and all(m.group(3) == "0" and m.group(4) == "0" for m in last), err[-600:]
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    assert st["nExactTests"] > 0
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), f"DAG (k_slow_leaves, {occ} CTAs/SM, centres {centre})")
    monkeypatch.setenv("SVB_SLOW_LEAVES", "0")
    u = pkg.GeomOctree(tris)
    su = u.build(9, 2)
    assert su["nPairsTotal"] == st["nPairsTotal"]   # the same pairs, decided elsewhere


@pytest.mark.parametrize("lattice", [64, 0], ids=["lattice", "random"])
def test_flat_triangles_with_three_oblique_edges(pkg, orc, lattice):
    """A box-mesh-like scene (every triangle flat: the FLATONLY kernels, k_slow_leaves) whose triangles have three oblique
    in-plane edges -- all three edge axes unsettled at once, vertices on a lattice so that edges run through voxel corners."""
    from test_classify_host import _oblique_flat_triangles
    tris = _oblique_flat_triangles(90, 5 if lattice else 6, lattice)
    bbox = (np.zeros(3), np.ones(3))
    o = orc.OracleOctree(tris)
    o.build(8, 2, bbox=bbox)
    t = pkg.GeomOctree(tris)
    st = t.build(8, 2, bbox=bbox)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG (oblique flat triangles)")
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")


@pytest.mark.parametrize("levels,step", [(7, 4), (6, 4), (5, 2), (4, 1)], ids=["L7s4", "L6s4", "L5s2", "L4s1"])
@pytest.mark.parametrize("mesh,kw", [("city", dict(lots=8)), ("soup", dict(n=300, seed=3))], ids=["city", "soup"])
def test_shallow_sub_octrees(pkg, orc, meshgen, mesh, kw, levels, step):
    """Sub-octrees of one or two levels: the fused leaf kernels then run on the sub-octree's ROOT pairs (flags still unset when
    the level starts) or not at all."""
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    t = pkg.GeomOctree(tris)
    st = t.build(levels, step)
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), f"DAG ({mesh}, levels {levels}, step {step})")


def test_leaf_level_without_first_touches_is_exercised(pkg, orc, meshgen, monkeypatch, capfd):
    """A scene whose last batches still bring new voxel masks goes through all three leaf-level routes (tracked first
    touches, none needed, direct query for the nodes with a new mask); SVB_VX_STATS reports the query on stderr
    (this scene: 8 batches, 3 leaf nodes queried directly -- tools/gpu_sanitize_quick.sh)."""
    monkeypatch.setenv("SVB_VX_STATS", "1")
    tris = meshgen.make_mesh("city", lots=8)
    o = orc.OracleOctree(tris)
    o.build(9, 2)
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(6 << 20)
    st = t.build(9, 2)
    err = capfd.readouterr().err
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "batched DAG")
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")
    if "queried directly" not in err and "voxelizing again" not in err:
        pytest.skip(f"no batch of this scene met a new voxel mask after a quiet batch ({st['nBatches']} batches)")


@pytest.mark.parametrize("mesh,kw,levels,step,budget", [
    ("city", dict(lots=16), 10, 2, 16 << 20),
    ("terrain", dict(n=128), 10, 2, 32 << 20),
    ("sphere_menger", dict(n_lat=64, n_lon=128, sponge_level=2), 9, 1, 4 << 20),
], ids=["city", "terrain", "spongeball"])
def test_k64_level_without_first_touches_is_exercised(pkg, orc, meshgen, mesh, kw, levels, step, budget, monkeypatch, capfd):
    """Later tile batches are voxelized without first touches on the 4^3 level too; the nodes of entries that are new to the
    level's table are listed by the insert pass and get their first touch from a direct query of the tile's triangles
    (k_k64_query).  General triangles (terrain, sphere) bring new 4^3 patterns in every batch, so the query runs often;
    the result must be the oracle's node for node."""
    monkeypatch.setenv("SVB_VX_STATS", "1")
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(budget)
    st = t.build(levels, step)
    err = capfd.readouterr().err
    for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG"):
        assert st[k] == o.stat(k), k
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "batched DAG")
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")
    assert st["nBatches"] >= 3
    if "4^3 level without first touches" not in err:
        pytest.skip(f"no later batch of this scene brought a new 4^3 pattern ({st['nBatches']} batches)")


@pytest.mark.parametrize("mesh,kw,levels,step", [
    ("city", dict(lots=16), 10, 2),
    ("terrain", dict(n=128), 10, 2),
    ("sphere_menger", dict(n_lat=64, n_lon=128, sponge_level=2), 9, 1),
    ("soup", dict(n=400, seed=7), 8, 2),
    ("sphere", dict(n_lat=16, n_lon=32), 3, 0),      # the smallest octree the SSVDAG format accepts
], ids=["city", "terrain", "spongeball", "soup", "three-levels"])
def test_gpu_encoders_equal_host_encoders(pkg, orc, meshgen, mesh, kw, levels, step, monkeypatch):
    """svb_encode writes the formats on the GPU (svb_encode.cu); SVB_ENCODE=host routes the same call through the host
    encoders (levels D2H + csrc/host/encoders.cpp).  Same bytes for all five files, and both equal the oracle's."""
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    t = pkg.GeomOctree(tris)
    t.build(levels, step)
    got, host, want = {}, {}, {}

    def grab(kinds, extra=""):
        for k in kinds:
            monkeypatch.delenv("SVB_ENCODE", raising=False)
            got[k + extra] = pkg.encoders.encode(t, k)
            assert pkg.encoders.encode_copy(t, k) == got[k + extra]            # the copying form of the same call (cached image)
            monkeypatch.setenv("SVB_ENCODE", "host")
            host[k + extra] = pkg.encoders.encode(t, k)
            want[k + extra] = o.encode(k)
        monkeypatch.delenv("SVB_ENCODE", raising=False)

    grab(("svdag", "esvdag"))
    t.to_sdag()
    o.to_sdag()
    grab(("ussvdag", "ssvdag"))
    t2 = pkg.GeomOctree(tris)
    t2.build(levels, step)
    t2.cross_merge()
    o2 = orc.OracleOctree(tris)
    o2.build(levels, step)
    o2.cross_merge()
    monkeypatch.delenv("SVB_ENCODE", raising=False)
    got["multi"] = pkg.encoders.encode(t2, "svdag")
    want["multi"] = o2.encode("svdag")
    for k in want:
        assert got[k] == want[k], f"{k}: GPU encoder differs from the oracle"
        if k in host:
            assert host[k] == want[k], f"{k}: host encoder differs from the oracle"


def test_svbuilder_cli_multi_gpu_writes_reference_files(pkg, tmp_path):
    """`svbuilder m.obj L s --devices 0,1`: one process, one host thread + one libsvb context per GPU, NCCL all-gathers for
    the triangle soup and the per-level node records (csrc/host/sharded_build.cpp).  Same files as the reference, and
    the result block carries the reference's own lines."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    tool = pkg.lib_path().parent / "svbuilder"
    ndev = min(torch.cuda.device_count(), 4)
    for path in [p for p in GOLDEN if p.stem in ("sphere_L7_s1", "city_L7_s2", "terrain_L7_s3", "city_L7_s1_c")]:
        g = golden_case(path)
        d = tmp_path / g["name"]
        d.mkdir()
        pkg.meshgen.write_obj(d / "m.obj", g["tris"])
        cmd = [str(tool), str(d / "m.obj"), str(g["levels"]), str(g["step"])] + (["-c"] if g["cross"] else []) + ["--devices", ",".join(str(i) for i in range(ndev))]
        r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert f"{ndev} GPUs" in r.stdout and "memory (bytes)," in r.stdout and "SSVDAG / DAG" in r.stdout
        for ext, data in g["files"].items():
            name = f"m_{g['levels']}" + ("-multi.svdag" if ext == "multi_svdag" else "." + ext)
            assert (d / name).read_bytes() == data, f"{g['name']}: {name} differs from the reference's file"
        if path.stem == "terrain_L7_s3":
            # the same once more with the tags of every first hash seed truncated: the level merge collides on every rank alike and
            # all of them repeat the build under the next merge seed (csrc/host/sharded_build.cpp)
            import os
            r2 = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=300, env=dict(os.environ, SVB_TEST_WEAK_HASH="6"))
            assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
            for ext, data in g["files"].items():
                name = f"m_{g['levels']}" + ("-multi.svdag" if ext == "multi_svdag" else "." + ext)
                assert (d / name).read_bytes() == data, f"{g['name']}: {name} differs after the merge retry"


@pytest.mark.parametrize("mesh,kw,levels,step,cross", [
    ("city", dict(lots=8), 9, 2, False),
    ("sphere", dict(n_lat=32, n_lon=64), 8, 0, False),
    ("terrain", dict(n=48), 8, 1, True),
], ids=["city-s2", "sphere-s0", "terrain-s1-c"])
def test_svbuilder_cli_text_equals_reference_text(pkg, orc, meshgen, tmp_path, mesh, kw, levels, step, cross):
    """the log text is part of the tool's surface.  The result block and the ./stats.txt block of the drop-in tool
    against those of the unmodified reference binary run on the same input here: every line that does not carry a time must
    be identical (counts, human-readable sizes, bits/vox, ratios, the 'memory (bytes)' row)."""
    import subprocess
    if not orc.REF_BIN.exists():
        pytest.skip("oracle/_ref/svbuilder_ref not built")
    tool = pkg.lib_path().parent / "svbuilder"
    tris = meshgen.make_mesh(mesh, **kw)
    ref = orc.run_reference(tmp_path / "ref", tris, levels, step, cross=cross)
    d = tmp_path / "gpu"
    d.mkdir()
    meshgen.write_scene(d / "m.obj", tris)
    r = subprocess.run([str(tool), str(d / "m.obj"), str(levels), str(step)] + (["-c"] if cross else []), cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]

    def block(text):
        lines = text.splitlines()
        a = next(i for i, l in enumerate(lines) if l.startswith("========= RESULTS"))
        keep = [l for l in lines[a:] if l.strip() and " time " not in l and not l.startswith(("time (ms)", "GPU build", "Cleaning up"))]
        return [l.replace(str(tmp_path / "ref"), "").replace(str(d), "") for l in keep]

    assert block(r.stdout) == block(ref["log"])
    want = [l for l in (tmp_path / "ref" / "stats.txt").read_text().splitlines() if not l.startswith("time (ms)")]
    got = [l for l in (d / "stats.txt").read_text().splitlines() if not l.startswith("time (ms)")]
    assert got == want


def test_hash_collisions_are_retried_not_reported(pkg, orc, meshgen, monkeypatch):
    """Inner dedup levels, SDAG class keys and cross-level subtree ids are 64-bit tags verified by an exact pass; a detected
    collision re-runs the stage under another seed instead of failing the build.  SVB_TEST_WEAK_HASH truncates the tags of
    a context's FIRST seed to 6 bits, so the stage collides and must come out with the oracle's result on the retry."""
    tris = meshgen.make_mesh("terrain", n=64)
    o = orc.OracleOctree(tris)
    o.build(9, 2)
    # (1) the build itself
    monkeypatch.setenv("SVB_TEST_WEAK_HASH", "6")
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(8 << 20)
    st = t.build(9, 2)
    assert st["nHashRetries"] >= 1, "the weak first-attempt hash did not collide: the retry path was not exercised"
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG after a hash retry")
    assert pkg.encoders.encode(t, "svdag") == o.encode("svdag")
    # (2) cross-level merge and (3) toSDAG on contexts that are still on their first seed
    for stage in ("cross", "sdag"):
        monkeypatch.delenv("SVB_TEST_WEAK_HASH")
        u = pkg.GeomOctree(tris)
        assert u.build(9, 2)["nHashRetries"] == 0
        monkeypatch.setenv("SVB_TEST_WEAK_HASH", "6")
        oo = orc.OracleOctree(tris)
        oo.build(9, 2)
        if stage == "cross":
            cm = u.cross_merge()
            assert cm["nHashRetries"] >= 1
            assert cm["nCrossLevelMerged"] == oo.cross_merge()
            assert pkg.encoders.encode(u, "svdag") == oo.encode("svdag")
        else:
            sd = u.to_sdag()
            oo.to_sdag()
            assert sd["nHashRetries"] >= 1
            assert sd["nNodesSDAG"] == oo.stat("nNodesSDAG")
            assert pkg.encoders.encode(u, "ssvdag") == oo.encode("ssvdag")


def test_merge_collision_is_reported_to_every_rank_and_retried(pkg, meshgen, monkeypatch):
    """Multi-GPU level merge: a tag collision surfaces in svb_shard_finish on EVERY rank (identical records, identical seed),
    so all of them repeat the build under the next merge seed (what sharded.build_sharded does)."""
    tris = meshgen.make_mesh("terrain", n=64)
    ref = pkg.GeomOctree(tris)
    ref.build(9, 2)
    want = ref.levels_host()
    monkeypatch.setenv("SVB_TEST_WEAK_HASH", "6")
    with pytest.raises(pkg.SvbError) as ei:
        _simulate_ranks(pkg, tris, 9, 2, 3, merge_seed=0)
    assert ei.value.code == -5
    octs, _ = _simulate_ranks(pkg, tris, 9, 2, 3, merge_seed=1)
    for r, oc in enumerate(octs):
        _assert_levels_equal(oc.levels_host(), want, f"rank {r} after the merge retry")


def test_context_on_a_caller_owned_stream(pkg, orc, meshgen):
    """svb_create_on_stream: the context enqueues everything on a stream the host owns (here a torch stream, as bench.py does
    for the multi-rank build); same octree, and the stream is still usable after the context is gone."""
    import torch
    tris = meshgen.make_mesh("city", lots=8)
    o = orc.OracleOctree(tris)
    o.build(9, 2)
    ts = torch.cuda.Stream()
    t = pkg.GeomOctree(tris, stream=ts)
    assert t.stream_ptr() == ts.cuda_stream
    st = t.build(9, 2)
    assert st["nNodesDAG"] == o.stat("nNodesDAG")
    _assert_levels_equal(t.levels_host(), _oracle_levels(o), "DAG built on a torch stream")
    t.to_sdag()
    o.to_sdag()
    assert pkg.encoders.encode(t, "ssvdag") == o.encode("ssvdag")
    t.close()
    with torch.cuda.stream(ts):
        x = torch.ones(1024, device="cuda").sum()
    ts.synchronize()
    assert float(x) == 1024.0
