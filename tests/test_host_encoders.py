"""Product host encoders (csrc/host/encoders.cpp, via svb_encode_levels) against the golden files:
the oracle supplies the node arrays, the product writes the bytes.  CPU only (no device needed)."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case


def _files_from_oracle_levels(pkg, orc, g):
    enc = pkg.encoders
    o = orc.OracleOctree(g["tris"])
    o.build(g["levels"], g["step"])
    lo, hi = o.scene_bbox()
    bboxF = np.concatenate([lo, hi]).astype(np.float32)
    rs = orc.lib().orc_root_side(o.h)
    out = {}
    lv = [o.level(l) for l in range(o.levels)]
    out["svdag"] = enc.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 2, "svdag")
    if g["cross"]:
        o.cross_merge()
        lv = [o.level(l) for l in range(o.levels)]
        out["multi_svdag"] = enc.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 2, "svdag")
        return out
    out["esvdag"] = enc.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 2, "esvdag")
    o.to_sdag()
    lv = [o.level(l) for l in range(o.levels)]
    out["ussvdag"] = enc.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 3, "ussvdag")
    out["ssvdag"] = enc.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 3, "ssvdag")
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_product_encoders_write_reference_bytes(pkg, orc, path):
    g = golden_case(path)
    mine = _files_from_oracle_levels(pkg, orc, g)
    assert set(mine) == set(g["files"])
    for k, data in mine.items():
        assert data == g["files"][k], f"{g['name']}: product encoder output for {k} differs from the reference file"


def test_encoder_state_checks(pkg):
    lv = [{"mask": np.array([1], np.uint8), "child": np.full((1, 8), 0xFFFFFFFE, np.uint32)}] * 3
    with pytest.raises(RuntimeError):
        pkg.encoders.encode_levels(lv, np.zeros(6, np.float32), 1.0, 3, 2, "ussvdag")   # USSVDAG needs SDAG state
    with pytest.raises(RuntimeError):
        pkg.encoders.encode_levels(lv, np.zeros(6, np.float32), 1.0, 3, 3, "svdag")     # SVDAG needs DAG state


@pytest.mark.parametrize("path", [p for p in GOLDEN if not p.stem.endswith("_c")], ids=lambda p: p.stem)
def test_svdag_decode_round_trip(pkg, orc, path):
    """EncodedSVDAG::load + decode (encoded_svdag.cpp:43-74, :200-270) in the product's host code: the reference's
    .svdag file decodes to the oracle's DAG levels, and re-encodes to the same bytes (what `svbuilder m.svdag L 0`
    relies on, main.cpp:96-101)."""
    g = golden_case(path)
    data = g["files"]["svdag"]
    levels, bboxF, rs, nn = pkg.encoders.decode_svdag(data)
    o = orc.OracleOctree(g["tris"])
    o.build(g["levels"], g["step"])
    assert len(levels) == o.levels
    for l in range(o.levels):
        want = o.level(l)
        assert np.array_equal(levels[l]["mask"], want["mask"]), l
        if l + 1 < o.levels:
            assert np.array_equal(levels[l]["child"], want["child"]), l
    assert pkg.encoders.encode_levels(levels, bboxF, rs, nn, 2, "svdag") == data
    assert pkg.encoders.encode_levels(levels, bboxF, rs, nn, 2, "esvdag") == g["files"]["esvdag"]


def test_svdag_decode_rejects_garbage(pkg):
    with pytest.raises(RuntimeError):
        pkg.encoders.decode_svdag(b"\x00" * 20)
    with pytest.raises(RuntimeError):
        pkg.encoders.decode_svdag(b"\xff" * 200)


def test_threaded_encoder_loops_on_a_larger_dag(pkg, orc, meshgen):
    """Levels of several thousand nodes (the golden scenes stay below the threshold of the OpenMP loops and of the
    concurrent per-level sorts of the SSVDAG encoder): the product encoders (default thread count) against the oracle's
    restatement of the reference encoders, DAG and SDAG state."""
    tris = meshgen.make_mesh("city", lots=24)
    o = orc.OracleOctree(tris)
    o.build(10, 2)
    lo, hi = o.scene_bbox()
    bboxF = np.concatenate([lo, hi]).astype(np.float32)
    rs = orc.lib().orc_root_side(o.h)
    lv = [o.level(l) for l in range(o.levels)]
    assert max(len(l["mask"]) for l in lv) > 8192
    want = {k: o.encode(k) for k in ("svdag", "esvdag")}
    got = {k: pkg.encoders.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 2, k) for k in ("svdag", "esvdag")}
    o.to_sdag()
    lv = [o.level(l) for l in range(o.levels)]
    for k in ("ussvdag", "ssvdag"):
        want[k] = o.encode(k)
        got[k] = pkg.encoders.encode_levels(lv, bboxF, rs, o.stat("nNodes"), 3, k)
    for k in want:
        assert got[k] == want[k], k


def _std_sort_order(orc, refs):
    """Plain std::sort with the reference's comparator (oracle/svdag_oracle.cpp::orc_sort_by_refs)."""
    import ctypes as C
    L = orc.lib()
    L.orc_sort_by_refs.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    refs = np.ascontiguousarray(refs, dtype=np.uint32)
    out = np.zeros(len(refs), np.uint32)
    L.orc_sort_by_refs(len(refs), refs.ctypes.data, out.ctypes.data)
    return out


def test_parallel_ref_sort_reproduces_std_sort_tie_order(pkg, orc):
    """The SSVDAG node order is the outcome of the reference's UNSTABLE std::sort (encoded_ssvdag.cpp:273-275): the product
    runs libstdc++'s partition / introsort routines with the independent halves as parallel tasks
    (csrc/host/encoders.cpp::sort_by_refs) and must land on the same permutation, ties included, for every input --
    sizes around the task threshold (16384) and far above it, heavy ties, sorted / reversed / constant / organ-pipe inputs."""
    rng = np.random.default_rng(5)
    cases = []
    for n in (0, 1, 2, 15, 16, 17, 100, 4097, 16384, 16385, 50_000, 300_000, 1_000_000):
        cases.append(rng.integers(1, 4, n))                       # almost everything ties (the usual reference-count picture)
        cases.append(rng.zipf(1.6, n).clip(1, 10**6))             # a few hot nodes, a long tail of 1s
    n = 200_000
    cases += [np.arange(n), np.arange(n)[::-1], np.full(n, 7), np.minimum(np.arange(n), np.arange(n)[::-1]),
              rng.integers(0, 2**31, n), np.repeat(np.arange(n // 100)[::-1], 100)]
    levels = [np.zeros(1, np.uint32)] + [np.asarray(c, dtype=np.uint32) for c in cases]
    got = pkg.encoders.ssvdag_order_from_refs(levels)
    assert list(got[0]) == [0]
    for k, (refs, g) in enumerate(zip(levels[1:], got[1:])):
        want = _std_sort_order(orc, refs)
        assert np.array_equal(g, want), f"case {k} (n={len(refs)}): permutation differs from std::sort's"
