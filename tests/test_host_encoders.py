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
