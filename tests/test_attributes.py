"""Material-id leaves (GeomOctree::buildSVO(..., putMaterialIdInLeaves = true), geom_octree.cpp:210-211, :252) and the
self-specified Gray-coded attribute bit-trees.

Goldens tests/golden/attr_*.npz hold the leaf level of the UNMODIFIED reference (minted through oracle/ref_attr_driver.cpp by
tests/golden/make_golden_attr.py).  CPU: the oracle restatement reproduces them.  GPU: svb_build_svo_materials reproduces
them and the oracle on further scenes; svb_attribute_bit_trees is checked against a plain numpy / dict model of the
specification."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = sorted((Path(__file__).resolve().parent / "golden").glob("attr_*.npz"))
NULL = 0xFFFFFFFE


def _case(path, meshgen):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    tris = meshgen.make_mesh(meta["mesh"], **meta["kw"])
    assert len(tris) == meta["triangles"]
    mats = meshgen.materials_for(tris, **meta["materials"])
    return tris, mats, int(z["levels"]), z["mask"], z["material"]


@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_oracle_material_leaves_equal_reference(orc, meshgen, path):
    tris, mats, levels, mask, material = _case(path, meshgen)
    o = orc.OracleOctree(tris)
    om, oc = o.build_svo_materials(levels, mats)
    assert np.array_equal(om, mask)
    assert np.array_equal(oc, material)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_gpu_material_leaves_equal_reference(pkg, meshgen, path):
    """Same leaf nodes in the same order, same masks, same 8 child slots as the unmodified reference."""
    tris, mats, levels, mask, material = _case(path, meshgen)
    t = pkg.GeomOctree(tris)
    gm, gc = t.build_svo_materials(levels, mats)
    assert len(gm) == len(mask)
    assert np.array_equal(gm, mask)
    assert np.array_equal(gc, material)


def _bit_tree_model(levels_host, leaf_mask, leaf_mat, nbits, gray):
    """The specification of svb_attribute_bit_trees in plain Python: per bit the voxel subset, reduced bottom-up by interning
    (child ids) per level; node count = unique nodes of levels >= 1 + the root, 0 for an empty tree."""
    L = len(levels_host)
    code = leaf_mat.astype(np.uint64)
    code = np.where(leaf_mat == NULL, 0, code)
    if gray:
        code = code ^ (code >> np.uint64(1))
    set_bits = ((leaf_mask[:, None] >> np.arange(8)) & 1).astype(bool) & (leaf_mat != NULL)
    nodes, voxels = [], []
    for b in range(nbits):
        on = set_bits & (((code >> np.uint64(b)) & np.uint64(1)) == 1)
        pm = (on * (1 << np.arange(8))).sum(axis=1).astype(np.int64)
        voxels.append(int(on.sum()))
        ids = [None if m == 0 else int(m) for m in pm]          # leaf level: a node IS its mask
        total = len({i for i in ids if i is not None})
        for l in range(L - 2, 0, -1):
            ch = levels_host[l]["child"]
            table, cur = {}, []
            for row in ch:
                key = tuple(None if c == NULL else ids[c] for c in row)
                if all(k is None for k in key):
                    cur.append(None)
                else:
                    cur.append(table.setdefault(key, len(table)))
            total += len(table)
            ids = cur
        nodes.append(total + 1 if voxels[-1] else 0)
    return np.array(nodes, np.uint64), np.array(voxels, np.uint64)


@pytest.mark.gpu
@pytest.mark.parametrize("mesh,kw,levels,nmat", [
    ("city", dict(lots=4), 7, 29),
    ("sphere", dict(n_lat=24, n_lon=48), 6, 13),
    ("soup", dict(n=300, seed=5), 6, 100),
], ids=["city", "sphere", "soup"])
def test_gpu_attribute_bit_trees_match_the_specification(pkg, orc, meshgen, mesh, kw, levels, nmat):
    tris = meshgen.make_mesh(mesh, **kw)
    mats = meshgen.materials_for(tris, n_materials=nmat, seed=17)
    o = orc.OracleOctree(tris)
    om, oc = o.build_svo_materials(levels, mats)                    # leaves in SVO order with material ids (pinned above)
    svo = [o.level(l) for l in range(levels)]                       # the SVO of the same call (child indices, nullNode = none)
    t = pkg.GeomOctree(tris)
    gm, gc = t.build_svo_materials(levels, mats)
    assert np.array_equal(gm, om) and np.array_equal(gc, oc)
    nbits = int(nmat - 1).bit_length()
    for gray in (False, True):
        nodes, vox = t.attribute_bit_trees(nbits, gray)
        wn, wv = _bit_tree_model(svo, om, oc, nbits, gray)
        assert np.array_equal(vox, wv), (gray, vox, wv)
        assert np.array_equal(nodes, wn), (gray, nodes, wn)
    # every voxel with a non-zero code sits in at least one bit-tree; Gray coding is a bijection on the ids
    assert pkg.lib().svb_gray_code(5) == 7 and pkg.lib().svb_gray_code(0) == 0
    a = np.arange(1 << nbits, dtype=np.uint32)
    g = a ^ (a >> 1)
    assert len(np.unique(g)) == len(a) and (np.bitwise_count(g[1:] ^ g[:-1]) == 1).all()
