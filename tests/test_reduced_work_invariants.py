"""CPU checks (numpy model of the parallel formulation, oracle/parallel_model.py) of the invariants the reduced-work
paths of DESIGN.md §3c rest on:

  * star stores: a child's first touch is never smaller than its parent's, and equals it exactly when the parent's
    first-touch triangle reaches the child -- so the children of a node's own first-touch pair can be stored plainly;
  * leaf query: a node's first touch is the smallest candidate whose triangle passes testTriBox at every box on the
    node's path (what k_leaf_query recomputes for the leaf nodes with a new voxel mask);
  * frozen entries: with sub-octrees reduced in ascending sequence order, an entry created by an earlier sub-octree is
    never improved by a later one."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from oracle import parallel_model as pm

ROOT = Path(__file__).resolve().parents[1]
spec = importlib.util.spec_from_file_location("_meshgen_inv", ROOT / "svdag-compression_b200" / "meshgen.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

MESHES = [("sphere", dict(n_lat=16, n_lon=32), 5), ("city", dict(lots=4), 6), ("soup", dict(n=150, seed=7), 5)]


def _tile(mesh, kw, levels):
    tris = mg.make_mesh(mesh, **kw).reshape(-1, 9)
    v = tris.reshape(-1, 3)
    lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
    side, _, _ = pm._float_root_side(lo, hi)
    return tris, pm.voxelize_tile(tris, np.arange(len(tris)), (lo + hi) * 0.5, side, levels), side


@pytest.mark.parametrize("mesh,kw,levels", MESHES, ids=[m[0] for m in MESHES])
def test_first_touch_of_children(mesh, kw, levels):
    tris, lv, side = _tile(mesh, kw, levels)
    for l in range(levels - 1):
        par, kid = lv[l], lv[l + 1]
        pc = np.array([bin(x).count("1") for x in range(256)])[par["mask"]]
        parent_of = np.repeat(np.arange(len(par["mask"])), pc)
        assert len(parent_of) == len(kid["tstar"])
        tp = par["tstar"][parent_of]
        assert (kid["tstar"] >= tp).all()
        # child boxes: centre kid["centre"], half side = the k of the step that created them
        k = side / float(1 << (l + 2))
        reached = pm.tri_box(kid["centre"], k, tris[tp])
        assert np.array_equal(kid["tstar"] == tp, reached), f"level {l + 1}"
        # exactly one pair per node carries the node's own first touch: first touches are triangle ids, pairs of a node
        # belong to distinct triangles (implicit in the model: pairs are (triangle, node) with unique triangles per node)


@pytest.mark.parametrize("mesh,kw,levels", MESHES, ids=[m[0] for m in MESHES])
def test_leaf_first_touch_by_direct_query(mesh, kw, levels):
    tris, lv, side = _tile(mesh, kw, levels)
    leaf = lv[levels - 1]
    rng = np.random.default_rng(1)
    pick = rng.choice(len(leaf["path"]), size=min(40, len(leaf["path"])), replace=False)
    c0 = lv[0]["centre"][0]
    for n in pick:
        path = int(leaf["path"][n])
        digits = [(path >> (3 * d)) & 7 for d in range(levels - 2, -1, -1)]   # levels-1 digits, root first
        alive = np.ones(len(tris), dtype=bool)
        c = c0.copy()
        for j, dig in enumerate(digits):
            k = side / float(1 << (j + 2))
            c = c + np.array([k if dig & 4 else -k, k if dig & 2 else -k, k if dig & 1 else -k])
            idx = np.nonzero(alive)[0]
            ok = pm.tri_box(np.broadcast_to(c, (len(idx), 3)), k, tris[idx])
            alive[idx[~ok]] = False
        assert np.allclose(c, leaf["centre"][n], rtol=0, atol=0)
        assert int(np.nonzero(alive)[0][0]) == int(leaf["tstar"][n])


@pytest.mark.parametrize("mesh,kw,L,step", [("sphere", dict(n_lat=24, n_lon=48), 6, 1), ("city", dict(lots=4), 7, 2), ("terrain", dict(n=24), 6, 2)],
                         ids=["sphere", "city", "terrain"])
def test_later_sub_octrees_never_improve_an_entry(mesh, kw, L, step, monkeypatch):
    tris = mg.make_mesh(mesh, **kw)
    late = []
    orig = pm.Builder._dedup_tile

    def spy(self, lv, tile_seq, lvl0):
        before = [dict(t) for t in self.tables]
        out = orig(self, lv, tile_seq, lvl0)
        for g, (b, a) in enumerate(zip(before, self.tables)):
            late.extend((g, key) for key, o in b.items() if a[key] != o)
        return out

    monkeypatch.setattr(pm.Builder, "_dedup_tile", spy)
    b = pm.Builder(tris)
    b.shard_build(L, step, None)
    assert b.ntiles > 1
    assert not late, f"{len(late)} entries were improved by a later sub-octree, e.g. {late[:3]}"
