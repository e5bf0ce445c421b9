"""Data-parallel (numpy) model of the GPU formulation of the svbuilder hot path.

TEST INFRASTRUCTURE ONLY (same rules as oracle/svdag_oracle.cpp).  This is *not* the
sequential reference algorithm: it is the level-synchronous / order-key / orbit-min
formulation the CUDA kernels implement, written with whole-array numpy ops
so that the formulation itself can be pinned against the sequential restatement and the
golden files on CPU, and so that GPU intermediates have something to be diffed against.

Formulation (reference lines it must reproduce in brackets):
  * voxelize: level-synchronous expansion of (triangle, node) pairs; a node exists iff some
    triangle passed the SAT test at every ancestor [geom_octree.cpp:205-261];
    first-touch triangle t*(node) = min triangle id over its pairs.
  * SVO order of a level inside one (sub-)octree = rank of (t*, parent path ascending,
    7 - childIdx) [children are created 7->0 but visited 0->7, :234,:249-250].
  * toDAG: unique subtrees, ordered by first occurrence in (tile sequence, SVO order)
    [:462-548 applied per sub-octree then globally, :289-435].
  * toSDAG: orbit-min over the 8 mirror variants [:551-697].
"""
from __future__ import annotations

import numpy as np

NULL = np.uint32(0xFFFFFFFE)


# --------------------------------------------------------------------------- SAT (vectorised)
def tri_box(c, h, tri):
    """test_triangle_box.cpp:105-184 for P (centre, triangle) pairs at once.
    c (P,3) float64, h scalar/array float64, tri (P,9) float32 -> bool (P,)"""
    t = tri.astype(np.float64)
    v0 = t[:, 0:3] - c
    v1 = t[:, 3:6] - c
    v2 = t[:, 6:9] - c
    e0 = v1 - v0
    e1 = v2 - v1
    e2 = v0 - v2
    ok = np.ones(len(t), dtype=bool)

    def rej(pa, pb, rad):
        mn = np.where(pa < pb, pa, pb)
        mx = np.where(pa < pb, pb, pa)
        return (mn > rad) | (mx < -rad)

    X, Y, Z = 0, 1, 2
    for e, tests in ((e0, ("X01", "Y02", "Z12")), (e1, ("X01", "Y02", "Z0")), (e2, ("X2", "Y1", "Z12"))):
        fex, fey, fez = np.abs(e[:, X]), np.abs(e[:, Y]), np.abs(e[:, Z])
        for name in tests:
            if name[0] == "X":
                a, b, rad = e[:, Z], e[:, Y], (fez + fey) * h
                f = lambda v: a * v[:, Y] - b * v[:, Z]
            elif name[0] == "Y":
                a, b, rad = e[:, Z], e[:, X], (fez + fex) * h
                f = lambda v: -a * v[:, X] + b * v[:, Z]
            else:
                a, b, rad = e[:, Y], e[:, X], (fey + fex) * h
                f = lambda v: a * v[:, X] - b * v[:, Y]
            pair = {"X01": (v0, v2), "X2": (v0, v1), "Y02": (v0, v2), "Y1": (v0, v1), "Z12": (v1, v2), "Z0": (v0, v1)}[name]
            ok &= ~rej(f(pair[0]), f(pair[1]), rad)
    for k in range(3):
        mn = np.minimum(np.minimum(v0[:, k], v1[:, k]), v2[:, k])
        mx = np.maximum(np.maximum(v0[:, k], v1[:, k]), v2[:, k])
        ok &= ~((mn > h) | (mx < -h))
    n = np.stack([e0[:, 1] * e1[:, 2] - e0[:, 2] * e1[:, 1],
                  e0[:, 2] * e1[:, 0] - e0[:, 0] * e1[:, 2],
                  e0[:, 0] * e1[:, 1] - e0[:, 1] * e1[:, 0]], axis=1)
    pos = n > 0.0
    hh = np.broadcast_to(np.asarray(h, dtype=np.float64).reshape(-1, 1) if np.ndim(h) else np.float64(h), v0.shape)
    vmin = np.where(pos, -hh - v0, hh - v0)
    vmax = np.where(pos, hh - v0, -hh - v0)
    d0 = n[:, 0] * vmin[:, 0]
    d0 = d0 + n[:, 1] * vmin[:, 1]
    d0 = d0 + n[:, 2] * vmin[:, 2]
    d1 = n[:, 0] * vmax[:, 0]
    d1 = d1 + n[:, 1] * vmax[:, 1]
    d1 = d1 + n[:, 2] * vmax[:, 2]
    ok &= ~(d0 > 0.0) & (d1 >= 0.0)
    return ok


# --------------------------------------------------------------------------- voxelizer
def _float_root_side(lo, hi):
    """geom_octree.cpp:177-184: side derived from the float-converted bbox."""
    lf = lo.astype(np.float32)
    hf = hi.astype(np.float32)
    sides = ((hf - lf) * np.float32(0.5)) * np.float32(2.0)
    return float(sides.max()), lf, hf


def voxelize_tile(tris, cand, centre, root_side, levels):
    """Level-synchronous build of one (sub-)octree of `levels` levels.

    Returns per level l: dict(path uint64 (octal digits, root = 0 digits), tstar, mask, childBase)
    with nodes in Morton (path) order; level 0 is the root."""
    tris = tris.reshape(-1, 9)
    lv = [{"path": np.zeros(1, np.uint64), "tstar": np.zeros(1, np.int64), "centre": np.asarray(centre, np.float64).reshape(1, 3)}]
    pt = np.asarray(cand, dtype=np.int64)          # pair triangle
    pn = np.zeros(len(pt), dtype=np.int64)         # pair node index
    for l in range(levels):
        nodes = lv[l]
        k = root_side / float(1 << (l + 2))        # getHalfSideD(l + 1)
        hit = np.zeros((len(pt), 8), dtype=bool)
        for c in range(8):
            off = np.array([k if c & 4 else -k, k if c & 2 else -k, k if c & 1 else -k])
            hit[:, c] = tri_box(nodes["centre"][pn] + off, k, tris[pt])
        mask = np.zeros(len(nodes["path"]), dtype=np.uint8)
        m8 = (hit * (1 << np.arange(8))).sum(axis=1).astype(np.uint8)
        np.bitwise_or.at(mask, pn, m8)
        nodes["mask"] = mask
        if l == levels - 1:
            break
        pc = np.array([bin(x).count("1") for x in range(256)], dtype=np.int64)[mask]
        base = np.concatenate([[0], np.cumsum(pc)])[:-1]
        nodes["childBase"] = base
        n_next = int(pc.sum())
        # children in (parent, c ascending) order == Morton order
        par = np.repeat(np.arange(len(mask)), pc)
        cidx = np.concatenate([np.nonzero((mask[i] >> np.arange(8)) & 1)[0] for i in range(len(mask))]) if n_next else np.zeros(0, np.int64)
        off = np.stack([np.where(cidx & 4, k, -k), np.where(cidx & 2, k, -k), np.where(cidx & 1, k, -k)], axis=1)
        nxt = {"path": (nodes["path"][par] << np.uint64(3)) | cidx.astype(np.uint64),
               "centre": nodes["centre"][par] + off, "tstar": np.full(n_next, np.iinfo(np.int64).max)}
        # emit child pairs
        pi, ci = np.nonzero(hit)
        below = (mask[pn[pi]].astype(np.int64) & ((1 << ci) - 1))
        rank = np.array([bin(x).count("1") for x in range(256)], dtype=np.int64)[below]
        child = base[pn[pi]] + rank
        np.minimum.at(nxt["tstar"], child, pt[pi])
        pt, pn = pt[pi], child
        lv.append(nxt)
    return lv


def clean_and_effective_masks(lv):
    """cleanEmptyNodes (geom_octree.cpp:437-456): bottom-up unlink of mask-0 nodes."""
    L = len(lv)
    eff = [None] * L
    eff[L - 1] = lv[L - 1]["mask"].copy()
    for l in range(L - 2, -1, -1):
        m = lv[l]["mask"].copy()
        base = lv[l]["childBase"]
        for i in range(len(m)):
            r = 0
            for c in range(8):
                if (m[i] >> c) & 1:
                    if eff[l + 1][base[i] + r] == 0:
                        m[i] &= ~(1 << c) & 0xFF
                    r += 1
        eff[l] = m
    return eff


def order_key(lv, l, tile_seq=0):
    """(tile_seq, t*, path') with path' = parent path, then 7 - childIdx."""
    p = lv[l]["path"].astype(np.uint64)
    if l > 0:
        p = (p & ~np.uint64(7)) | (np.uint64(7) - (p & np.uint64(7)))
    return np.stack([np.full(len(p), tile_seq, dtype=np.uint64), lv[l]["tstar"].astype(np.uint64), p], axis=1)


def _lexrank(keys):
    """rank of each row of (n,k) uint64 keys in lexicographic order (rows unique)."""
    order = np.lexsort(keys.T[::-1])
    r = np.empty(len(order), dtype=np.int64)
    r[order] = np.arange(len(order))
    return r, order


class Builder:
    """mesh -> DAG levels, parallel formulation.  levels[l] = dict(mask (U,), child (U,8)).

    Also speaks the multi-GPU protocol of include/svb.h (shard_* methods) on CPU memory, so that
    svdag-compression_b200/sharded.py can be exercised over gloo: keys here are structural (nested
    tuples), records are pickled dicts."""

    def __init__(self, tris):
        self.tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        self.stats = {}

    # ------------------------------------------------------------------ local phase
    def _dedup_tile(self, lv, tile_seq, lvl0):
        """bottom-up insert of one tile's levels into the tables; returns per-level keys."""
        Lt = len(lv)
        eff = clean_and_effective_masks(lv)
        ids = [None] * Lt
        for l in range(Lt - 1, 0 if lvl0 == 0 else -1, -1):
            ok = order_key(lv, l, tile_seq)
            keys = []
            for i in range(len(eff[l])):
                if eff[l][i] == 0:
                    keys.append(None)
                    continue
                if l == Lt - 1:
                    key = (int(eff[l][i]),) + (None,) * 8
                else:
                    ch, r = [], 0
                    for c in range(8):
                        if (lv[l]["mask"][i] >> c) & 1:
                            ch.append(ids[l + 1][lv[l]["childBase"][i] + r])
                            r += 1
                        else:
                            ch.append(None)
                    key = (int(eff[l][i]),) + tuple(ch)
                keys.append(key)
                o = tuple(int(x) for x in ok[i])
                tab = self.tables[lvl0 + l]
                if key not in tab or o < tab[key]:
                    tab[key] = o
            ids[l] = keys
        return ids, eff

    def shard_build(self, L, step, bbox, rank=0, world=1):
        tris = self.tris
        v = tris.reshape(-1, 3)
        lo, hi = bbox if bbox is not None else (v.min(0), v.max(0))
        lo = np.asarray(lo, np.float64)
        hi = np.asarray(hi, np.float64)
        all_t = np.arange(len(tris))
        root_side, lf, hf = _float_root_side(lo, hi)
        self.L, self.step, self.rank, self.world = L, step, rank, world
        self.root_side, self.bboxF = root_side, (lf, hf)
        self.tables = [dict() for _ in range(L)]
        self.n_svo = 0
        self.leaf_vox = 0
        self.top = None
        centre = (lo + hi) * 0.5
        if step == 0:
            assert world == 1
            lv = voxelize_tile(tris, all_t, centre, root_side, L)
            self.n_svo = sum(len(x["path"]) for x in lv[1:])
            ids, eff = self._dedup_tile(lv, 0, 0)
            self.leaf_vox = int(sum(bin(int(m)).count("1") for m in eff[L - 1]))
            self.root_children = (lv[0], ids, eff)
            self.ntiles = 1
            return
        s1 = step + 1
        base = voxelize_tile(tris, all_t, centre, root_side, s1)
        beff = clean_and_effective_masks(base)
        self.base_svo = sum(len(x["path"]) for x in base[1:])
        rank_, order = _lexrank(order_key(base, s1 - 1)) if s1 > 1 else (np.zeros(1, np.int64), np.zeros(1, np.int64))
        lhs = root_side / float(1 << s1)      # getHalfSideD(stepLevels - 1)
        self.base_vox = int(sum(bin(int(m)).count("1") for m in beff[s1 - 1]))
        self.tile_of = {}                      # (base leaf index, child) -> tile_seq
        self.tile_root_key = {}                # tile_seq -> key (only tiles built here, until the roots are exchanged)
        seq = 0
        for i in order:
            for j in range(7, -1, -1):
                if not (beff[s1 - 1][i] >> j) & 1:
                    continue
                self.tile_of[(int(i), j)] = seq
                if seq % world == rank:
                    p1 = base[s1 - 1]["centre"][i]
                    p2 = p1 + np.array([lhs if j & 4 else -lhs, lhs if j & 2 else -lhs, lhs if j & 1 else -lhs])
                    tlo, thi = np.minimum(p1, p2), np.maximum(p1, p2)
                    t_side, _, _ = _float_root_side(tlo, thi)
                    lv = voxelize_tile(tris, all_t, (tlo + thi) * 0.5, t_side, L - s1)
                    self.n_svo += sum(len(x["path"]) for x in lv[1:])
                    ids, eff = self._dedup_tile(lv, seq, s1)
                    self.leaf_vox += int(sum(bin(int(m)).count("1") for m in eff[-1]))
                    self.tile_root_key[seq] = ids[0][0]
                seq += 1
        self.ntiles = seq
        self.top = (base, beff, s1)

    # ------------------------------------------------------------------ exchange (CPU memory)
    def shard_info(self):
        s1 = self.step + 1 if self.step else 0
        return s1, self.L - 1, self.ntiles, [self.leaf_vox, self.n_svo, 0, 0, 0]

    def shard_level_count(self, g):
        import pickle
        if not hasattr(self, "_blobs"):
            self._blobs = {}
        self._blobs[g] = pickle.dumps(self.tables[g])   # structural keys: final as soon as the local phase ends
        return len(self._blobs[g]), 1

    def shard_export_level(self, g, ptr):
        import ctypes
        ctypes.memmove(ptr, self._blobs[g], len(self._blobs[g]))

    def shard_import_level(self, g, ptr, counts, stride):
        import ctypes
        import pickle
        merged = {}
        for r, n in enumerate(counts):
            blob = ctypes.string_at(ptr + r * int(stride), int(n))
            for key, o in pickle.loads(blob).items():
                if key not in merged or o < merged[key]:
                    merged[key] = o
        self.tables[g] = merged

    def _root_uids(self):
        s1 = self.step + 1
        items = sorted(self.tables[s1].items(), key=lambda kv: kv[1])
        return {key: u for u, (key, _) in enumerate(items)}, [key for key, _ in items]

    def shard_export_roots(self, ptr):
        uid, _ = self._root_uids()
        a = np.full(max(self.ntiles, 1), 0xFFFFFFFF, dtype=np.uint32)
        for seq, key in self.tile_root_key.items():
            a[seq] = 0xFFFFFFFE if key is None else uid[key]
        import ctypes
        ctypes.memmove(ptr, a.ctypes.data, a.nbytes)

    def shard_import_roots(self, ptr):
        import ctypes
        n = max(self.ntiles, 1)
        a = np.frombuffer(ctypes.string_at(ptr, 4 * n * self.world), dtype=np.uint32).reshape(self.world, n).min(axis=0)
        _, keys = self._root_uids()
        self.tile_root_key = {seq: (None if a[seq] == 0xFFFFFFFE else keys[a[seq]]) for seq in range(self.ntiles)}

    # ------------------------------------------------------------------ finish
    def shard_finish(self, totals=None):
        L, tables = self.L, self.tables
        leaf_vox, n_svo = (self.leaf_vox, self.n_svo) if totals is None else (int(totals[0]), int(totals[1]))
        final = [dict() for _ in range(L)]
        for l in range(L - 1, 0, -1):
            items = sorted(tables[l].items(), key=lambda kv: kv[1])
            for r, (key, _) in enumerate(items):
                final[l][key] = r
        levels_out = [None] * L
        if self.top is not None:
            base, beff, s1 = self.top
            self.stats["nTotalVoxels"] = self.base_vox + leaf_vox - self.ntiles
            self.stats["nNodesSVO"] = self.base_svo + n_svo
            ids_top = [None] * s1
            for l in range(s1 - 1, -1, -1):
                keys = []
                for i in range(len(beff[l])):
                    if beff[l][i] == 0:
                        keys.append(None)
                        continue
                    ch = []
                    if l == s1 - 1:
                        for c in range(8):
                            if (beff[l][i] >> c) & 1:
                                k = self.tile_root_key[self.tile_of[(i, c)]]
                                ch.append(("id", final[s1][k]) if k is not None else ("id", 0))
                            else:
                                ch.append(None)
                    else:
                        r = 0
                        for c in range(8):
                            if (base[l]["mask"][i] >> c) & 1:
                                ch.append(ids_top[l + 1][base[l]["childBase"][i] + r])
                                r += 1
                            else:
                                ch.append(None)
                    key = (int(beff[l][i]),) + tuple(ch)
                    keys.append(key)
                    if l > 0:
                        o = tuple(int(x) for x in order_key(base, l)[i])
                        if key not in tables[l] or o < tables[l][key]:
                            tables[l][key] = o
                if l > 0:
                    items = sorted(tables[l].items(), key=lambda kv: kv[1])
                    final[l] = {key: r for r, (key, _) in enumerate(items)}
                    ids_top[l] = [None if k is None else ("id", final[l][k]) for k in keys]
                else:
                    ids_top[l] = keys
            root_key = ids_top[0][0]
        else:
            self.stats["nTotalVoxels"] = leaf_vox
            self.stats["nNodesSVO"] = n_svo
            lv0, ids, eff = self.root_children
            ch, r = [], 0
            for c in range(8):
                if (lv0["mask"][0] >> c) & 1:
                    ch.append(ids[1][lv0["childBase"][0] + r])
                    r += 1
                else:
                    ch.append(None)
            root_key = (int(eff[0][0]),) + tuple(ch)

        def resolve(l, key):
            out = np.full(8, NULL, dtype=np.uint32)
            for c in range(8):
                k = key[1 + c]
                if k is None:
                    continue
                out[c] = k[1] if (isinstance(k, tuple) and len(k) == 2 and k[0] == "id") else final[l + 1][k]
            return out

        for l in range(1, L):
            U = len(final[l])
            mask = np.zeros(U, np.uint8)
            child = np.full((U, 8), NULL, dtype=np.uint32)
            for key, r in final[l].items():
                mask[r] = key[0]
                if l < L - 1:
                    child[r] = resolve(l, key)
            levels_out[l] = {"mask": mask, "child": child}
        levels_out[0] = {"mask": np.array([root_key[0]], np.uint8), "child": resolve(0, root_key).reshape(1, 8)}
        self.levels = levels_out
        self.stats["nNodesDAG"] = 1 + sum(len(x["mask"]) for x in levels_out[1:])
        return dict(self.stats)

    def build(self, L, step, lo=None, hi=None):
        self.shard_build(L, step, None if lo is None else (lo, hi))
        self.shard_finish()
        return self.levels


# --------------------------------------------------------------------------- SDAG (orbit-min)
_PERM = np.array([[i ^ s for i in range(8)] for s in range(8)])          # slot i takes slot i^s
_PERMBITS = np.array([[sum((((m >> (i ^ s)) & 1) << i) for i in range(8)) for m in range(256)] for s in range(8)], dtype=np.uint8)
# priority order of the reference's lookups: id, X, Y, Z, XY, XZ, YZ, XYZ as (x,y,z) -> s = 4x+2y+z
PRIORITY = [0, 4, 2, 1, 6, 5, 3, 7]


def _variant(mask, child, mir, s, child_inv):
    """mirror by axes s (X=4,Y=2,Z=1), toggle child mirror bits, clear them where the child is invariant."""
    m2 = _PERMBITS[s][mask]
    c2 = child[:, _PERM[s]]
    mir2 = np.stack([_PERMBITS[s][mir[:, a]] for a in range(3)], axis=1)
    if s:
        has = np.zeros(len(mask), dtype=np.uint8)
        for i in range(8):
            has |= ((c2[:, i] != NULL).astype(np.uint8) << i)
        for a, bit in ((0, 4), (1, 2), (2, 1)):
            if s & bit:
                mir2[:, a] ^= has
                if child_inv is not None:
                    clear = np.zeros(len(mask), dtype=np.uint8)
                    for i in range(8):
                        ok = c2[:, i] != NULL
                        idx = np.where(ok, c2[:, i], 0)
                        inv_bit = {4: 1, 2: 2, 1: 4}[bit]
                        clear |= ((ok & ((child_inv[idx] & inv_bit) != 0)).astype(np.uint8) << i)
                    mir2[:, a] &= ~clear
    return m2, c2, mir2


def _pack(mask, child, mir):
    return np.concatenate([mask.reshape(-1, 1).astype(np.uint64), child.astype(np.uint64), mir.astype(np.uint64)], axis=1)


def to_sdag(levels):
    """levels: list of dict(mask, child) in DAG state -> SDAG levels with mirror/inv (orbit-min formulation)."""
    L = len(levels)
    out = [dict(mask=x["mask"].copy(), child=x["child"].copy(), mirror=np.zeros((len(x["mask"]), 3), np.uint8),
                inv=np.zeros(len(x["mask"]), np.uint8)) for x in levels]
    for lev in range(L - 1, 0, -1):
        cur = out[lev]
        n = len(cur["mask"])
        child_inv = out[lev + 1]["inv"] if lev < L - 1 else None
        kid = _pack(cur["mask"], cur["child"], cur["mirror"])
        inv = np.zeros(n, np.uint8)
        for bit, s in ((1, 4), (2, 2), (4, 1)):
            m2, c2, mir2 = _variant(cur["mask"], cur["child"], cur["mirror"], s, None)   # raw mirror, no invertInvs
            inv |= (np.all(_pack(m2, c2, mir2) == kid, axis=1).astype(np.uint8) * bit)
        variants = [kid] + [None] * 7
        for s in range(1, 8):
            variants[s] = _pack(*_variant(cur["mask"], cur["child"], cur["mirror"], s, child_inv))
        # class key = lexicographic min over the 8 variants
        allv = np.stack(variants, axis=1)                       # (n, 8, 12)
        cls = np.empty((n, 12), dtype=np.uint64)
        for i in range(n):
            cls[i] = min(map(tuple, allv[i]))
        # group by class key; representative = min original index
        _, inverse = np.unique(cls, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        rep = np.full(inverse.max() + 1 if n else 0, n, dtype=np.int64)
        np.minimum.at(rep, inverse, np.arange(n))
        rep_of = rep[inverse]
        is_rep = rep_of == np.arange(n)
        new_id = np.cumsum(is_rep) - 1
        corr_id = new_id[rep_of]
        flags = np.zeros(n, dtype=np.int64)
        for i in range(n):
            if is_rep[i]:
                continue
            target = tuple(kid[rep_of[i]])
            for s in PRIORITY:
                if tuple(allv[i, s]) == target:
                    flags[i] = s
                    break
            else:
                raise AssertionError("orbit-min formulation broke: no variant equals the representative")
        out[lev] = dict(mask=cur["mask"][is_rep], child=cur["child"][is_rep], mirror=cur["mirror"][is_rep], inv=inv[is_rep])
        par = out[lev - 1]
        for j in range(8):
            ok = par["child"][:, j] != NULL
            old = np.where(ok, par["child"][:, j], 0)
            par["child"][:, j] = np.where(ok, corr_id[old].astype(np.uint32), NULL)
            f = flags[old]
            par["mirror"][:, 0] |= ((ok & ((f & 4) != 0)).astype(np.uint8) << j)
            par["mirror"][:, 1] |= ((ok & ((f & 2) != 0)).astype(np.uint8) << j)
            par["mirror"][:, 2] |= ((ok & ((f & 1) != 0)).astype(np.uint8) << j)
    return out
