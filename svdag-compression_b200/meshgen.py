"""Procedural triangle meshes for the svbuilder hot path.

All generators return a float32 array of shape (T, 3, 3) -- a triangle soup in file
order, which is exactly what the reference hands to the voxelizer
(`Scene::_triangles`, /root/reference/src/symvox/scene.cpp:394-416).  Every mesh pins
its bounding box to exactly [0,1]^3 with two tiny corner triangles so that
`_rootSide == 1` and every voxel centre / half side is an exact dyadic double
(SL's `center()`/`half_side_lengths()` formulas are not in the tree).

`write_obj` / `write_bincache` emit the two on-disk inputs the reference's
`Scene::loadObj` accepts (scene.cpp:51-71 prefers `<obj>.bincache`; layout of the
cache: scene.hpp:112-119 header, scene.cpp:324-374 body).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

__all__ = [
    "sphere", "menger_sponge", "sphere_menger", "terrain", "city", "composite", "composite_crop", "soup",
    "corner_pins", "write_obj", "write_bincache", "make_mesh",
]


# --------------------------------------------------------------------------- helpers
def corner_pins(eps: float = 2.0 ** -7) -> np.ndarray:
    """Two small triangles touching (0,0,0) and (1,1,1): pin the bbox to [0,1]^3."""
    a = np.array([[[0, 0, 0], [eps, 0, 0], [0, eps, 0]],
                  [[1, 1, 1], [1 - eps, 1, 1], [1, 1 - eps, 1]]], dtype=np.float32)
    return a


def _quads_to_tris(p00, p10, p11, p01) -> np.ndarray:
    """Each quad (p00,p10,p11,p01) -> two triangles. Inputs (..., 3) -> (2N, 3, 3)."""
    t1 = np.stack([p00, p10, p11], axis=-2)
    t2 = np.stack([p00, p11, p01], axis=-2)
    return np.concatenate([t1.reshape(-1, 3, 3), t2.reshape(-1, 3, 3)], axis=0)


def _box_tris(lo, hi) -> np.ndarray:
    """12 triangles of an axis-aligned box."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    c = np.array([[lo[0] if not (i & 4) else hi[0],
                   lo[1] if not (i & 2) else hi[1],
                   lo[2] if not (i & 1) else hi[2]] for i in range(8)])
    faces = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for a, b, cc, d in faces:
        tris.append([c[a], c[b], c[cc]])
        tris.append([c[a], c[cc], c[d]])
    return np.asarray(tris)


# --------------------------------------------------------------------------- meshes
def sphere(n_lat: int = 64, n_lon: int = 128, center=(0.5, 0.5, 0.5), radius: float = 0.4,
           pins: bool = True) -> np.ndarray:
    """UV sphere, n_lat x n_lon quads (degenerate cap triangles dropped)."""
    th = np.linspace(0.0, np.pi, n_lat + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    pts = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=-1)
    pts = np.asarray(center) + radius * pts
    tris = _quads_to_tris(pts[:-1, :-1], pts[1:, :-1], pts[1:, 1:], pts[:-1, 1:])
    tris = tris.astype(np.float32)
    # drop zero-area cap triangles (two coincident vertices)
    keep = ~((tris[:, 0] == tris[:, 1]).all(-1) | (tris[:, 1] == tris[:, 2]).all(-1)
             | (tris[:, 0] == tris[:, 2]).all(-1))
    tris = tris[keep]
    if pins:
        tris = np.concatenate([tris, corner_pins()], axis=0)
    return np.ascontiguousarray(tris, dtype=np.float32)


def menger_sponge(level: int = 3, lo=(0.5, 0.5, 0.5), size: float = 0.5) -> np.ndarray:
    """Level-`level` Menger sponge: 20^level cubes, 12 triangles each (inner faces kept)."""
    cells = np.zeros((1, 3), dtype=np.int64)
    for _ in range(level):
        offs = np.array([(i, j, k) for i in range(3) for j in range(3) for k in range(3)
                         if (i == 1) + (j == 1) + (k == 1) < 2], dtype=np.int64)
        cells = (cells[:, None, :] * 3 + offs[None, :, :]).reshape(-1, 3)
    h = size / (3 ** level)
    unit = _box_tris((0, 0, 0), (1, 1, 1))  # (12,3,3)
    tris = (cells[:, None, None, :] + unit[None]) * h + np.asarray(lo)
    return tris.reshape(-1, 3, 3).astype(np.float32)


def sphere_menger(n_lat: int = 256, n_lon: int = 512, sponge_level: int = 3) -> np.ndarray:
    """BASELINE.json configs[0]: tessellated sphere + Menger sponge (off the voxel lattice)."""
    s = sphere(n_lat, n_lon, center=(0.27, 0.27, 0.27), radius=0.2, pins=False)
    nudge = -(1.0 / 3.0) * 2.0 ** -10
    m = menger_sponge(sponge_level, lo=(0.5 + nudge,) * 3, size=0.5)
    return np.ascontiguousarray(np.concatenate([s, m, corner_pins()], axis=0), dtype=np.float32)


def _value_noise(n: int, octaves: int, seed: int) -> np.ndarray:
    """(n+1)x(n+1) multi-octave value noise in [0,1] (bilinear lattice interpolation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.linspace(0.0, 1.0, n + 1)
    h = np.zeros((n + 1, n + 1))
    amp, freq = 1.0, 4
    for _ in range(octaves):
        g = rng.random((freq + 2, freq + 2))
        u = x * freq
        i = np.minimum(u.astype(np.int64), freq - 1)
        f = u - i
        f = f * f * (3.0 - 2.0 * f)
        a = g[i][:, i]
        b = g[i + 1][:, i]
        c = g[i][:, i + 1]
        d = g[i + 1][:, i + 1]
        fx = f[:, None]
        fy = f[None, :]
        h += amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy)
        amp *= 0.5
        freq *= 2
    h -= h.min()
    h /= h.max()
    return h


def terrain(n: int = 1024, octaves: int = 5, seed: int = 1234, z_scale: float = 1.0) -> np.ndarray:
    """BASELINE.json configs[1]: heightfield on an (n x n)-vertex grid -> 2(n-1)^2 triangles."""
    m = n - 1
    h = _value_noise(m, octaves, seed) * z_scale
    x = np.linspace(0.0, 1.0, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    pts = np.stack([X, Y, h], axis=-1)
    tris = _quads_to_tris(pts[:-1, :-1], pts[1:, :-1], pts[1:, 1:], pts[:-1, 1:])
    tris = np.concatenate([tris.astype(np.float32), corner_pins()], axis=0)
    return np.ascontiguousarray(tris, dtype=np.float32)


def city(lots: int = 256, seed: int = 2024, window_rows: int = 3) -> np.ndarray:
    """BASELINE.json configs[2]: lots x lots city blocks on a z=0 ground plane.

    Each lot carries a tower of 1-3 set-back boxes plus `window_rows` rings of inset
    window boxes on the lowest tier (~150 triangles per lot at the defaults)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cell = 1.0 / lots
    out = [np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]], dtype=np.float64)]
    unit = _box_tris((0, 0, 0), (1, 1, 1))
    heights = rng.uniform(0.02, 0.9, size=(lots, lots))
    heights[lots // 2, lots // 2] = 1.0
    tiers = rng.integers(1, 4, size=(lots, lots))
    margin = rng.uniform(0.08, 0.2, size=(lots, lots))
    los, his = [], []
    for i in range(lots):
        for j in range(lots):
            x0, y0 = i * cell, j * cell
            m = margin[i, j] * cell
            z0 = 0.0
            nt = int(tiers[i, j])
            for t in range(nt):
                z1 = heights[i, j] * (t + 1) / nt
                inset = m + t * 0.12 * cell
                los.append((x0 + inset, y0 + inset, z0))
                his.append((x0 + cell - inset, y0 + cell - inset, z1))
                if t == 0:
                    # window boxes protruding from the four walls of the lowest tier
                    for r in range(window_rows):
                        zc = z0 + (z1 - z0) * (r + 0.5) / window_rows
                        dz = (z1 - z0) * 0.2 / window_rows
                        w = 0.18 * cell
                        d = 0.03 * cell
                        xm = x0 + 0.5 * cell
                        ym = y0 + 0.5 * cell
                        los.append((xm - w, y0 + inset - d, zc - dz)); his.append((xm + w, y0 + inset, zc + dz))
                        los.append((xm - w, y0 + cell - inset, zc - dz)); his.append((xm + w, y0 + cell - inset + d, zc + dz))
                        los.append((x0 + inset - d, ym - w, zc - dz)); his.append((x0 + inset, ym + w, zc + dz))
                        los.append((x0 + cell - inset, ym - w, zc - dz)); his.append((x0 + cell - inset + d, ym + w, zc + dz))
                z0 = z1
    los = np.asarray(los)
    his = np.asarray(his)
    boxes = los[:, None, None, :] + unit[None] * (his - los)[:, None, None, :]
    out.append(boxes.reshape(-1, 3, 3))
    out.append(corner_pins().astype(np.float64))
    tris = np.concatenate(out, axis=0)
    return np.ascontiguousarray(np.clip(tris, 0.0, 1.0), dtype=np.float32)


def composite(n_terrain: int = 1024, lots: int = 256) -> np.ndarray:
    """BASELINE.json configs[4]: low terrain (z in [0,0.25]) with the city above it."""
    t = terrain(n_terrain, z_scale=0.25)[:-2]
    c = city(lots)
    return np.ascontiguousarray(np.concatenate([t, c], axis=0), dtype=np.float32)


def composite_crop(n_terrain: int = 513, lots: int = 128, octant: int = 0) -> np.ndarray:
    """BASELINE.md row 5 ("CPU parity on a cropped octant"): the triangles of composite(n_terrain, lots) that lie entirely
    inside top-level octant `octant` (bit 2 = x, bit 1 = y, bit 0 = z high half), mapped onto [0,1]^3 by v -> 2v - o -- exact
    in float32 for every vertex (a doubling and, for the high half, the subtraction of 1 from a value in [0.5, 1] x 2) -- plus
    the corner pins.  The sub-scene has the mix of the full composite (general terrain triangles under axis-aligned city
    boxes) at a size the reference's CPU build finishes in minutes."""
    t = composite(n_terrain, lots)[:-2].astype(np.float32)
    o = np.array([(octant >> 2) & 1, (octant >> 1) & 1, octant & 1], dtype=np.float32)
    lo, hi = 0.5 * o, 0.5 * o + 0.5
    inside = np.all((t >= lo) & (t <= hi), axis=(1, 2))
    u = t[inside] * np.float32(2.0) - o
    return np.ascontiguousarray(np.concatenate([u, corner_pins()], axis=0), dtype=np.float32)


def soup(n: int = 400, seed: int = 7) -> np.ndarray:
    """Adversarial triangle soup for parity tests: random triangles of all sizes mixed with degenerate ones
    (points, segments, axis-aligned segments, zero-area slivers), triangles lying exactly in voxel faces, and
    scene-spanning triangles.  Vertices on a 1/64 lattice make exact ties with voxel boundaries common."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for i in range(n):
        kind = i % 8
        c = rng.integers(4, 60, size=3) / 64.0
        if kind == 0:      # point
            t = np.stack([c, c, c])
        elif kind == 1:    # segment along an axis
            d = np.zeros(3); d[rng.integers(0, 3)] = rng.integers(1, 6) / 64.0
            t = np.stack([c, c + d, c])
        elif kind == 2:    # oblique segment (two equal vertices)
            t = np.stack([c, c + rng.integers(-5, 6, size=3) / 64.0, c])
        elif kind == 3:    # collinear sliver
            d = rng.integers(-4, 5, size=3) / 64.0
            t = np.stack([c, c + d, c + 2 * d])
        elif kind == 4:    # axis-aligned face triangle on the lattice
            a = rng.integers(0, 3); u, w = [k for k in range(3) if k != a]
            p1, p2 = c.copy(), c.copy()
            p1[u] += rng.integers(1, 8) / 64.0; p2[w] += rng.integers(1, 8) / 64.0
            t = np.stack([c, p1, p2])
        elif kind == 5:    # large random triangle
            t = rng.random((3, 3))
        elif kind == 6:    # tiny random triangle (sub-voxel at 256^3)
            t = c + rng.random((3, 3)) / 512.0
        else:              # medium random triangle off the lattice
            t = c + (rng.random((3, 3)) - 0.5) / 8.0
        out.append(np.clip(t, 0.0, 1.0))
    tris = np.concatenate([np.asarray(out), corner_pins().astype(np.float64)], axis=0)
    return np.ascontiguousarray(tris, dtype=np.float32)


def make_mesh(name: str, **kw) -> np.ndarray:
    """Look a generator up by name ('sphere', 'sphere_menger', 'terrain', 'city', 'composite')."""
    return {"sphere": sphere, "sphere_menger": sphere_menger, "terrain": terrain,
            "city": city, "composite": composite, "composite_crop": composite_crop, "menger": menger_sponge, "soup": soup}[name](**kw)


# --------------------------------------------------------------------------- writers
def write_obj(path, tris: np.ndarray) -> None:
    """ASCII OBJ: 3 `v` lines + one `f a b c` line per triangle (no sharing, triangles only;
    scene.cpp:183 mis-parses quads).  %.9g round-trips float32 exactly through `%f`-sscanf."""
    tris = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    with open(path, "w") as f:
        for k, t in enumerate(tris):
            for v in t:
                f.write("v %.9g %.9g %.9g\n" % (float(v[0]), float(v[1]), float(v[2])))
            f.write("f %d %d %d\n" % (3 * k + 1, 3 * k + 2, 3 * k + 3))


def write_bincache(path, tris: np.ndarray, materials=None) -> None:
    """`<obj>.bincache` as `Scene::saveBinObj` writes it (scene.cpp:324-348): 56-byte
    header, float32 vertices, 300-byte materials (one default material unless `materials` -- a per-triangle array of
    material indices, scene.hpp:41-45 `TIndexedTri::material` -- asks for more), 56-byte indexed triangles."""
    tris = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
    nt = tris.shape[0]
    verts = tris.reshape(-1, 3)
    lo = verts.min(axis=0)
    hi = verts.max(axis=0)
    nmat = 1 if materials is None else int(np.max(materials)) + 1
    with open(path, "wb") as f:
        f.write(struct.pack("<4Q6f", nt * 3, 0, nmat, nt, *[float(x) for x in lo], *[float(x) for x in hi]))
        f.write(verts.astype("<f4").tobytes())
        for k in range(nmat):
            name = b"Voxelator Default Mat" if nmat == 1 else b"mat%d" % k
            mat = name + b"\0" * (256 - len(name))
            mat += struct.pack("<9f", 0.9, 0.9, 0.9, 0.2, 0.2, 0.2, 0.1, 0.1, 0.1)
            mat += struct.pack("<2f", 0.0, 0.0)
            assert len(mat) == 300
            f.write(mat)
        idx = np.zeros((nt, 7), dtype="<u8")
        idx[:, 0:3] = np.arange(nt * 3, dtype=np.uint64).reshape(nt, 3)
        if materials is not None:
            idx[:, 6] = np.asarray(materials, dtype=np.uint64)
        f.write(idx.tobytes())


def materials_for(tris: np.ndarray, n_materials: int = 13, seed: int = 99) -> np.ndarray:
    """Synthetic per-triangle material ids: runs of consecutive triangles share a material (as the faces of one object do
    in an OBJ with `usemtl` groups), ids in [0, n_materials)."""
    nt = int(np.asarray(tris).reshape(-1, 9).shape[0])
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros(nt, np.uint32)
    i = 0
    while i < nt:
        run = int(rng.integers(1, 40))
        out[i:i + run] = rng.integers(0, n_materials)
        i += run
    return out


def write_scene(obj_path, tris: np.ndarray, ascii_obj: bool = False, materials=None) -> Path:
    """Write `<obj_path>` (a stub unless ascii_obj) and `<obj_path>.bincache`."""
    obj_path = Path(obj_path)
    if ascii_obj:
        write_obj(obj_path, tris)
    else:
        obj_path.write_text("# geometry lives in the .bincache next to this file\n")
    write_bincache(str(obj_path) + ".bincache", tris, materials)
    return obj_path
