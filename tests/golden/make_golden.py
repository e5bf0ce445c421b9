"""Mint golden vectors from the UNMODIFIED reference binary (oracle/_ref/svbuilder_ref).

The reference ships no tests or fixtures, so its own compiled svbuilder is the
source of truth.  Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz holding the input triangle soup, the (levels, step, cross)
arguments, every file svbuilder wrote, and the per-level "Reduced level" counts parsed from
its log (geom_octree.cpp:509).  Cases are small so the fixtures stay a few hundred KB.
"""
import importlib.util
import re
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

spec = importlib.util.spec_from_file_location("_meshgen", ROOT / "svdag-compression_b200" / "meshgen.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

CASES = [
    # name, mesh, mesh kwargs, levels, step, cross
    ("sphere_L6_s0", "sphere", dict(n_lat=16, n_lon=32), 6, 0, False),
    ("sphere_L7_s1", "sphere", dict(n_lat=16, n_lon=32), 7, 1, False),
    ("sphere_L7_s2", "sphere", dict(n_lat=16, n_lon=32), 7, 2, False),
    ("sphere_L6_s0_c", "sphere", dict(n_lat=16, n_lon=32), 6, 0, True),
    ("city_L7_s0", "city", dict(lots=4), 7, 0, False),
    ("city_L7_s2", "city", dict(lots=4), 7, 2, False),
    ("city_L7_s1_c", "city", dict(lots=4), 7, 1, True),
    ("terrain_L6_s0", "terrain", dict(n=24), 6, 0, False),
    ("terrain_L7_s3", "terrain", dict(n=24), 7, 3, False),
    ("spongeball_L7_s0", "sphere_menger", dict(n_lat=12, n_lon=24, sponge_level=1), 7, 0, False),
    ("spongeball_L7_s2_c", "sphere_menger", dict(n_lat=12, n_lon=24, sponge_level=1), 7, 2, True),
    # adversarial soup: degenerate triangles (points, segments, slivers), lattice ties, scene-spanning triangles
    ("soup_L7_s1", "soup", dict(n=400, seed=7), 7, 1, False),
    ("soup_L7_s0_c", "soup", dict(n=400, seed=7), 7, 0, True),
    # scaled + shifted scenes (bbox is not the unit cube): float narrowing of sub-octree boxes, empty sub-octree roots
    ("city_affine_L8_s2", "city", dict(lots=4, affine=((3.7, 2.9, 5.3), (11.3, -5.1, 2.9))), 8, 2, False),
    ("terrain_affine_L7_s1", "terrain", dict(n=24, affine=((3.7, 2.9, 5.3), (11.3, -5.1, 2.9))), 7, 1, False),
    ("city_affine_L7_s1_c", "city", dict(lots=4, affine=((3.7, 2.9, 5.3), (11.3, -5.1, 2.9))), 7, 1, True),
]


def main():
    out_dir = Path(__file__).resolve().parent
    only = set(sys.argv[1:])
    for name, mesh, kw, levels, step, cross in CASES:
        if only and name not in only:
            continue
        kw = dict(kw)
        affine = kw.pop("affine", None)
        tris = mg.make_mesh(mesh, **kw)
        if affine is not None:
            t = tris.reshape(-1, 3).astype(np.float64) * np.asarray(affine[0]) + np.asarray(affine[1])
            tris = np.ascontiguousarray(t.astype(np.float32).reshape(-1, 9))
        with tempfile.TemporaryDirectory() as td:
            r = orc.run_reference(td, tris, levels, step, cross=cross)
        reduced = np.zeros((levels, 2), dtype=np.int64)
        for m in re.finditer(r"Reduced level (\d+) from (\d+) to (\d+) nodes", r["log"]):
            reduced[int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
        stats = {k: int(re.search(rf"{k}:\s+.*\((\d+)\)", r["log"]).group(1))
                 for k in ("Voxels", "SVO Nodes", "DAG Nodes", "SDAG Nodes")}
        arrs = {"tris": tris, "levels": levels, "step": step, "cross": cross, "reduced": reduced,
                "stats": np.array([stats["Voxels"], stats["SVO Nodes"], stats["DAG Nodes"], stats["SDAG Nodes"]], dtype=np.int64)}
        for ext, data in r["files"].items():
            arrs["file_" + ext.replace(".", "_")] = np.frombuffer(data, dtype=np.uint8)
        np.savez_compressed(out_dir / f"{name}.npz", **arrs)
        print(name, tris.shape[0], "tris", {k: len(v) for k, v in r["files"].items()}, stats)


if __name__ == "__main__":
    main()
