"""Mint tests/golden/attr_*.npz: the leaf level of the UNMODIFIED reference's
GeomOctree::buildSVO(levels, bbox, false, NULL, putMaterialIdInLeaves = true) -- per leaf node its child mask and its
8 child slots (material id of the last triangle touching the voxel, 0xFFFFFFFE for unset voxels) in the reference's node
order -- obtained through oracle/_ref/ref_attr_driver (our 40-line driver around the reference sources,
oracle/ref_attr_driver.cpp; built by `make -C oracle ref_attr`).  Meshes and materials are regenerated from their
generator arguments, which are stored with the arrays.

    python tests/golden/make_golden_attr.py
"""
import importlib.util
import json
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
    ("attr_sphere_L7", "sphere", dict(n_lat=32, n_lon=64), 7, dict(n_materials=13, seed=99)),
    ("attr_city_L8", "city", dict(lots=6), 8, dict(n_materials=40, seed=5)),
    ("attr_soup_L7", "soup", dict(n=400, seed=7), 7, dict(n_materials=200, seed=11)),
    ("attr_terrain_L8", "terrain", dict(n=48), 8, dict(n_materials=7, seed=3)),
]


def main():
    out_dir = Path(__file__).resolve().parent
    for name, mesh, kw, levels, mkw in CASES:
        tris = mg.make_mesh(mesh, **kw)
        mats = mg.materials_for(tris, **mkw)
        with tempfile.TemporaryDirectory() as td:
            mask, mat = orc.run_reference_materials(td, tris, mats, levels)
        np.savez_compressed(out_dir / f"{name}.npz", mask=mask, material=mat, levels=levels,
                            meta=json.dumps(dict(mesh=mesh, kw=kw, materials=mkw, triangles=int(len(tris)))))
        print(name, len(mask), "leaf nodes,", int(np.unique(mat[mat != 0xFFFFFFFE]).size), "materials in use")


if __name__ == "__main__":
    main()
