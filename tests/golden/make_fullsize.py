"""Mint tests/golden/fullsize_city16k.json: run the UNMODIFIED reference svbuilder (oracle/_ref/svbuilder_ref) on
BASELINE.json's headline configuration -- procedural city (lots=256, 11.0 M triangles), 14 levels, step 4 -- and
record the sizes and SHA-256 of the files it writes plus its result block.  About one hour on 8 cores, ~6 GB RAM.

    python tests/golden/make_fullsize.py [fullsize|midsize|bigsize|terrain|spongeball|composite_crop] [workdir]

`midsize` (city lots=64 at 4096^3, levels 12 step 3) is the same generator at 1/16 of the ground area and finishes in
minutes: tests/golden/midsize_city4k.json.
"""
import hashlib
import importlib.util
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

spec = importlib.util.spec_from_file_location("_meshgen", ROOT / "svdag-compression_b200" / "meshgen.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


CONFIGS = {
    # name: (mesh generator, kwargs, levels, step, output file)
    "fullsize": ("city", dict(lots=256), 14, 4, "fullsize_city16k.json"),    # BASELINE.json's headline configuration (hours of CPU)
    "midsize": ("city", dict(lots=64), 12, 3, "midsize_city4k.json"),        # same generator, 1/16 of the ground area at the same voxels per lot (minutes)
    "bigsize": ("city", dict(lots=128), 13, 3, "bigsize_city8k.json"),       # 1/4 of the ground area: 2.75 M triangles at 8192^3 (tens of minutes)
    "terrain": ("terrain", dict(n=1024), 12, 3, "size_terrain4k.json"),      # BASELINE.json configs[1]: 2.09 M general triangles at 4096^3
    "spongeball": ("sphere_menger", dict(), 10, 1, "size_spongeball1k.json"),  # BASELINE.json configs[0]: sphere + Menger sponge at 1024^3
    # one top-level octant of the terrain + city composite (configs[4]'s geometry mix), cropped and blown up to the unit cube
    "composite_crop": ("composite_crop", dict(n_terrain=513, lots=128, octant=0), 12, 3, "size_composite_crop4k.json"),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in CONFIGS else "fullsize"
    rest = [a for a in sys.argv[1:] if a not in CONFIGS]
    mesh, kw, levels, step, out_name = CONFIGS[which]
    work = rest[0] if rest else f"/tmp/fullref/work_{which}"
    tris = mg.make_mesh(mesh, **kw)
    r = orc.run_reference(work, tris, levels, step)
    out = {"workload": f"meshgen.make_mesh({mesh!r}, **{kw}), levels {levels}, step {step}", "mesh": mesh, "kw": kw, "levels": levels, "step": step,
           "triangles": int(len(tris)), "reference_seconds": r["seconds"],
           "files": {k: {"sha256": hashlib.sha256(v).hexdigest(), "bytes": len(v)} for k, v in r["files"].items()}}
    if mesh == "city":
        out["lots"] = kw["lots"]
    for k in ("Voxels", "SVO Nodes", "DAG Nodes", "SDAG Nodes"):
        out[k] = int(re.search(rf"{k}:\s+.*\((\d+)\)", r["log"]).group(1))
    (Path(__file__).resolve().parent / out_name).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
