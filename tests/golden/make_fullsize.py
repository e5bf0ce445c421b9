"""Mint tests/golden/fullsize_city16k.json: run the UNMODIFIED reference svbuilder (oracle/_ref/svbuilder_ref) on
BASELINE.json's headline configuration -- procedural city (lots=256, 11.0 M triangles), 14 levels, step 4 -- and
record the sizes and SHA-256 of the files it writes plus its result block.  About one hour on 8 cores, ~6 GB RAM.

    python tests/golden/make_fullsize.py [workdir]
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


def main():
    work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/fullref/work"
    tris = mg.city(256)
    r = orc.run_reference(work, tris, 14, 4)
    out = {"workload": "meshgen.city(lots=256), levels 14, step 4", "reference_seconds": r["seconds"],
           "files": {k: {"sha256": hashlib.sha256(v).hexdigest(), "bytes": len(v)} for k, v in r["files"].items()}}
    for k in ("Voxels", "SVO Nodes", "DAG Nodes", "SDAG Nodes"):
        out[k] = int(re.search(rf"{k}:\s+.*\((\d+)\)", r["log"]).group(1))
    (Path(__file__).resolve().parent / "fullsize_city16k.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
