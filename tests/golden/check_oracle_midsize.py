"""One-off pin of the CPU restatement (oracle/svdag_oracle.cpp) at a size far above the golden scenes: builds the city of
tests/golden/midsize_city4k.json (64x64 lots, 4096^3, levels 12 step 3; 1.42 G voxels) with the sequential oracle --
about 12 minutes on one core -- and compares counts and the SHA-256 of all four encoded files with what the UNMODIFIED
reference svbuilder wrote (make_fullsize.py midsize).  Last runs (round 1, session 3): midsize_city4k.json all equal, 711 s; size_terrain4k.json (BASELINE configs[1] at full
size) all equal, 526 s; size_spongeball1k.json (configs[0]) all equal, 6.6 s.

    python tests/golden/check_oracle_midsize.py [midsize_city4k.json | size_terrain4k.json | size_spongeball1k.json]
"""
import hashlib
import importlib.util
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

spec = importlib.util.spec_from_file_location("_meshgen", ROOT / "svdag-compression_b200" / "meshgen.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def main():
    gold = json.loads((Path(__file__).resolve().parent / (sys.argv[1] if len(sys.argv) > 1 else "midsize_city4k.json")).read_text())
    tris = mg.make_mesh(gold.get("mesh", "city"), **gold.get("kw", {"lots": gold.get("lots", 64)}))
    t0 = time.time()
    o = orc.OracleOctree(tris)
    o.build(gold["levels"], gold["step"])
    ok = (o.stat("nTotalVoxels"), o.stat("nNodesSVO"), o.stat("nNodesDAG")) == (gold["Voxels"], gold["SVO Nodes"], gold["DAG Nodes"])
    print(f"oracle build {time.time() - t0:.1f} s; counts equal: {ok}", flush=True)
    res = {k: hashlib.sha256(o.encode(k)).hexdigest() for k in ("svdag", "esvdag")}
    o.to_sdag()
    ok &= o.stat("nNodesSDAG") == gold["SDAG Nodes"]
    res.update({k: hashlib.sha256(o.encode(k)).hexdigest() for k in ("ussvdag", "ssvdag")})
    for k, v in res.items():
        same = v == gold["files"][k]["sha256"]
        ok &= same
        print(f"{k}: {'equal' if same else 'DIFFERENT'}")
    print("PINNED" if ok else "MISMATCH", f"({time.time() - t0:.1f} s)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
