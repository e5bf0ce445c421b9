#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over the round-2 code paths on small multi-batch builds:
# warm-up batches, bin-once root pairs, warp-granular emit, fused scans, 4^3 level without first touches (+ direct query),
# list-based inner passes, toSDAG, cross-level merge, the GPU encoders of all five files, the shard export / import kernels
# (two simulated ranks on one device) and the hash-retry path.  Logs -> gpurun_out/san_r2_*.log (copied to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import os, sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
for mesh, kw, L, s, budget in [("city", dict(lots=12), 10, 2, 6 << 20), ("terrain", dict(n=48), 8, 1, 4 << 20)]:
    tris = pkg.meshgen.make_mesh(mesh, **kw)
    t = pkg.GeomOctree(tris)
    t.set_batch_budget(budget)
    st = t.build(L, s)
    sizes = [len(pkg.encoders.encode(t, k)) for k in ("svdag", "esvdag")]
    t.to_sdag()
    sizes += [len(pkg.encoders.encode(t, k)) for k in ("ussvdag", "ssvdag")]
    u = pkg.GeomOctree(tris)
    u.build(L, s)
    cm = u.cross_merge()
    sizes.append(len(pkg.encoders.encode(u, "svdag")))
    print(mesh, st["nTotalVoxels"], st["nNodesDAG"], st["nBatches"], cm["nCrossLevelMerged"], sizes)
if os.environ.get("SAN_SHARDED", "1") == "1":
    import torch
    from test_gpu_parity import _simulate_ranks
    tris = pkg.meshgen.make_mesh("city", lots=8)
    octs, stats = _simulate_ranks(pkg, tris, 9, 2, 2)
    print("sharded", stats[0]["nNodesDAG"], stats[1]["nNodesDAG"])
os.environ["SVB_TEST_WEAK_HASH"] = "6"
tris = pkg.meshgen.make_mesh("terrain", n=32)
t = pkg.GeomOctree(tris)
st = t.build(8, 2)
print("retries", st["nHashRetries"])
PY
run() {  # tool extra-env
  env $2 SVB_VX_STATS=1 timeout 170 compute-sanitizer --tool $1 --error-exitcode 3 --print-limit 10 python /tmp/san2.py > gpurun_out/san_r2_$1.log 2>&1
  echo "$1 exit $?" | tee -a gpurun_out/san_r2_$1.log
  grep -v "vx-stats\] tiles" gpurun_out/san_r2_$1.log | tail -7
}
run memcheck
run racecheck
run initcheck SAN_SHARDED=0
run synccheck
