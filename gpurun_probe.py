# scratch: feasibility + stage timings of the big workloads
import sys, time, json
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g._pkg()
import numpy as np
which = sys.argv[1:] or ["terrain", "city128", "city256"]
for w in which:
    t0 = time.time()
    if w == "terrain": tris, L, s = pkg.meshgen.terrain(1024), 12, 3
    elif w == "city64": tris, L, s = pkg.meshgen.city(64), 12, 3
    elif w == "city128": tris, L, s = pkg.meshgen.city(128), 13, 4
    elif w == "city256": tris, L, s = pkg.meshgen.city(256), 14, 4
    print(w, "mesh", tris.shape[0], "tris gen %.1fs" % (time.time() - t0), flush=True)
    t = pkg.GeomOctree(tris)
    t.set_profiling(True)
    for it in range(3):
        t0 = time.time()
        try:
            st = t.build(L, s)
        except Exception as e:
            print("  FAILED", e, flush=True); break
        t1 = time.time()
        sd = t.to_sdag()
        t2 = time.time()
        print("  it%d build %.3fs sdag %.3fs | vox %.3e svo %.3e dag %d sdag %d tiles %d batches %d pairs %.3e exact %.3e launches %d | ms total %.1f vox %.1f dedup %.1f fin %.1f sdag %.1f" % (
            it, t1 - t0, t2 - t1, st["nTotalVoxels"], st["nNodesSVO"], st["nNodesDAG"], sd["nNodesSDAG"], st["nTiles"], st["nBatches"], st["nPairsTotal"], st["nExactTests"], st["nKernelLaunches"],
            st["msTotal"], st["msVoxelize"], st["msDedup"], st["msFinalize"], sd["msSdag"]), flush=True)
    prof = t.profile()
    agg = {}
    for r in prof:
        k = (r["name"], r["level"])
        a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += r["n_in"]; a[1] += r["ms"]; a[2] += r["bytes"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print("   ", k, "n_in %.3e ms %.3f GB/s(alg) %.1f" % (a[0], a[1], a[2] / a[1] / 1e6 if a[1] else 0), flush=True)
    t0 = time.time(); b = pkg.encoders.encode(t, "ssvdag"); print("  encode ssvdag %.3fs %d bytes" % (time.time() - t0, len(b)), flush=True)
    t.close()
