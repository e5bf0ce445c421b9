import os, sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
def affine(tris, scale, offset):
    t = tris.reshape(-1, 3).astype(np.float64) * np.asarray(scale) + np.asarray(offset)
    return np.ascontiguousarray(t.astype(np.float32).reshape(-1, 9))
tris = affine(pkg.meshgen.make_mesh("city", lots=8), (3.7, 2.9, 5.3), (11.3, -5.1, 2.9))
os.environ["SVB_CENTRE"] = "chain"
t = pkg.GeomOctree(tris)
st = t.build(9, 2)
print(tuple(st[k] for k in ("nTotalVoxels", "nNodesSVO", "nNodesDAG")), "want (1829834, 565167, 9612)")
