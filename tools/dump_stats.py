"""Every timing field of the stats block of one 16K^3 city build (after two warm-up builds)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import __graft_entry__ as g
pkg = g._pkg()
tris = pkg.meshgen.city(256)
v = tris.reshape(-1, 3)
bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
t = pkg.GeomOctree(tris)
for _ in range(3):
    st = t.build(14, 4, bbox=bbox)
print({k: (round(x, 2) if isinstance(x, float) else x) for k, x in st.items() if k.startswith("ms") or k.startswith("wall") or k in ("nBatches", "nKernelLaunches")})
