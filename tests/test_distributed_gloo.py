"""world_size-2 (and 3) gloo runs of the multi-GPU build protocol on CPU.

The exchange code under test is the product's svdag-compression_b200/sharded.py (the same function
bench.py and the GPU tests call over NCCL); the per-rank "device" is oracle/parallel_model.Builder,
the numpy model of the GPU formulation, which speaks the shard_* protocol on CPU memory.  The merged
result of every rank must equal the single-process oracle node for node."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mesh, kw, levels, step, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from conftest import load_pkg
    from oracle import parallel_model as pm
    pkg = load_pkg()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris = pkg.meshgen.make_mesh(mesh, **kw)
    b = pm.Builder(tris)
    st = pkg.sharded.build_sharded(b, levels, step, None, device=torch.device("cpu"))
    np.savez(Path(out_dir) / f"rank{rank}.npz", nvox=st["nTotalVoxels"], nsvo=st["nNodesSVO"], ndag=st["nNodesDAG"],
             exchanged=st["bytesExchanged"],
             **{f"mask{l}": lv["mask"] for l, lv in enumerate(b.levels)}, **{f"child{l}": lv["child"] for l, lv in enumerate(b.levels)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mesh,kw,levels,step", [
    (2, "sphere", dict(n_lat=16, n_lon=32), 6, 1),
    (2, "city", dict(lots=4), 6, 2),
    (3, "terrain", dict(n=24), 6, 2),
], ids=["sphere-w2", "city-w2", "terrain-w3"])
def test_sharded_build_matches_single_process_oracle(tmp_path, orc, meshgen, world, mesh, kw, levels, step):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mesh, kw, levels, step, str(tmp_path)), nprocs=world, join=True)
    tris = meshgen.make_mesh(mesh, **kw)
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert int(z["nvox"]) == o.stat("nTotalVoxels")
        assert int(z["nsvo"]) == o.stat("nNodesSVO")
        assert int(z["ndag"]) == o.stat("nNodesDAG")
        assert int(z["exchanged"]) > 0
        for l in range(levels):
            want = o.level(l)
            assert np.array_equal(z[f"mask{l}"], want["mask"]), f"rank {r} level {l} masks differ"
            assert np.array_equal(z[f"child{l}"], want["child"]), f"rank {r} level {l} children differ"
