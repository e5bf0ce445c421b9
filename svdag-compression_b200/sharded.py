"""Multi-GPU build: one process per GPU, torch.distributed for the plumbing.

`build_sharded()` drives the protocol documented in include/svb.h ("multi-GPU"): every rank
voxelizes and reduces its share of the sub-octrees, then, bottom-up, the ranks all-gather their
unique nodes of a level (key with global child ids + min order key) and rebuild the identical
global table; finally every rank holds the same DAG a single GPU would have built
(GeomOctree::buildDAG's "join + last DAG pass", geom_octree.cpp:397-425, across devices).

The exchange is backend-agnostic: `octree` only has to provide the shard_* methods
(capi.GeomOctree does, over the C ABI; the CPU test-suite plugs a numpy model in and runs the
same code over gloo)."""
from __future__ import annotations

import numpy as np


def _all_gather_ints(dist, vals, device, group):
    import torch
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    out = torch.empty(dist.get_world_size(group) * len(vals), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(dist.get_world_size(group), len(vals))


def _sync(device):
    import torch
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()


def build_sharded(octree, levels: int, step: int, bbox, group=None, device=None):
    """Collective: call on every rank with the same arguments.  Returns the stats dict of shard_finish."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    octree.shard_build(levels, step, bbox, rank, world)
    first, last, ntiles, counters = octree.shard_info()
    exchanged = 0
    for g in range(last, first - 1, -1):
        n, rec = octree.shard_level_count(g)
        counts = _all_gather_ints(dist, [n], device, group)[:, 0]
        stride = int(max(16, (int(counts.max()) * rec + 15) // 16 * 16))
        mine = torch.zeros(stride, dtype=torch.uint8, device=device)
        octree.shard_export_level(g, mine.data_ptr())
        allb = torch.empty(world * stride, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allb, mine, group=group)
        _sync(device)
        octree.shard_import_level(g, allb.data_ptr(), np.ascontiguousarray(counts, dtype=np.uint64), stride)
        exchanged += world * stride
    nt = max(int(ntiles), 1)
    roots = torch.zeros(nt, dtype=torch.int32, device=device)
    octree.shard_export_roots(roots.data_ptr())
    allr = torch.empty(world * nt, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allr, roots, group=group)
    _sync(device)
    octree.shard_import_roots(allr.data_ptr())
    tot = torch.tensor([int(x) for x in counters], dtype=torch.int64, device=device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    _sync(device)
    st = octree.shard_finish([int(x) for x in tot.cpu().numpy()])
    st["bytesExchanged"] = exchanged + world * nt * 4
    return st
