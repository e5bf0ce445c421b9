"""Multi-GPU build: one process per GPU, torch.distributed for the plumbing.

`build_sharded()` drives the protocol documented in include/svb.h ("multi-GPU"): every rank
voxelizes and reduces its share of the sub-octrees, then, bottom-up, the ranks all-gather their
unique nodes of a level (key with global child ids + min order key) and rebuild the identical
global table; finally every rank holds the same DAG a single GPU would have built
(GeomOctree::buildDAG's "join + last DAG pass", geom_octree.cpp:397-425, across devices).

Host round trips: ONE.  The record counts of every level and the per-rank counters are known as
soon as the local phase ends, so a single small all-gather (read back once) sizes every later
buffer and yields the summed counters; from there on export -> all-gather -> import of all levels
and of the sub-octree roots is enqueued on the octree's own CUDA stream (svb_stream(), wrapped as
a torch ExternalStream so NCCL orders itself against it) without waiting for the device.

The exchange is backend-agnostic: `octree` only has to provide the shard_* methods
(capi.GeomOctree does, over the C ABI; the CPU test-suite plugs a numpy model in and runs the
same code over gloo)."""
from __future__ import annotations

import contextlib

import numpy as np


def octree_stream(octree, device):
    """torch stream context in which the octree's shard_* calls and the collectives are mutually ordered."""
    import torch
    if device.type != "cuda" or not hasattr(octree, "stream_ptr"):
        return contextlib.nullcontext()
    ts = getattr(octree, "_torch_stream", None)      # the context was created on a torch stream (capi.GeomOctree(stream=...))
    if ts is not None:
        return torch.cuda.stream(ts)
    return torch.cuda.stream(torch.cuda.ExternalStream(octree.stream_ptr(), device=device))


def build_sharded(octree, levels: int, step: int, bbox, group=None, device=None):
    """Collective: call on every rank with the same arguments.  Returns the stats dict of shard_finish.

    A 64-bit tag collision in the level merge (error -5 from shard_finish) is seen by every rank alike -- all of them
    import identical records under the same seed -- so all of them repeat the build under the next merge seed."""
    for seed in range(4):
        try:
            return _build_sharded_once(octree, levels, step, bbox, group, device)
        except Exception as e:  # capi.SvbError; the numpy model of the CPU tests never raises it
            if getattr(e, "code", None) != -5 or seed == 3 or not hasattr(octree, "set_merge_seed"):
                raise
            octree.set_merge_seed(seed + 1)


def _build_sharded_once(octree, levels: int, step: int, bbox, group=None, device=None):
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    import time
    t0 = time.perf_counter()
    octree.shard_build(levels, step, bbox, rank, world)          # returns with this rank's local phase complete
    t1 = time.perf_counter()
    first, last, ntiles, counters = octree.shard_info()
    order = list(range(last, first - 1, -1))
    local = [octree.shard_level_count(g) for g in order]          # (records, bytes per record) per level, final after the local phase
    with octree_stream(octree, device):
        # the one host round trip: every rank's record counts for all levels + its counters
        mine = torch.tensor([n for n, _ in local] + [int(x) for x in counters], dtype=torch.int64, device=device)
        allv = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allv, mine, group=group)
        table = allv.cpu().numpy().reshape(world, mine.numel())
        exchanged = 0
        keep = []                                                 # gathered buffers stay alive until the device is done with them
        for k, g in enumerate(order):
            counts = np.ascontiguousarray(table[:, k], dtype=np.uint64)
            rec = local[k][1]
            stride = int(max(16, (int(counts.max()) * rec + 15) // 16 * 16))
            buf = torch.empty(stride, dtype=torch.uint8, device=device)      # padding beyond the records is never read
            octree.shard_export_level(g, buf.data_ptr())
            allb = torch.empty(world * stride, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allb, buf, group=group)
            octree.shard_import_level(g, allb.data_ptr(), counts, stride)
            keep.append((buf, allb))
            exchanged += world * stride
        nt = max(int(ntiles), 1)
        roots = torch.empty(nt, dtype=torch.int32, device=device)
        octree.shard_export_roots(roots.data_ptr())
        allr = torch.empty(world * nt, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(allr, roots, group=group)
        octree.shard_import_roots(allr.data_ptr())
        totals = [int(x) for x in table[:, len(order):].sum(axis=0)]
        st = octree.shard_finish(totals)                          # synchronises: the buffers above may go
    del keep
    st["bytesExchanged"] = exchanged + world * nt * 4
    st["wallLocalMs"] = (t1 - t0) * 1e3                           # host wall clock: local phase / exchange + finish
    st["wallMergeMs"] = (time.perf_counter() - t1) * 1e3
    return st


def set_triangles_sharded(octree, host_tris, ntris: int, group=None, device=None):
    """Collective upload of the triangle soup: every rank copies only its 1/world slice of the (pinned) host array over
    its own PCIe link, an all-gather over NVLink completes the soup in every rank's HBM, and the octree borrows the
    device buffer (svb_set_triangles_device).  host_tris: 1-D float32 torch tensor of ntris*9 values, identical on all ranks."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    nfloat = int(ntris) * 9
    chunk = (-(-nfloat // world) + 3) // 4 * 4                  # floats per rank, 16-byte granular
    with octree_stream(octree, device):
        buf = torch.empty(world * chunk + 16, dtype=torch.float32, device=device)   # a few floats of slack, as svb_set_triangles keeps
        full = buf[:world * chunk]
        lo, hi = rank * chunk, min((rank + 1) * chunk, nfloat)
        mine = full[rank * chunk:(rank + 1) * chunk]
        if hi > lo:
            mine[:hi - lo].copy_(host_tris[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full, mine, group=group)     # in place: rank r's slice is already where it belongs
        torch.cuda.current_stream(device).synchronize()
    octree._dev_tris = buf                                       # keep the buffer alive while the octree borrows it
    octree.set_triangles_device(full.data_ptr(), int(ntris))
    return int((hi - lo) * 4 if hi > lo else 0)
