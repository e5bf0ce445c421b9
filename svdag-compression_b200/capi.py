"""ctypes binding of include/svb.h and a Python mirror of the reference's GeomOctree call
surface (src/symvox/geom_octree.hpp:107-157): buildSVO+toDAG / buildDAG -> build(),
toSDAG -> to_sdag(), mergeAcrossAllLevels -> cross_merge(), getNodeData -> level()."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
STATE_NAMES = {0: "EMPTY", 1: "SVO", 2: "DAG", 3: "SDAG"}
NULL_NODE = 0xFFFFFFFE


class SvbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svb error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nTotalVoxels", "nNodesSVO", "nNodesDAG", "nNodesSDAG", "nNodesLastLevSVO",
                                           "nNodesLastLevDAG", "nCrossLevelMerged", "nNodes", "nTiles", "nBatches", "nPairsTotal")]
    _fields_ += [("rootSide", C.c_double), ("bboxF", C.c_float * 6)]
    _fields_ += [(n, C.c_double) for n in ("msVoxelize", "msDedup", "msFinalize", "msSdag", "msCrossMerge", "msTotal")]
    _fields_ += [("nKernelLaunches", C.c_uint64), ("nExactTests", C.c_uint64), ("nHashRetries", C.c_uint64)]

    def as_dict(self):
        d = {}
        for n, _t in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if n == "bboxF" else v
        return d


class ProfRec(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("level", C.c_uint32), ("n_in", C.c_uint64), ("n_out", C.c_uint64),
                ("ms", C.c_double), ("bytes", C.c_double)]


def lib_path() -> Path:
    return HERE / "libsvb.so"


_lib = None


def lib() -> C.CDLL:
    """Load libsvb.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise SvbError(-2, f"{p} missing: build the CUDA extension first (python svdag-compression_b200/build.py)")
        L = C.CDLL(str(p))
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.svb_create.restype = vp
        L.svb_create.argtypes = [C.c_int]
        L.svb_create_on_stream.restype = vp
        L.svb_create_on_stream.argtypes = [C.c_int, vp]
        L.svb_destroy.argtypes = [vp]
        L.svb_last_error.restype = C.c_char_p
        L.svb_last_error.argtypes = [vp]
        L.svb_version.restype = C.c_char_p
        L.svb_stream.restype = vp
        L.svb_stream.argtypes = [vp]
        L.svb_synchronize.argtypes = [vp]
        L.svb_set_merge_seed.argtypes = [vp, u64]
        L.svb_set_triangles.argtypes = [vp, vp, u64]
        L.svb_set_triangles_device.argtypes = [vp, vp, u64]
        L.svb_build.argtypes = [vp, u32, u32, vp, vp, C.POINTER(Stats)]
        L.svb_to_sdag.argtypes = [vp, C.POINTER(Stats)]
        L.svb_cross_merge.argtypes = [vp, C.POINTER(Stats)]
        L.svb_state.argtypes = [vp]
        L.svb_levels.argtypes = [vp]
        L.svb_levels.restype = u32
        L.svb_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.svb_level_count.argtypes = [vp, u32, C.POINTER(u64)]
        L.svb_level_count_svo.argtypes = [vp, u32, C.POINTER(u64)]
        L.svb_download_level.argtypes = [vp, u32, vp, vp, vp, vp, vp]
        L.svb_upload_levels.argtypes = [vp, u32, vp, vp, vp, vp, C.c_double, u64]
        L.svb_shard_build.argtypes = [vp, u32, u32, vp, vp, u32, u32]
        L.svb_shard_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64), vp]
        L.svb_shard_level_count.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u32)]
        L.svb_shard_export_level.argtypes = [vp, u32, vp]
        L.svb_shard_import_level.argtypes = [vp, u32, vp, vp, u64]
        L.svb_shard_export_roots.argtypes = [vp, vp]
        L.svb_shard_import_roots.argtypes = [vp, vp]
        L.svb_shard_finish.argtypes = [vp, vp, C.POINTER(Stats)]
        L.svb_build_svo_materials.argtypes = [vp, u32, vp, vp, vp, C.POINTER(u64)]
        L.svb_download_leaf_materials.argtypes = [vp, vp, vp]
        L.svb_attribute_bit_trees.argtypes = [vp, u32, C.c_int, vp, vp]
        L.svb_gray_code.argtypes = [u32]
        L.svb_gray_code.restype = u32
        L.svb_set_profiling.argtypes = [vp, C.c_int]
        L.svb_profile_count.argtypes = [vp]
        L.svb_profile_get.argtypes = [vp, C.c_int, C.POINTER(ProfRec)]
        L.svb_set_batch_budget.argtypes = [vp, u64]
        L.svb_raycast_depth.argtypes = [C.c_int, vp, u64, C.c_int, vp, vp, u32, u32, u32, u32, C.c_float, vp]
        _lib = L
    return _lib


class GeomOctree:
    """One GPU-resident octree.  Method names follow the reference class; data stays in HBM until
    `level()` / `levels_host()` copies it out (what the encoders read through getNodeData())."""

    def __init__(self, tris=None, device: int = 0, stream=None):
        """stream: a torch.cuda.Stream the context should work on (kept alive by this object); default: its own stream."""
        self._L = lib()
        self._torch_stream = stream
        self._h = self._L.svb_create_on_stream(device, stream.cuda_stream) if stream is not None else self._L.svb_create(device)
        if not self._h:
            raise SvbError(-2, "svb_create failed: no usable CUDA device (there is no CPU fallback)")
        self._tris = None
        if tris is not None:
            self.set_triangles(tris)

    def close(self):
        if getattr(self, "_h", None):
            self._L.svb_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise SvbError(rc, self._L.svb_last_error(self._h).decode())

    # ---- Scene
    def set_triangles(self, tris):
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        self._tris = t
        self._check(self._L.svb_set_triangles(self._h, t.ctypes.data, t.shape[0]))

    def set_triangles_ptr(self, host_ptr: int, ntris: int):
        """Raw host pointer (e.g. pinned memory) to ntris*9 float32; copied H2D during the call."""
        self._tris = None
        self._check(self._L.svb_set_triangles(self._h, host_ptr, ntris))

    def set_triangles_device(self, dev_ptr: int, ntris: int):
        self._check(self._L.svb_set_triangles_device(self._h, dev_ptr, ntris))

    def scene_bbox(self):
        v = self._tris.reshape(-1, 3)
        return v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64)

    # ---- GeomOctree
    def build(self, levels: int, step: int = 0, bbox=None, group=None, sharded: bool = False) -> dict:
        """step == 0: buildSVO + toDAG; step > 0: buildDAG.  With sharded=True (inside an initialised
        torch.distributed job) the sub-octrees are spread over the ranks and merged (sharded.py)."""
        if sharded:
            from .sharded import build_sharded
            return build_sharded(self, levels, step, bbox if bbox is not None else self.scene_bbox(), group=group)
        lo, hi = bbox if bbox is not None else self.scene_bbox()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        st = Stats()
        self._check(self._L.svb_build(self._h, levels, step, lo.ctypes.data, hi.ctypes.data, C.byref(st)))
        return st.as_dict()

    def stream_ptr(self) -> int:
        """cudaStream_t of this context (the shard_export_* / shard_import_* calls are stream-ordered on it)."""
        return int(self._L.svb_stream(self._h) or 0)

    def set_merge_seed(self, seed: int):
        self._check(self._L.svb_set_merge_seed(self._h, int(seed)))

    def synchronize(self):
        self._check(self._L.svb_synchronize(self._h))

    # ---- multi-GPU protocol (include/svb.h); driven by sharded.build_sharded()
    def shard_build(self, levels, step, bbox, rank, world):
        lo, hi = bbox if bbox is not None else self.scene_bbox()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        self._check(self._L.svb_shard_build(self._h, levels, step, lo.ctypes.data, hi.ctypes.data, rank, world))

    def shard_info(self):
        a, b, n = C.c_uint32(), C.c_uint32(), C.c_uint64()
        cnt = np.zeros(5, np.uint64)
        self._check(self._L.svb_shard_info(self._h, C.byref(a), C.byref(b), C.byref(n), cnt.ctypes.data))
        return int(a.value), int(b.value), int(n.value), [int(x) for x in cnt]

    def shard_level_count(self, g):
        n, r = C.c_uint64(), C.c_uint32()
        self._check(self._L.svb_shard_level_count(self._h, g, C.byref(n), C.byref(r)))
        return int(n.value), int(r.value)

    def shard_export_level(self, g, dev_ptr):
        self._check(self._L.svb_shard_export_level(self._h, g, dev_ptr))

    def shard_import_level(self, g, dev_ptr, counts, stride):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self._check(self._L.svb_shard_import_level(self._h, g, dev_ptr, counts.ctypes.data, int(stride)))

    def shard_export_roots(self, dev_ptr):
        self._check(self._L.svb_shard_export_roots(self._h, dev_ptr))

    def shard_import_roots(self, dev_ptr):
        self._check(self._L.svb_shard_import_roots(self._h, dev_ptr))

    def shard_finish(self, totals) -> dict:
        t = np.ascontiguousarray(totals, dtype=np.uint64)
        st = Stats()
        self._check(self._L.svb_shard_finish(self._h, t.ctypes.data, C.byref(st)))
        return st.as_dict()

    def to_sdag(self) -> dict:
        st = Stats()
        self._check(self._L.svb_to_sdag(self._h, C.byref(st)))
        return st.as_dict()

    def cross_merge(self) -> dict:
        st = Stats()
        self._check(self._L.svb_cross_merge(self._h, C.byref(st)))
        return st.as_dict()

    @property
    def state(self) -> int:
        return int(self._L.svb_state(self._h))

    @property
    def levels(self) -> int:
        return int(self._L.svb_levels(self._h))

    def stats(self) -> dict:
        st = Stats()
        self._check(self._L.svb_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def level_sizes(self):
        out = []
        for l in range(self.levels):
            n = C.c_uint64()
            self._check(self._L.svb_level_count(self._h, l, C.byref(n)))
            out.append(int(n.value))
        return out

    def level_sizes_svo(self):
        out = []
        for l in range(self.levels):
            n = C.c_uint64()
            self._check(self._L.svb_level_count_svo(self._h, l, C.byref(n)))
            out.append(int(n.value))
        return out

    def level(self, lev: int) -> dict:
        n = C.c_uint64()
        self._check(self._L.svb_level_count(self._h, lev, C.byref(n)))
        n = int(n.value)
        mask = np.zeros(n, np.uint8)
        child = np.zeros((n, 8), np.uint32)
        mir = np.zeros((n, 3), np.uint8)
        inv = np.zeros(n, np.uint8)
        chl = np.zeros((n, 8), np.uint32)
        self._check(self._L.svb_download_level(self._h, lev, mask.ctypes.data, child.ctypes.data, mir.ctypes.data,
                                               inv.ctypes.data, chl.ctypes.data))
        return {"mask": mask, "child": child, "mirror": mir, "inv": inv, "childLevel": chl}

    def levels_host(self):
        return [self.level(l) for l in range(self.levels)]

    def upload_levels(self, levels, bboxF, root_side: float, n_voxels: int):
        counts = np.array([len(l["mask"]) for l in levels], dtype=np.uint64)
        mask = np.ascontiguousarray(np.concatenate([l["mask"] for l in levels]), dtype=np.uint8)
        child = np.ascontiguousarray(np.concatenate([l["child"].reshape(-1, 8) for l in levels]), dtype=np.uint32)
        bb = np.ascontiguousarray(bboxF, dtype=np.float32)
        self._check(self._L.svb_upload_levels(self._h, len(levels), counts.ctypes.data, mask.ctypes.data, child.ctypes.data,
                                              bb.ctypes.data, float(root_side), int(n_voxels)))

    # ---- material-id leaves + attribute bit-trees (include/svb.h)
    def build_svo_materials(self, levels: int, materials, bbox=None):
        """buildSVO(levels, bbox, false, NULL, putMaterialIdInLeaves=true): (mask (n,), material (n, 8)) of the leaf level in
        the reference's node order; unset voxels hold 0xFFFFFFFE."""
        lo, hi = bbox if bbox is not None else self.scene_bbox()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        m = np.ascontiguousarray(materials, dtype=np.uint32)
        n = C.c_uint64()
        self._check(self._L.svb_build_svo_materials(self._h, levels, lo.ctypes.data, hi.ctypes.data, m.ctypes.data, C.byref(n)))
        mask = np.zeros(int(n.value), np.uint8)
        mat = np.zeros((int(n.value), 8), np.uint32)
        self._check(self._L.svb_download_leaf_materials(self._h, mask.ctypes.data, mat.ctypes.data))
        return mask, mat

    def attribute_bit_trees(self, nbits: int, gray: bool):
        """(nodes[nbits], voxels[nbits]) of the bit-trees of the last build_svo_materials (self-specified)."""
        nodes = np.zeros(nbits, np.uint64)
        vox = np.zeros(nbits, np.uint64)
        self._check(self._L.svb_attribute_bit_trees(self._h, nbits, 1 if gray else 0, nodes.ctypes.data, vox.ctypes.data))
        return nodes, vox

    # ---- instrumentation
    def set_profiling(self, on=True):
        """True / 1: records of the last build; 2: records accumulate over builds (read them once with profile()); 3: as 2, but only
        the launches of the "emit" family are bracketed by events; False: off."""
        self._L.svb_set_profiling(self._h, int(on))

    def profile(self):
        out = []
        for i in range(self._L.svb_profile_count(self._h)):
            r = ProfRec()
            self._L.svb_profile_get(self._h, i, C.byref(r))
            out.append({"name": r.name.decode(), "level": int(r.level), "n_in": int(r.n_in), "n_out": int(r.n_out),
                        "ms": float(r.ms), "bytes": float(r.bytes)})
        return out

    def set_batch_budget(self, nbytes: int):
        self._L.svb_set_batch_budget(self._h, int(nbytes))


FILE_KIND = {"svdag": 0, "multi.svdag": 0, "ussvdag": 1, "ssvdag": 2, "esvdag": 2}


def raycast_depth(file_bytes: bytes, kind: str, view_inv, proj_inv, width: int, height: int, max_iters: int = 512,
                  draw_level: int = 0, projection_factor: float = 1e30, device: int = 0) -> np.ndarray:
    """Depth image (h, w, 3) = (t, level, iterations) of an encoded file: svb_raycast_depth (CUDA)."""
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    vi = np.ascontiguousarray(view_inv, dtype=np.float32).reshape(16)
    pi = np.ascontiguousarray(proj_inv, dtype=np.float32).reshape(16)
    out = np.zeros((height, width, 3), np.float32)
    rc = lib().svb_raycast_depth(device, buf.ctypes.data, len(buf), FILE_KIND[kind], vi.ctypes.data, pi.ctypes.data, width, height,
                                 max_iters, draw_level, float(projection_factor), out.ctypes.data)
    if rc != 0:
        raise SvbError(rc, "svb_raycast_depth failed" + (" (no CUDA device: there is no CPU fallback)" if rc == -6 else ""))
    return out
