"""File images of the reference's output formats, produced by the product's C++ host encoders
(csrc/host/encoders.cpp) through the C ABI entry point svb_encode()."""
from __future__ import annotations

import ctypes as C

import numpy as np

KINDS = {"svdag": 0, "ussvdag": 1, "ssvdag": 2, "esvdag": 2}


def encode_view(octree, kind: str) -> np.ndarray:
    """The file image as a uint8 view of the context's pinned buffer (svb_encode_view: written on the GPU, one D2H copy of
    the finished image, no further host copy).  Valid until the octree changes or another kind is encoded."""
    L = octree._L
    L.svb_encode_view.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    ptr, size = C.c_void_p(), C.c_uint64()
    octree._check(L.svb_encode_view(octree._h, KINDS[kind], C.byref(ptr), C.byref(size)))
    if size.value == 0:
        return np.zeros(0, np.uint8)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(int(size.value),))


def encode(octree, kind: str) -> bytes:
    """octree: capi.GeomOctree in the state the format needs (DAG for svdag/esvdag, SDAG for ussvdag/ssvdag)."""
    return encode_view(octree, kind).tobytes()


def encode_copy(octree, kind: str) -> bytes:
    """The two-call form of svb_encode (size query, then copy into a caller-owned buffer)."""
    L = octree._L
    L.svb_encode.restype = C.c_int64
    L.svb_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    k = KINDS[kind]
    n = L.svb_encode(octree._h, k, None, 0)
    if n < 0:
        octree._check(int(n))
    buf = np.zeros(n, np.uint8)
    n2 = L.svb_encode(octree._h, k, buf.ctypes.data, n)
    if n2 != n:
        octree._check(int(n2) if n2 < 0 else -1)
    return buf.tobytes()


def ssvdag_order_from_refs(refs_per_level):
    """svb_ssvdag_order_from_refs (host only): list of uint32 arrays (one per level, level 0 first) -> list of orders."""
    from .capi import lib
    L = lib()
    L.svb_ssvdag_order_from_refs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    start = np.zeros(len(refs_per_level) + 1, np.uint32)
    start[1:] = np.cumsum([len(r) for r in refs_per_level])
    refs = np.ascontiguousarray(np.concatenate(refs_per_level), dtype=np.uint32)
    order = np.zeros(len(refs), np.uint32)
    rc = L.svb_ssvdag_order_from_refs(refs.ctypes.data, start.ctypes.data, len(refs_per_level), order.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"svb_ssvdag_order_from_refs failed: {rc}")
    return [order[start[l]:start[l + 1]].copy() for l in range(len(refs_per_level))]


def encode_levels(levels, bboxF, root_side: float, n_nodes: int, state: int, kind: str) -> bytes:
    """Host-only path (no GPU): levels = list of dicts with mask (n,), child (n,8), optional mirror (n,3),
    childLevel (n,8).  Goes through the same C++ encoders via svb_encode_levels()."""
    from .capi import lib
    L = lib()
    L.svb_encode_levels.restype = C.c_int64
    L.svb_encode_levels.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_double, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    counts = np.array([len(l["mask"]) for l in levels], dtype=np.uint64)
    mask = np.ascontiguousarray(np.concatenate([l["mask"] for l in levels]), dtype=np.uint8)
    child = np.ascontiguousarray(np.concatenate([np.asarray(l["child"]).reshape(-1, 8) for l in levels]), dtype=np.uint32)
    mir = np.ascontiguousarray(np.concatenate([np.asarray(l.get("mirror", np.zeros((len(l["mask"]), 3)))).reshape(-1, 3) for l in levels]), dtype=np.uint8)
    has_cl = all("childLevel" in l for l in levels)
    cl = np.ascontiguousarray(np.concatenate([np.asarray(l["childLevel"]).reshape(-1, 8) for l in levels]), dtype=np.uint32) if has_cl else None
    bb = np.ascontiguousarray(bboxF, dtype=np.float32)
    args = [len(levels), counts.ctypes.data, mask.ctypes.data, child.ctypes.data, mir.ctypes.data,
            cl.ctypes.data if cl is not None else None, bb.ctypes.data, float(root_side), int(n_nodes), int(state), KINDS[kind]]
    n = L.svb_encode_levels(*args, None, 0)
    if n < 0:
        raise RuntimeError(f"svb_encode_levels({kind}) failed: {n}")
    buf = np.zeros(n, np.uint8)
    L.svb_encode_levels(*args, buf.ctypes.data, n)
    return buf.tobytes()


def decode_svdag(file_bytes: bytes):
    """EncodedSVDAG::load + decode through the product's host code (svb_decode_svdag; no GPU): returns
    (levels, bboxF, root_side, n_nodes) with levels = list of dicts {mask (n,), child (n,8)}."""
    from .capi import lib
    L = lib()
    L.svb_decode_svdag.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    nlev = C.c_uint32()
    counts = np.zeros(32, np.uint64)
    rc = L.svb_decode_svdag(buf.ctypes.data, len(buf), C.byref(nlev), counts.ctypes.data, None, None, None, None, None)
    if rc != 0:
        raise RuntimeError(f"svb_decode_svdag failed: {rc}")
    n = int(counts[:nlev.value].sum())
    mask = np.zeros(n, np.uint8)
    child = np.zeros((n, 8), np.uint32)
    bb = np.zeros(6, np.float32)
    rs = C.c_double()
    nn = C.c_uint64()
    rc = L.svb_decode_svdag(buf.ctypes.data, len(buf), C.byref(nlev), counts.ctypes.data, mask.ctypes.data, child.ctypes.data,
                            bb.ctypes.data, C.byref(rs), C.byref(nn))
    if rc != 0:
        raise RuntimeError(f"svb_decode_svdag failed: {rc}")
    levels, off = [], 0
    for l in range(nlev.value):
        c = int(counts[l])
        levels.append({"mask": mask[off:off + c].copy(), "child": child[off:off + c].copy()})
        off += c
    return levels, bb, float(rs.value), int(nn.value)
