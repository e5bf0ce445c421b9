"""File images of the reference's output formats, produced by the product's C++ host encoders
(csrc/host/encoders.cpp) through the C ABI entry point svb_encode()."""
from __future__ import annotations

import ctypes as C

import numpy as np

KINDS = {"svdag": 0, "ussvdag": 1, "ssvdag": 2, "esvdag": 2}


def encode(octree, kind: str) -> bytes:
    """octree: capi.GeomOctree in the state the format needs (DAG for svdag/esvdag, SDAG for ussvdag/ssvdag)."""
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
