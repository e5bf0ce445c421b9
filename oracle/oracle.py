"""ctypes front-end of oracle/_ref/libsvdag_oracle.so (the CPU restatement) and a runner
for oracle/_ref/svbuilder_ref (the unmodified reference binary).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/svdag_oracle.cpp.  Importable only
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_BIN = HERE / "_ref" / "svbuilder_ref"
LIB = HERE / "_ref" / "libsvdag_oracle.so"
DDA_LIB = HERE / "_ref" / "libdda_oracle.so"

STAT = {"nTotalVoxels": 0, "nNodesSVO": 1, "nNodesDAG": 2, "nNodesSDAG": 3,
        "nNodesLastLevSVO": 4, "nNodesLastLevDAG": 5, "nCrossLevelMerged": 6, "nNodes": 7}
NULLNODE = 0xFFFFFFFE


def build(ref: bool = True) -> None:
    """(Re)build the oracle library (and the reference binary when /root/reference exists)."""
    targets = ["oracle"] + (["ref", "ref_attr"] if ref else [])
    subprocess.run(["make", "-C", str(HERE)] + targets, check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build(ref=False)
        L = C.CDLL(str(LIB))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_build.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
        L.orc_build_svo_only.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        L.orc_build_svo_materials.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_to_dag.argtypes = [C.c_void_p]
        L.orc_to_sdag.argtypes = [C.c_void_p]
        L.orc_cross_merge.argtypes = [C.c_void_p]
        L.orc_cross_merge.restype = C.c_uint
        L.orc_state.argtypes = [C.c_void_p]
        L.orc_levels.argtypes = [C.c_void_p]
        L.orc_levels.restype = C.c_uint
        L.orc_level_size.argtypes = [C.c_void_p, C.c_uint]
        L.orc_level_size.restype = C.c_uint64
        L.orc_level_size_before_dag.argtypes = [C.c_void_p, C.c_uint]
        L.orc_level_size_before_dag.restype = C.c_uint64
        L.orc_stat.argtypes = [C.c_void_p, C.c_int]
        L.orc_stat.restype = C.c_uint64
        L.orc_root_side.argtypes = [C.c_void_p]
        L.orc_root_side.restype = C.c_double
        L.orc_get_level.argtypes = [C.c_void_p, C.c_uint] + [C.c_void_p] * 5
        L.orc_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.orc_encode.restype = C.c_int64
        L.orc_test_tri_box.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        _lib = L
    return _lib


class OracleOctree:
    """Mirror of the reference's GeomOctree call sequence (svbuilder/main.cpp:147-271)."""

    def __init__(self, tris: np.ndarray):
        self.tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        self.h = lib().orc_create(self.tris.ctypes.data, self.tris.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def scene_bbox(self):
        v = self.tris.reshape(-1, 3)
        return v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64)

    def build(self, levels: int, step: int, bbox=None) -> None:
        lo, hi = bbox if bbox is not None else self.scene_bbox()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        rc = lib().orc_build(self.h, levels, step, lo.ctypes.data, hi.ctypes.data)
        if rc != 0:
            raise RuntimeError("orc_build failed")

    def build_svo_materials(self, levels: int, materials, bbox=None):
        """buildSVO(levels, bbox, false, NULL, putMaterialIdInLeaves=true): returns (mask (n,), material (n, 8)) of the leaf
        level in SVO node order; material slots of unset voxels hold 0xFFFFFFFE (nullNode)."""
        lo, hi = bbox if bbox is not None else self.scene_bbox()
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        m = np.ascontiguousarray(materials, dtype=np.uint32)
        lib().orc_build_svo_materials(self.h, levels, lo.ctypes.data, hi.ctypes.data, m.ctypes.data)
        lv = self.level(levels - 1)
        return lv["mask"], lv["child"]

    def to_sdag(self) -> None:
        lib().orc_to_sdag(self.h)

    def cross_merge(self) -> int:
        return int(lib().orc_cross_merge(self.h))

    @property
    def levels(self) -> int:
        return int(lib().orc_levels(self.h))

    def stat(self, name: str) -> int:
        return int(lib().orc_stat(self.h, STAT[name]))

    def level_sizes(self):
        return [int(lib().orc_level_size(self.h, l)) for l in range(self.levels)]

    def level_sizes_before_dag(self):
        return [int(lib().orc_level_size_before_dag(self.h, l)) for l in range(self.levels)]

    def level(self, lev: int):
        n = int(lib().orc_level_size(self.h, lev))
        mask = np.zeros(n, np.uint8)
        child = np.zeros((n, 8), np.uint32)
        mir = np.zeros((n, 3), np.uint8)
        inv = np.zeros(n, np.uint8)
        chlev = np.zeros((n, 8), np.uint32)
        lib().orc_get_level(self.h, lev, mask.ctypes.data, child.ctypes.data, mir.ctypes.data,
                            inv.ctypes.data, chlev.ctypes.data)
        return {"mask": mask, "child": child, "mirror": mir, "inv": inv, "childLevel": chlev}

    def encode(self, kind: str) -> bytes:
        k = {"svdag": 0, "ussvdag": 1, "ssvdag": 2, "esvdag": 2}[kind]
        n = lib().orc_encode(self.h, k, None, 0)
        if n < 0:
            raise RuntimeError(f"oracle encode({kind}) refused (wrong state)")
        buf = np.zeros(n, np.uint8)
        lib().orc_encode(self.h, k, buf.ctypes.data, n)
        return buf.tobytes()


def test_tri_box(center, half: float, tri9) -> bool:
    c = np.ascontiguousarray(center, dtype=np.float64)
    t = np.ascontiguousarray(tri9, dtype=np.float32).reshape(9)
    return bool(lib().orc_test_tri_box(c.ctypes.data, float(half), t.ctypes.data))


def svbuilder_files(tris: np.ndarray, levels: int, step: int, cross: bool = False) -> dict:
    """Everything `svbuilder m.obj L s [-c]` writes (main.cpp:220-271), from the restatement."""
    o = OracleOctree(tris)
    o.build(levels, step)
    out = {"svdag": o.encode("svdag")}
    if cross:
        o.cross_merge()
        out["multi.svdag"] = o.encode("svdag")
        return out
    out["esvdag"] = o.encode("esvdag")
    o.to_sdag()
    out["ussvdag"] = o.encode("ussvdag")
    out["ssvdag"] = o.encode("ssvdag")
    return out


def run_reference(workdir, tris: np.ndarray, levels: int, step: int, cross: bool = False,
                  threads: int | None = None, name: str = "m", timeout: float | None = None) -> dict:
    """Run the unmodified reference binary on `tris`; returns {'files': {...}, 'log': str, 'seconds': wall}."""
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location("_svb_meshgen", HERE.parent / "svdag-compression_b200" / "meshgen.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    if not REF_BIN.exists():
        raise FileNotFoundError(f"{REF_BIN} missing: run `make -C oracle ref` where /root/reference exists")
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    obj = wd / f"{name}.obj"
    mg.write_scene(obj, tris)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    cmd = [str(REF_BIN), str(obj), str(levels), str(step)] + (["-c"] if cross else [])
    t0 = time.time()
    p = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=timeout)
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(f"reference svbuilder failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    base = wd / f"{name}_{levels}"
    files = {}
    for ext in (["svdag", "multi.svdag"] if cross else ["svdag", "esvdag", "ussvdag", "ssvdag"]):
        f = Path(str(base) + ("-multi.svdag" if ext == "multi.svdag" else "." + ext))
        if f.exists():
            files[ext] = f.read_bytes()
    return {"files": files, "log": p.stdout, "seconds": dt}


ATTR_BIN = HERE / "_ref" / "ref_attr_driver"


def run_reference_materials(workdir, tris: np.ndarray, materials, levels: int, name: str = "m", timeout: float | None = None):
    """The unmodified reference's buildSVO with putMaterialIdInLeaves = true (through oracle/ref_attr_driver.cpp):
    returns (mask (n,), material (n, 8)) of its leaf level, in its node order."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_svb_meshgen", HERE.parent / "svdag-compression_b200" / "meshgen.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    if not ATTR_BIN.exists():
        raise FileNotFoundError(f"{ATTR_BIN} missing: run `make -C oracle ref_attr` where /root/reference exists")
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    obj = wd / f"{name}.obj"
    mg.write_scene(obj, tris, materials=np.asarray(materials))
    out = wd / f"{name}_leaves.bin"
    p = subprocess.run([str(ATTR_BIN), str(obj), str(levels), str(out)], cwd=wd, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"ref_attr_driver failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    raw = out.read_bytes()
    n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    rec = np.frombuffer(raw[8:8 + 33 * n], dtype=np.dtype([("mask", "u1"), ("child", "<u4", (8,))]))
    return rec["mask"].copy(), rec["child"].copy()


# ------------------------------------------------------------------ DDA ray caster (oracle/dda_oracle.c)
_dda = None
FILE_KIND = {"svdag": 0, "ussvdag": 1, "ssvdag": 2, "esvdag": 2}


def dda_render(file_bytes: bytes, kind: str, view_inv, proj_inv, width: int, height: int,
               max_iters: int = 512, draw_level: int = 0, projection_factor: float = 0.0) -> np.ndarray:
    """Depth image (h, w, 3) = (t, level, iterations) of an encoded file, CPU restatement of the viewer's DEPTH_MODE
    shader.  view_inv / proj_inv: 4x4 float32, column-major flattened (what glUniformMatrix4fv receives)."""
    global _dda
    if _dda is None:
        if not DDA_LIB.exists():
            build(ref=False)
        _dda = C.CDLL(str(DDA_LIB))
        _dda.dda_oracle_render.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    vi = np.ascontiguousarray(view_inv, dtype=np.float32).reshape(16)
    pi = np.ascontiguousarray(proj_inv, dtype=np.float32).reshape(16)
    out = np.zeros((height, width, 3), np.float32)
    rc = _dda.dda_oracle_render(buf.ctypes.data, len(buf), FILE_KIND[kind], vi.ctypes.data, pi.ctypes.data, width, height,
                                max_iters, draw_level, float(projection_factor), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("dda_oracle_render failed (bad file?)")
    return out
