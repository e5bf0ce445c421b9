"""Build libsvb.so (CUDA kernels + C ABI) and the svbuilder host tool for sm_100a, in-tree.

    python svdag-compression_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands next to this file so it travels to the GPU
box with the repo snapshot (it is git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsvb.so"
SVBUILDER = HERE / "svbuilder"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"   # the environment's CXX points at a compiler without libgomp; pin the system one

CU_SOURCES = ["svb_prims.cu", "svb_voxelize.cu", "svb_dedup.cu", "svb_sdag.cu", "svb_cross.cu", "svb_encode.cu", "svb_attr.cu", "svb_api.cu", "svb_raycast.cu", "host/encoders.cpp"]
# the ray caster must round every float operation separately (pixel-exact against oracle/dda_oracle.c)
EXTRA_FLAGS = {"svb_raycast.cu": ["--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false"]}
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp,-Wall,-Wno-unused-function,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [CSRC / "host" / "encoders.cpp", CSRC / "host" / "octree_data.hpp"] + [HERE.parent / "include" / "svb.h", Path(__file__)]
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    procs = []
    for src in CU_SOURCES:
        obj = objdir / (src.replace("/", "_") + ".o")
        objs.append(str(obj))
        cmd = [NVCC] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-ccbin", HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + objs + ["-lcudart", "-Xcompiler", "-fopenmp"]
    subprocess.run(cmd, check=True)
    return LIB


def build_svbuilder(force: bool = False) -> Path:
    host = CSRC / "host"
    srcs = sorted(p for p in host.glob("*.cpp") if p.name != "encoders.cpp")   # encoders.cpp lives in libsvb.so
    if not srcs:
        return SVBUILDER
    deps = srcs + list(host.glob("*.hpp")) + [LIB]
    if not force and not _stale(SVBUILDER, deps):
        return SVBUILDER
    cmd = [HOSTCXX, "-O2", "-std=c++14", "-ffp-contract=off", "-Wall", "-I", str(HERE.parent / "include"), "-I", str(host)]
    cmd += [str(s) for s in srcs] + ["-o", str(SVBUILDER), f"-L{HERE}", "-lsvb", f"-Wl,-rpath,{HERE}", "-Wl,-rpath,$ORIGIN",
                                     "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-I/usr/local/cuda/include",
                                     "-lcudart", "-lnccl", "-pthread"]   # NCCL only in the tool (multi-GPU from one process); libsvb.so itself is NCCL-free
    subprocess.run(cmd, check=True)
    return SVBUILDER


def build_all(force: bool = False, verbose: bool = False):
    build_lib(force, verbose)
    build_svbuilder(force)
    return LIB


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
