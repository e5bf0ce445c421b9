"""CPU check of the voxelizer's per-pair decision (svb_classify.cuh compiled with g++, see
tests/harness/classify_harness.cpp): the FP64 interval filter, the inherited settled-axis flags, the exact
single-axis fast path for flat triangles and the closed-form node centres must give, for EVERY (triangle, node)
pair of a hierarchical build, the 8-child mask of the reference-order predicate.  Test infrastructure only --
the product never runs this code on the CPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "harness" / "classify_harness.cpp"
OUT = ROOT / "tests" / "harness" / "build" / "libclassify_harness.so"


# Two builds of the same source.  "separate": every operation rounded on its own (-ffp-contract=off).  "contracted": the
# filter's plain a * b + c expressions fused into FMAs (-ffp-contract=fast -mfma), which is how nvcc compiles them for the
# device, while the explicitly rounded operations of the predicate and the centre chain stay single operations
# (SVB_HOST_STRICT_OPS, svb_sat.cuh).  The filter's 2^-40 margin must make its verdicts independent of that choice.
@pytest.fixture(scope="module", params=["separate", "contracted"])
def harness(request):
    OUT.parent.mkdir(exist_ok=True)
    contracted = request.param == "contracted"
    if contracted and " fma " not in Path("/proc/cpuinfo").read_text().replace("\n", " "):
        pytest.skip("host CPU without FMA")
    out = OUT.with_name("libclassify_harness_fma.so") if contracted else OUT
    deps = [SRC] + list((ROOT / "svdag-compression_b200" / "csrc").glob("svb_*.cuh"))
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        flags = ["-ffp-contract=fast", "-mfma", "-DSVB_HOST_STRICT_OPS"] if contracted else ["-ffp-contract=off"]
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17"] + flags + ["-fPIC", "-shared", "-Wno-unknown-pragmas",
                        str(SRC), "-o", str(out)], check=True)
    if contracted:   # the build really contains fused multiply-adds (and the predicate's operations really are out of line)
        dis = subprocess.run(["objdump", "-d", "--no-show-raw-insn", str(out)], capture_output=True, text=True).stdout
        assert dis.count("vfmadd") + dis.count("vfmsub") + dis.count("vfnmadd") > 20, "no FMA contraction in the contracted harness"
    L = C.CDLL(str(out))
    L.harness_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.harness_chain_exact.argtypes = [C.c_void_p, C.c_double, C.c_int]
    return L


def _run(L, tris, lo, hi, levels, direct, flat_only=False):
    tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
    bf = np.concatenate([lo, hi]).astype(np.float32)                      # geom_octree.cpp:177-184
    side = max(np.float32(np.float32(np.float32(bf[3 + k] - bf[k]) * np.float32(0.5)) * np.float32(2.0)) for k in range(3))
    centre = np.ascontiguousarray((np.asarray(lo, np.float64) + np.asarray(hi, np.float64)) * 0.5)   # bbox.center(), :214
    out = np.zeros(10, np.uint64)
    L.harness_run(tris.ctypes.data, tris.shape[0], centre.ctypes.data, float(side), levels, int(direct), int(flat_only), out.ctypes.data)
    return dict(pairs=int(out[0]), bad=int(out[1]), fast=int(out[2]), exact=int(out[3]),
                first=dict(level=int(out[4]), tri=int(out[5]), code=int(out[6]), got=int(out[7]), want=int(out[8]), fl=int(out[9])))


def _affine(tris, scale, offset):
    t = tris.reshape(-1, 3).astype(np.float64) * np.asarray(scale) + np.asarray(offset)
    return np.ascontiguousarray(t.astype(np.float32).reshape(-1, 9))


CASES = [
    ("city", dict(lots=8), 8),
    ("terrain", dict(n=32), 8),
    ("sphere", dict(n_lat=16, n_lon=32), 8),
    ("sphere_menger", dict(n_lat=16, n_lon=32, sponge_level=2), 7),
    ("soup", dict(n=400, seed=7), 8),       # degenerate triangles, lattice ties, scene-spanning triangles
    ("soup", dict(n=240, seed=11), 9),
]


@pytest.mark.parametrize("direct", [True, False], ids=["direct", "chain"])
@pytest.mark.parametrize("mesh,kw,levels", CASES, ids=[f"{c[0]}{i}" for i, c in enumerate(CASES)])
def test_filter_equals_predicate_unit_cube(harness, meshgen, mesh, kw, levels, direct):
    tris = meshgen.make_mesh(mesh, **kw)
    v = tris.reshape(-1, 3).astype(np.float64)
    r = _run(harness, tris, v.min(axis=0), v.max(axis=0), levels, direct)
    assert r["bad"] == 0, r
    if mesh == "city":
        assert r["fast"] > 0.1 * r["pairs"], r   # the flat fast path must actually be exercised
        r2 = _run(harness, tris, v.min(axis=0), v.max(axis=0), levels, direct, flat_only=True)   # box-mesh kernel variant
        assert r2["bad"] == 0 and r2["pairs"] == r["pairs"], r2


@pytest.mark.parametrize("direct", [True, False], ids=["direct", "chain"])
@pytest.mark.parametrize("mesh,kw,levels", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_filter_equals_predicate_arbitrary_float_bbox(harness, meshgen, mesh, kw, levels, direct):
    """Scaled / shifted scene: centres are (lo+hi)/2 of float bounds -- the chain is still exact, so the
    closed form must agree with it."""
    tris = _affine(meshgen.make_mesh(mesh, **kw), (3.7, 2.9, 5.3), (11.3, -5.1, 2.9))
    v = tris.reshape(-1, 3).astype(np.float64)
    r = _run(harness, tris, v.min(axis=0), v.max(axis=0), levels, direct)
    assert r["bad"] == 0, r


def test_filter_equals_predicate_full_double_bbox(harness, meshgen):
    """Caller-supplied bbox with 53-bit doubles: the chain rounds at every level; the kernels then replay it
    (DIRECT=false, chosen on the host by centre_chain_exact)."""
    tris = _affine(meshgen.make_mesh("city", lots=8), (3.7, 2.9, 5.3), (11.3, -5.1, 2.9))
    v = tris.reshape(-1, 3).astype(np.float64)
    lo, hi = v.min(axis=0) - np.array([0.1, 0.2, 0.3]) / 3.0, v.max(axis=0) + np.array([0.7, 0.1, 0.05]) / 7.0
    r = _run(harness, tris, lo, hi, 8, False)
    assert r["bad"] == 0, r


def _oblique_flat_triangles(n, seed, lattice):
    """Flat triangles whose three in-plane edges are all oblique (a box mesh only ever leaves ONE edge axis unsettled: the
    hypotenuse's), on every flat axis; vertices on a 1/lattice grid so that edges pass exactly through voxel corners
    (ties: the reference-order predicate has to decide them) or, lattice = 0, anywhere."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        a = i % 3
        u, w = [k for k in range(3) if k != a]
        while True:
            if lattice:
                p = rng.integers(0, lattice + 1, size=(3, 2)).astype(np.float64) / lattice
            else:
                p = rng.random((3, 2))
            e = np.array([p[1] - p[0], p[2] - p[1], p[0] - p[2]])
            area = abs(e[0][0] * e[1][1] - e[0][1] * e[1][0])
            if area > 1e-3 and np.all(np.abs(e) > 1e-6):      # not degenerate, no axis-aligned edge
                break
        out[i, :, a] = (rng.integers(0, 65) / 64.0) if lattice else rng.random()
        out[i, :, u] = p[:, 0]
        out[i, :, w] = p[:, 1]
    return np.ascontiguousarray(out.reshape(n, 9))


@pytest.mark.parametrize("direct", [True, False], ids=["direct", "chain"])
@pytest.mark.parametrize("lattice", [64, 0], ids=["lattice", "random"])
def test_flat_triangles_with_three_oblique_edges(harness, lattice, direct):
    """All three edge axes of a flat triangle unsettled at once: the three-edge loop of slow_leaf_voxels / the three edge
    blocks of classify_flat_slow, voxel by voxel against the predicate; the lattice variant is full of exact ties."""
    tris = _oblique_flat_triangles(90, 5 if lattice else 6, lattice)
    lo, hi = np.zeros(3), np.ones(3)
    for flat_only in (1, 0):
        r = _run(harness, tris, lo, hi, 7, direct, flat_only=flat_only)
        assert r["bad"] == 0, r
        assert r["pairs"] > 100000
        if lattice:
            assert r["exact"] > 0, r


def _chain_partial_sums_exact(centre, root_side, levels):
    """Brute force with exact rational arithmetic: is every partial sum c0 + sum(+-k_j) a representable double?"""
    from fractions import Fraction
    for c0 in centre:
        for signs in ((1,) * (levels + 1), (-1,) * (levels + 1), tuple((-1) ** j for j in range(levels + 1))):
            acc, k = Fraction(c0), Fraction(root_side) / 4
            for j in range(levels + 1):        # levels steps of the chain + the child step
                acc += signs[j] * k
                if Fraction(float(acc)) != acc:
                    return False
                k /= 2
    return True


@pytest.mark.parametrize("centre,side,levels", [
    ((0.5, 0.5, 0.5), 1.0, 16),                                   # unit cube, 64K^3
    ((13.149999618530273, -3.6499998569488525, 5.550000190734863), 5.300000190734863, 14),   # float scene bbox
    ((0.1, 0.2, 0.3), 1.0, 9),                                    # full 53-bit doubles: the chain rounds
    ((1e6 + 1 / 3.0, 2.0, 3.0), 0.37, 12),                        # large offset + full mantissa
    ((12345.678, -0.001, 7.0), 1e-3, 10),                         # tiny cube far from the origin
    ((0.0, 0.0, 0.0), 3.0, 20),
])
def test_chain_exactness_predicate_is_sound(harness, centre, side, levels):
    """centre_chain_exact() may say "no" too often, never "yes" wrongly: whenever it allows closed-form centres, every
    partial sum of the chain must be a representable double (checked with exact rationals)."""
    c = np.ascontiguousarray(centre, dtype=np.float64)
    says = bool(harness.harness_chain_exact(c.ctypes.data, float(np.float32(side)), levels))
    truth = _chain_partial_sums_exact([float(x) for x in c], float(np.float32(side)), levels)
    assert not (says and not truth), (centre, side, levels)
    if centre == (0.5, 0.5, 0.5):
        assert says
    if centre == (0.1, 0.2, 0.3):
        assert not says
