"""DDA depth-image checker: the CPU restatement of the viewer's DEPTH_MODE shader
(oracle/dda_oracle.c) against the encoded files' geometry, and the CUDA ray caster (svb_raycast_depth) against
the restatement, pixel-exact."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case

W = H = 96
CASES = ["city_L7_s2", "sphere_L7_s2", "terrain_L7_s3", "city_affine_L8_s2", "spongeball_L7_s0"]
CROSS = ["city_L7_s1_c", "sphere_L6_s0_c", "spongeball_L7_s2_c", "city_affine_L7_s1_c"]
KIND = {"svdag": "svdag", "multi_svdag": "svdag", "ussvdag": "ussvdag", "ssvdag": "ssvdag", "esvdag": "esvdag"}


def _case(name):
    return golden_case([p for p in GOLDEN if p.stem == name][0])


def _cameras(pkg, tris):
    v = tris.reshape(-1, 3)
    lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
    c, d = (lo + hi) / 2, float(np.linalg.norm(hi - lo))
    cams = []
    for off, fov in (((0.9, 0.7, 0.6), 50.0), ((-0.3, 1.1, 0.25), 35.0), ((0.05, 0.02, 0.45), 70.0)):   # the last one sits inside the bbox
        cams.append((pkg.camera.look_at_inv(c + np.array(off) * d, c), pkg.camera.perspective_inv(fov, 1.0), fov))
    return cams


def _occupancy(svdag_bytes):
    """Dense occupancy grid decoded from a .svdag image (encoded_svdag.cpp:155-170: mask word + child words 7..0)."""
    hdr = np.frombuffer(svdag_bytes[:44], dtype=np.uint32)
    levels = int(hdr[7])
    words = np.frombuffer(svdag_bytes[44:], dtype=np.uint32)
    n = 1 << levels
    grid = np.zeros((n, n, n), bool)

    def rec(ptr, lev, x, y, z):
        m = int(words[ptr]) & 0xFF
        size = n >> (lev + 1)
        for c in range(8):
            if not (m >> c) & 1:
                continue
            cx, cy, cz = x + ((c >> 2) & 1) * size, y + ((c >> 1) & 1) * size, z + (c & 1) * size
            if lev == levels - 1:
                grid[cx, cy, cz] = True
            else:
                rec(int(words[ptr + bin(m >> c).count("1")]), lev + 1, cx, cy, cz)
    rec(0, 0, 0, 0, 0)
    return grid, levels


@pytest.mark.parametrize("name", CASES[:3])
def test_oracle_hits_lie_on_set_voxels(pkg, orc, name):
    """Semantics of the restated DDA: every reported hit is the entry point of a set voxel of the decoded grid."""
    g = _case(name)
    data = g["files"]["svdag"]
    grid, levels = _occupancy(data)
    hdrf = np.frombuffer(data[:28], dtype=np.float32)
    bbmin, bbmax, root = hdrf[0:3].astype(np.float64), hdrf[3:6].astype(np.float64), float(hdrf[6])
    omin = (bbmin + bbmax) / 2 - root / 2
    n = 1 << levels
    for vi, pi, _fov in _cameras(pkg, g["tris"])[:2]:
        img = orc.dda_render(data, "svdag", vi, pi, W, H, 2000, 0, 1e30)
        hit = img[..., 0] > 0
        assert hit.mean() > 0.05
        m = np.asarray(vi, np.float64).reshape(4, 4).T
        pinv = np.asarray(pi, np.float64).reshape(4, 4).T
        ys, xs = np.nonzero(hit)
        bad = 0
        for y, x in zip(ys, xs):
            sx, sy = (x + 0.5) / W * 2 - 1, (y + 0.5) / H * 2 - 1
            a, b = pinv @ [sx, sy, 0, 1], pinv @ [sx, sy, 1, 1]
            d = b[:3] / b[3] - a[:3] / a[3]
            d /= np.linalg.norm(d)
            o = (m @ [0, 0, 0, 1])[:3]
            e = (m @ [d[0], d[1], d[2], 1])[:3]
            d = (e - o) / np.linalg.norm(e - o)
            p = o + (float(img[y, x, 0]) + 0.02 * root / n) * d        # a fiftieth of a voxel past the entry point
            v = np.floor((p - omin) / root * n).astype(int)
            if (v < 0).any() or (v >= n).any() or not grid[v[0], v[1], v[2]]:
                bad += 1
        assert bad <= 0.01 * len(ys), f"{bad} of {len(ys)} hit points are not inside a set voxel"


@pytest.mark.parametrize("name", CASES + CROSS)
def test_oracle_images_agree_across_formats(pkg, orc, name):
    """.svdag == .ussvdag == -multi.svdag pixel for pixel (same traversal arithmetic, different pointers / mirror
    bits); .ssvdag / .esvdag walk 4^3 leaves, so t may differ in the last ulps but never the hit set."""
    g = _case(name)
    for vi, pi, fov in _cameras(pkg, g["tris"]):
        for pf in (1e30, pkg.camera.projection_factor(fov, H)):
            ref = orc.dda_render(g["files"]["svdag"], "svdag", vi, pi, W, H, 2000, 0, pf)
            for k, data in g["files"].items():
                if k == "svdag":
                    continue
                img = orc.dda_render(data, KIND[k], vi, pi, W, H, 2000, 0, pf)
                if k in ("ussvdag", "multi_svdag"):
                    assert np.array_equal(img, ref), (name, k)
                elif pf > 1e29:   # with LOD on, the 4^3 leaves stop one level earlier than 2^3 nodes: compare only at full depth
                    assert np.array_equal(img[..., 0] > 0, ref[..., 0] > 0), (name, k)
                    assert np.abs(img[..., 0] - ref[..., 0]).max() <= 1e-5 * max(1.0, float(ref[..., 0].max())), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES + CROSS)
def test_cuda_raycast_equals_oracle_pixel_exact(pkg, orc, name):
    g = _case(name)
    for vi, pi, fov in _cameras(pkg, g["tris"]):
        for pf in (1e30, pkg.camera.projection_factor(fov, H)):
            for k, data in g["files"].items():
                want = orc.dda_render(data, KIND[k], vi, pi, W, H, 2000, 0, pf)
                got = pkg.raycast_depth(data, KIND[k], vi, pi, W, H, 2000, 0, pf)
                assert np.array_equal(got, want), f"{name} {k}: {(got != want).any(axis=2).sum()} pixels differ"


@pytest.mark.gpu
def test_cuda_raycast_gpu_built_vs_reference_built_files(pkg, meshgen):
    """north_star sanity check: depth image of the GPU-built outputs == depth image of the reference-built files."""
    g = _case("city_L7_s2")
    t = pkg.GeomOctree(g["tris"])
    t.build(g["levels"], g["step"])
    mine = {"svdag": pkg.encoders.encode(t, "svdag")}
    t.to_sdag()
    mine["ussvdag"] = pkg.encoders.encode(t, "ussvdag")
    mine["ssvdag"] = pkg.encoders.encode(t, "ssvdag")
    vi, pi, _ = _cameras(pkg, g["tris"])[0]
    imgs = {}
    for k in mine:
        a = pkg.raycast_depth(mine[k], k, vi, pi, 256, 256, 2000)
        b = pkg.raycast_depth(g["files"][k], k, vi, pi, 256, 256, 2000)
        assert np.array_equal(a, b), k
        imgs[k] = a
    assert (imgs["svdag"][..., 0] > 0).mean() > 0.1
    assert np.array_equal(imgs["svdag"], imgs["ussvdag"])
    assert np.array_equal(imgs["svdag"][..., 0] > 0, imgs["ssvdag"][..., 0] > 0)


@pytest.mark.gpu
def test_cuda_raycast_large_scene_formats_agree(pkg, meshgen):
    """1024^3 city built on the GPU: SVDAG, USSVDAG, -multi.svdag images identical; SSVDAG same hit set."""
    tris = meshgen.make_mesh("city", lots=16)
    a = pkg.GeomOctree(tris)
    a.build(10, 2)
    sv = pkg.encoders.encode(a, "svdag")
    b = pkg.GeomOctree(tris)
    b.build(10, 2)
    b.cross_merge()
    multi = pkg.encoders.encode(b, "svdag")
    a.to_sdag()
    us, ss = pkg.encoders.encode(a, "ussvdag"), pkg.encoders.encode(a, "ssvdag")
    vi, pi, _ = _cameras(pkg, tris)[0]
    ref = pkg.raycast_depth(sv, "svdag", vi, pi, 512, 512, 4000)
    assert (ref[..., 0] > 0).mean() > 0.1
    assert np.array_equal(pkg.raycast_depth(multi, "svdag", vi, pi, 512, 512, 4000), ref)
    assert np.array_equal(pkg.raycast_depth(us, "ussvdag", vi, pi, 512, 512, 4000), ref)
    s4 = pkg.raycast_depth(ss, "ssvdag", vi, pi, 512, 512, 4000)
    assert np.array_equal(s4[..., 0] > 0, ref[..., 0] > 0)
    assert np.abs(s4[..., 0] - ref[..., 0]).max() <= 1e-5 * float(ref[..., 0].max())


def test_raycast_without_device_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    g = _case("sphere_L6_s0")
    vi, pi, _ = _cameras(pkg, g["tris"])[0]
    with pytest.raises(pkg.SvbError):
        pkg.raycast_depth(g["files"]["svdag"], "svdag", vi, pi, 16, 16)
