"""CPU check of the permute-based key builder of the 4^3 dedup level (svb_dedup.cu::build_key64_u8): the selector table is
parsed from the CUDA source and the funnel-shift / byte-permute / byte-mask sequence is emulated with numpy for every
child mask, every alignment of childBase and random child bytes, against the byte-by-byte definition of the key
(build_key: byte c of the key = voxel mask of child c, children stored contiguously in ascending child order)."""
import re
from pathlib import Path

import numpy as np

SRC = Path(__file__).resolve().parents[1] / "svdag-compression_b200" / "csrc" / "svb_dedup.cu"


def _rank_sel():
    txt = SRC.read_text()
    body = re.search(r"RANK_SEL\[128\]\s*=\s*\{(.*?)\};", txt, re.S).group(1)
    vals = [int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{8})u", body)]
    assert len(vals) == 128
    return vals


def _byte_perm(x, y, s):
    """CUDA __byte_perm(x, y, s): result byte i = byte (s >> 4i) & 7 of the 8-byte pool {x bytes 0-3, y bytes 0-3}."""
    pool = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(pool[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def _funnelshift_r(lo, hi, sh):
    return (((hi << 32) | lo) >> (sh & 31)) & 0xFFFFFFFF


def _key_perm(mask, base, refs, sel):
    w = base >> 2
    word = lambda i: int.from_bytes(bytes(refs[4 * i: 4 * i + 4]), "little")
    a0, a1, a2 = word(w), word(w + 1), word(w + 2)
    sh = (base & 3) * 8
    lo, hi = _funnelshift_r(a0, a1, sh), _funnelshift_r(a1, a2, sh)
    s = sel[mask & 127]
    bm_lo = ((((mask & 15) * 0x00204081) & 0x01010101) * 0xFF) & 0xFFFFFFFF
    bm_hi = ((((mask >> 4) * 0x00204081) & 0x01010101) * 0xFF) & 0xFFFFFFFF
    k_lo = _byte_perm(lo, hi, s & 0xFFFF) & bm_lo
    k_hi = _byte_perm(lo, hi, s >> 16) & bm_hi
    return (k_hi << 32) | k_lo


def _key_plain(mask, base, refs):
    key, r = 0, 0
    for c in range(8):
        if (mask >> c) & 1:
            key |= int(refs[base + r]) << (8 * c)
            r += 1
    return key


def test_rank_selector_table():
    sel = _rank_sel()
    for m in range(256):
        for c in range(8):
            assert (sel[m & 127] >> (4 * c)) & 0xF == bin(m & ((1 << c) - 1)).count("1"), (m, c)


def test_permute_key_equals_bytewise_key():
    sel = _rank_sel()
    rng = np.random.default_rng(5)
    for m in range(256):
        for base in (0, 1, 2, 3, 5, 10, 23):
            refs = rng.integers(0, 256, size=48, dtype=np.uint8)   # neighbours' bytes beyond the node's run are arbitrary
            refs[rng.integers(0, 48, size=6)] = 0                  # empty leaf children contribute a zero byte = no child
            assert _key_perm(m, base, refs, sel) == _key_plain(m, base, refs), (m, base)
