"""The C-ABI shared library loads (no GPU needed for dlopen) and exports every entry point include/svb.h declares;
compute entry points fail loudly -- not silently on a CPU path -- when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "svb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.svb_version().decode().startswith("svb ")


def test_no_device_means_no_context(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = pkg.lib()
    assert not L.svb_create(0)                      # NULL: there is no CPU fallback
    with pytest.raises(pkg.SvbError):
        pkg.GeomOctree()
