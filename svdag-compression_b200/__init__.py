"""svdag-compression_b200 -- B200-native svbuilder hot path (mesh -> SVO -> SVDAG -> SSVDAG).

The product is `libsvb.so` (hand-written CUDA for sm_100a behind the C ABI of include/svb.h)
plus the C++ host tool `svbuilder`.  This Python package is only the thin ctypes binding used
by tests and bench.py; it contains no compute and no CPU fallback: without the CUDA library
and a GPU every build call raises.

The directory name carries a hyphen, so import it by path (see tests/conftest.py::load_pkg)
under the module name `svdag_compression_b200`.
"""
from . import meshgen  # noqa: F401
from . import build as _build  # noqa: F401
from .capi import SvbError, GeomOctree, lib, lib_path, STATE_NAMES, raycast_depth  # noqa: F401
from . import encoders  # noqa: F401
from . import sharded  # noqa: F401
from . import camera  # noqa: F401

__all__ = ["meshgen", "GeomOctree", "SvbError", "lib", "lib_path", "encoders", "build_native"]


def build_native(force: bool = False, verbose: bool = False):
    """Compile libsvb.so (+ svbuilder) in-tree for sm_100a."""
    return _build.build_all(force=force, verbose=verbose)
