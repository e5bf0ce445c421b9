"""pytest configuration: `gpu` marker, repo-root imports, shared fixtures."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_pkg():
    """Import the product package.  Its directory name has a hyphen (svdag-compression_b200),
    so it is loaded by path under the module name `svdag_compression_b200`."""
    name = "svdag_compression_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "svdag-compression_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py", submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def meshgen(pkg):
    return pkg.meshgen


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as o
    o.lib()
    return o


GOLDEN = sorted(p for p in (ROOT / "tests" / "golden").glob("*.npz") if not p.name.startswith("attr_"))   # (attr_*.npz: tests/test_attributes.py)


def golden_case(path):
    z = np.load(path)
    files = {k[5:]: z[k].tobytes() for k in z.files if k.startswith("file_")}
    return {"name": Path(path).stem, "tris": z["tris"], "levels": int(z["levels"]), "step": int(z["step"]),
            "cross": bool(z["cross"]), "reduced": z["reduced"], "stats": z["stats"], "files": files}
