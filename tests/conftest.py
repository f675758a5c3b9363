"""pytest configuration: the `gpu` marker, the package loader (the package
directory has a dash in its name) and the oracle / reference fixtures.

Only tests/ (here), __graft_entry__.smoke() and bench.py's cpu_baseline leg may
touch oracle/.
"""
import importlib.util
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(Path(__file__).resolve().parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    name = "sigmod2018_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "sigmod-2018_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def b200():
    """The ctypes binding; loading it does not need a GPU."""
    return load_package()


@pytest.fixture(scope="session")
def gpu(b200):
    """The binding with device 0 initialised (gpu tests only)."""
    rc = b200.lib().b200_init(0)
    assert rc == 0
    return b200


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def ref():
    """The UNMODIFIED reference compiled by oracle/Makefile (oracle/_ref/);
    present in the build container and on boxes the snapshot travelled to."""
    import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libref_ops.so not built (needs /root/reference)")
    return refbind
