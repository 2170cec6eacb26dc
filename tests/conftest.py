"""pytest wiring: the `gpu` marker, import paths and shared fixtures."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "voxelengine-cpp_b200"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _ensure_port_oracle():
    lib = os.path.join(ROOT, "oracle", "libvxoracle.so")
    src = os.path.join(ROOT, "oracle", "vxoracle.cpp")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"],
                              stdout=subprocess.DEVNULL)


def _ensure_reference_oracle():
    """Built only where /root/reference exists (never on the GPU box, which
    uses the prebuilt oracle/_ref/libvxref.so that travelled with the repo)."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libvxref.so")
    if not os.path.exists(lib) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "reference"],
                              stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session", autouse=True)
def _oracles():
    _ensure_port_oracle()
    _ensure_reference_oracle()


@pytest.fixture(scope="session")
def table():
    import cases
    return cases.table()
