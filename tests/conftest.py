import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "cad_golden.npz")
REFERENCE_DATA = "/root/reference/data"        # only present in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from rpmplv2_b200 import engine
        return engine.load_library().rvs_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle binding (built on demand with g++)."""
    from oracle import oracle_binding
    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def host_sim():
    """Device headers compiled for the host (logic check without a GPU); see tests/host_sim/geom_host.cpp."""
    d = os.path.join(ROOT, "tests", "host_sim")
    out = os.path.join(d, "_build", "libgeom_host.so")
    srcs = [os.path.join(d, "geom_host.cpp")] + [os.path.join(ROOT, "rpmplv2_b200", "csrc", f)
                                                   for f in ("rvs_geom.cuh", "rvs_fk.cuh", "rvs_robot_host.h", "rvs_types.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                        srcs[0], "-o", out], check=True)
    return ctypes.CDLL(out)


def same(a, b):
    """Bitwise-style equality that treats NaN == NaN."""
    a, b = np.asarray(a), np.asarray(b)
    return (a == b) | (np.isnan(a) & np.isnan(b))
