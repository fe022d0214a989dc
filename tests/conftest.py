import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from hairmsnn_b200 import api
        return api.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: the C ABI reports
    # HM_ERR_CUDA and the tests assert on results.  Nothing to do here on purpose.
    return


@pytest.fixture(scope="session")
def probe():
    """Host build of the product's shared host/device headers (tests/cpu_probe.cpp)."""
    import ctypes
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcpu_probe.so")
    srcs = [os.path.join(ROOT, "tests", "cpu_probe.cpp"), os.path.join(ROOT, "hairmsnn_b200", "csrc", "hm_bvh_build.cpp")]
    deps = srcs + [os.path.join(ROOT, "hairmsnn_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "hairmsnn_b200", "csrc")) if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-mfma", "-pthread", "-shared",
                               "-I/usr/local/cuda/include"] + srcs + ["-o", so])
    return ctypes.CDLL(so)
