import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C-ABI.  On a GPU box a missing library is a hard failure."""
    import torch
    from plonky2_ecdsa_b200 import Engine
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container")
    eng = Engine(0)          # raises P2EError (test failure) if libp2e.so is absent
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def hostcheck():
    """Host-compiled copy of the device arithmetic cores (tests/hostcheck, test infrastructure)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "hostcheck.cpp")
    hdrs = [os.path.join(ROOT, "plonky2_ecdsa_b200", "csrc", h) for h in ("p2e_field.cuh", "p2e_f8.cuh")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    for f in ("hc_gl_mul", "hc_gl_add", "hc_gl_sub", "hc_gl_reduce128"):
        getattr(lib, f).restype = ctypes.c_uint64
        getattr(lib, f).argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    return lib
