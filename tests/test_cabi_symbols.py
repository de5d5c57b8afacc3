"""CPU tier: the C-ABI library loads and exports every symbol include/p2e.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "p2e.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2e_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import plonky2_ecdsa_b200 as pkg
    if not os.path.exists(pkg.LIB_PATH):
        from plonky2_ecdsa_b200.build import build
        build()
    lib = pkg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/p2e.h but not exported by libp2e.so"
        assert s in pkg.SIGNATURES, f"{s} has no ctypes signature in engine.py"
    assert sorted(pkg.SIGNATURES) == syms


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry must fail loudly (never route to a CPU path)."""
    import torch
    import plonky2_ecdsa_b200 as pkg
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    with pytest.raises(pkg.P2EError):
        pkg.Engine(0)
    lib = pkg.load_library()
    assert lib.p2e_mul_nonnative_witness(0, 1, None, None, None, None) != 0
    assert b"p2e_init" in lib.p2e_last_error()


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "plonky2_ecdsa_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no CPU", ""), f"{f} mentions the oracle"
