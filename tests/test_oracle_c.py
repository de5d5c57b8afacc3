"""Pins the C restatement (oracle/p2e_oracle.c, the CPU baseline) against the Python spec."""
import random

import numpy as np
import pytest

from oracle import cpu_baseline as CB
from oracle import spec as S
from plonky2_ecdsa_b200 import workload as W
from tests import util as U


@pytest.fixture(scope="module")
def lib():
    return CB.load()


@pytest.mark.parametrize("modulus", [0, 1])
def test_c_mul_nonnative_vs_spec(lib, modulus):
    x, y = W.mulmod_workload(1200, modulus, seed=3 + modulus)
    out, err = CB.mul_nonnative(lib, modulus, x, y, threads=2)
    eo, ee = U.expect_mul(x, y, modulus)
    assert np.array_equal(out, eo) and np.array_equal(err, ee)


@pytest.mark.parametrize("modulus", [0, 1])
def test_c_inverse_vs_spec(lib, modulus):
    M = S.MODULI[modulus]
    rng = random.Random(8 + modulus)
    xs = [1, 2, M - 1, M + 1, (1 << 256) - 1, 0, M, 1 << 256] + [rng.randrange(1 << 256) for _ in range(60)]
    x = W.ints_to_limbs(xs)
    out, err = CB.inv_nonnative(lib, modulus, x, threads=2)
    eo, ee = U.expect_inv(x, modulus)
    assert np.array_equal(out, eo) and np.array_equal(err, ee)


def test_c_gate_eval_vs_spec(lib):
    w = W.random_goldilocks(4, 53, 700)
    assert np.array_equal(CB.eval_mul_nonnative(lib, 0, w, threads=2), U.expect_eval_mul(w, 0))
    assert np.array_equal(CB.eval_mul_nonnative(lib, 1, w), U.expect_eval_mul(w, 1))
    w2 = W.random_goldilocks(5, 33, 700)
    assert np.array_equal(CB.eval_check_sum(lib, w2), U.expect_eval_check_sum(w2))


def test_c_fast_inverse_mode_gives_the_same_bits(lib):
    """The "best-CPU" baseline variant (extended Euclid instead of the Fermat ladder) must not change a witness bit."""
    rng = random.Random(77)
    for modulus in (0, 1):
        M = S.MODULI[modulus]
        xs = [1, 2, M - 1, M + 1, 0, M, 3] + [rng.randrange(1 << 256) for _ in range(200)]
        x = W.ints_to_limbs(xs)
        ref, e0 = CB.inv_nonnative(lib, modulus, x)
        lib.orc_set_inverse_mode(1)
        try:
            got, e1 = CB.inv_nonnative(lib, modulus, x)
        finally:
            lib.orc_set_inverse_mode(0)
        assert np.array_equal(ref, got) and np.array_equal(e0, e1)
    wl = W.ecdsa_workload(3, seed=5)
    args = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    a, ea = CB.ecdsa_witness(lib, *args)
    lib.orc_set_inverse_mode(1)
    try:
        b, eb = CB.ecdsa_witness(lib, *args)
    finally:
        lib.orc_set_inverse_mode(0)
    assert np.array_equal(a, b) and np.array_equal(ea, eb)
