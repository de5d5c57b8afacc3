"""CPU tier: the device arithmetic cores of csrc/p2e_field.cuh, compiled for the host by
tests/hostcheck, against the oracle.  (The CUDA kernels run the very same functions.)"""
import ctypes
import random

import numpy as np
import pytest

from oracle import spec as S
from plonky2_ecdsa_b200 import workload as W
from tests import util as U

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)


def P(a):
    return a.ctypes.data_as(u64p if a.dtype == np.uint64 else u32p)


def aos(m):
    """wire-major [rows][n] -> instance-major [n][rows] (hostcheck's layout)."""
    return np.ascontiguousarray(m.T)


@pytest.mark.parametrize("modulus", [0, 1])
def test_mul_witness_core(hostcheck, modulus):
    n = 1500
    x, y = W.mulmod_workload(n, modulus, seed=W.SEED + 17 * modulus)
    out = np.zeros((n, 51), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_mul_witness(modulus, ctypes.c_size_t(n), P(aos(x)), P(aos(y)), P(out), P(err))
    eo, ee = U.expect_mul(x, y, modulus)
    assert np.array_equal(out.T, eo)
    assert np.array_equal(err, ee)
    assert (ee != 0).any() and (ee == 0).any()      # both valid rows and would-panic rows covered


@pytest.mark.parametrize("modulus", [0, 1])
def test_mul_wire_checks(hostcheck, modulus):
    x = np.zeros((3, 9), dtype=np.uint64)
    y = np.ones((3, 9), dtype=np.uint64)
    x[0, 0] = 1 << 29
    x[1, 3] = 1 << 32
    x[2, 8] = (1 << 64) - (1 << 32)       # p_GL - 1
    out = np.zeros((3, 51), dtype=np.uint64)
    err = np.zeros(3, dtype=np.uint32)
    hostcheck.hc_mul_witness(modulus, ctypes.c_size_t(3), P(x), P(y), P(out), P(err))
    assert list(err) == [2, 1, 1] and not out.any()


@pytest.mark.parametrize("modulus", [0, 1])
@pytest.mark.parametrize("is_sub", [0, 1])
def test_addsub_core(hostcheck, modulus, is_sub):
    M = S.MODULI[modulus]
    rng = random.Random(100 + modulus * 2 + is_sub)
    vals = [0, 1, M - 1, M, (1 << 256) - 1, M + 1, 2, M // 2, M // 2 + 1]
    pairs = [(a, b) for a in vals for b in vals] + [(M - 5, 5), (5, M - 5), ((1 << 256) + 3, 5), (7, (1 << 261) - 1)]
    pairs += [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(800)]
    a = W.ints_to_limbs([p[0] for p in pairs])
    b = W.ints_to_limbs([p[1] for p in pairs])
    n = len(pairs)
    out = np.zeros((n, 10), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_addsub(modulus, is_sub, ctypes.c_size_t(n), P(aos(a)), P(aos(b)), P(out), P(err))
    eo, ee = U.expect_addsub(a, b, modulus, is_sub)
    assert np.array_equal(out.T, eo) and np.array_equal(err, ee)


@pytest.mark.parametrize("modulus", [0, 1])
def test_inverse_core(hostcheck, modulus):
    M = S.MODULI[modulus]
    rng = random.Random(200 + modulus)
    xs = [1, 2, M - 1, M + 1, (1 << 256) - 1, 0, M, 1 << 256] + [rng.randrange(1 << 256) for _ in range(150)]
    x = W.ints_to_limbs(xs)
    n = len(xs)
    out = np.zeros((n, 18), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_inv(modulus, ctypes.c_size_t(n), P(aos(x)), P(out), P(err))
    eo, ee = U.expect_inv(x, modulus)
    assert np.array_equal(out.T, eo) and np.array_equal(err, ee)


def test_glv_core(hostcheck):
    rng = random.Random(300)
    n_ = S.SECP_N
    ks = [0, 1, n_ - 1, n_, n_ // 2, n_ // 2 + 1, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(1500)]
    k = W.ints_to_limbs(ks)
    n = len(ks)
    out = np.zeros((n, 12), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_glv(ctypes.c_size_t(n), P(aos(k)), P(out), P(err))
    eo, ee = U.expect_glv(k)
    assert np.array_equal(out.T, eo) and np.array_equal(err, ee)


def test_check_sum_core(hostcheck):
    x, y = W.mulmod_workload(600, 0, seed=77)
    eo, _ = U.expect_mul(x, y, 0)
    a = np.ascontiguousarray(eo[18:35])
    n = a.shape[1]
    # add rows of arbitrary field elements (always the slow path)
    a[:, :40] = W.random_goldilocks(5, 17, 40)
    b = np.zeros((n, 16), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_check_sum(ctypes.c_size_t(n), P(aos(a)), P(b), P(err))
    for i in range(n):
        eb, panic = S.check_sum_gen_nopanic(U.col(a, i))
        assert [int(v) for v in b[i]] == eb and bool(err[i] & 16) == panic


def test_goldilocks_ops(hostcheck):
    rng = random.Random(9)
    p = S.GL_P
    edge = [0, 1, p - 1, p - 2, 1 << 32, (1 << 32) - 1, (1 << 63), p >> 1]
    cases = [(a, b) for a in edge for b in edge] + [(rng.randrange(p), rng.randrange(p)) for _ in range(5000)]
    for a, b in cases:
        assert hostcheck.hc_gl_mul(a, b) == a * b % p
        assert hostcheck.hc_gl_add(a, b) == (a + b) % p
        assert hostcheck.hc_gl_sub(a, b) == (a - b) % p
    for _ in range(2000):
        lo, hi = rng.getrandbits(64), rng.getrandbits(64)
        assert hostcheck.hc_gl_reduce128(lo, hi) == ((hi << 64) | lo) % p


# ---------------------------------------------------------------------------- saturated 8 x 32-bit field (p2e_f8.cuh)
P_SECP = S.MODULI[0]


def _words(vals, width=9):
    a = np.zeros((len(vals), width), dtype=np.uint32)
    for i, v in enumerate(vals):
        for k in range(8):
            a[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return a


def _from_words(a):
    return [sum(int(a[i, k]) << (32 * k) for k in range(8)) for i in range(a.shape[0])]


F8_EDGES = [0, 1, 2, P_SECP - 1, P_SECP, P_SECP + 1, (1 << 256) - 1, (1 << 256) - 2, (1 << 32) + 977, (1 << 32) + 976,
            (1 << 255), (1 << 255) - 1, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1, ((1 << 256) - 1) ^ 0xFFFFFFFF]


@pytest.mark.parametrize("op", range(9))
def test_f8_portable_ops(hostcheck, op):
    rng = random.Random(4000 + op)
    pairs = [(a, b) for a in F8_EDGES for b in F8_EDGES] + [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(3000)]
    a, b = _words([p[0] for p in pairs]), _words([p[1] for p in pairs])
    out = np.zeros_like(a)
    hostcheck.hc_f8(op, ctypes.c_size_t(len(pairs)), P(a), P(b), P(out))
    got = _from_words(out)
    for (x, y), g in zip(pairs, got):
        want = [x * y, x * x, x + y, x - y, 2 * x, 3 * x, 4 * x, 8 * x, x][op] % P_SECP
        assert g % P_SECP == want and g < (1 << 256), (op, hex(x), hex(y), hex(g))
        if op == 8:
            assert g < P_SECP


@pytest.mark.parametrize("modulus", [0, 1])
def test_division_step_inverse(hostcheck, modulus):
    """p2e_modinv.cuh (what every inverse of the product path runs) against pow(x, -1, m) and against the ladder the
    reference runs: 0 -> 0, the inverse is unique."""
    M = S.MODULI[modulus]
    rng = random.Random(900 + modulus)
    vals = [0, 1, 2, 3, M - 1, M - 2, (M - 1) // 2, (M + 1) // 2, 1 << 255, (1 << 255) - 1, 1 << 30, (1 << 30) - 1, 1 << 60,
            (1 << 240) + 1, M >> 1, M >> 30] + [1 << k for k in range(0, 256, 7)] + [M - (1 << k) for k in range(0, 255, 11)]
    vals = [v % M for v in vals] + [rng.randrange(M) for _ in range(4000)] + [rng.randrange(1 << rng.randrange(1, 256)) for _ in range(500)]
    a = np.array([[(v >> (32 * k)) & 0xFFFFFFFF for k in range(8)] for v in vals], dtype=np.uint32)
    out = np.zeros_like(a)
    hostcheck.hc_modinv(modulus, 0, ctypes.c_size_t(len(vals)), P(a), P(out))
    got = [sum(int(w) << (32 * k) for k, w in enumerate(row)) for row in out]
    for v, g in zip(vals, got):
        assert g == (pow(v, -1, M) if v else 0), (hex(v), hex(g))
    m = 300
    lad = np.zeros((m, 8), dtype=np.uint32)
    hostcheck.hc_modinv(modulus, 1, ctypes.c_size_t(m), P(np.ascontiguousarray(a[:m])), P(lad))
    assert np.array_equal(lad, out[:m])


def test_f8_limb_conversions(hostcheck):
    rng = random.Random(77)
    vals = F8_EDGES + [rng.randrange(1 << 256) for _ in range(500)]
    limbs = np.ascontiguousarray(W.ints_to_limbs(vals).T.astype(np.uint32))
    out = np.zeros_like(limbs)
    hostcheck.hc_f8(9, ctypes.c_size_t(len(vals)), P(limbs), P(limbs), P(out))
    assert _from_words(out) == vals
    back = np.zeros_like(limbs)
    hostcheck.hc_f8(10, ctypes.c_size_t(len(vals)), P(out), P(out), P(back))
    assert np.array_equal(back, limbs)


def test_mulw_fast_core_matches_the_generic_core(hostcheck):
    """K4's fast core (factors < 2^256, q < 2^256) against the oracle: r, q, check_sum columns, carries."""
    rng = random.Random(91)
    vals = [0, 1, 2, P_SECP - 1, P_SECP - 2, P_SECP, (P_SECP - 1) // 2, 1 << 255, (1 << 32) + 977, (1 << 128) - 1]
    pairs = [(a, b) for a in vals for b in vals] + [(rng.randrange(P_SECP), rng.randrange(P_SECP)) for _ in range(3000)]
    x = W.ints_to_limbs([p[0] for p in pairs])
    y = W.ints_to_limbs([p[1] for p in pairs])
    n = len(pairs)
    out = np.zeros((n, 51), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_mulw_fast(ctypes.c_size_t(n), P(aos(x)), P(aos(y)), P(out), P(err))
    eo, ee = U.expect_mul(x, y, 0)
    assert not ee.any() and not err.any()
    assert np.array_equal(out.T, eo)
