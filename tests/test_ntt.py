"""Goldilocks NTT / coset LDE: the oracle against the defining DFT and the field
constants' identities (CPU tier); the CUDA path against the oracle and, at full size, against direct
polynomial evaluation and the round trip (GPU tier)."""
import random

import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from oracle import ntt_spec as N

P = N.GL_P


# ------------------------------------------------------------------------------------------ CPU tier
def test_field_constants_identities():
    # 7 generates the whole multiplicative group: 7^((p-1)/q) != 1 for every prime q | p - 1 = 2^32 * 3 * 5 * 17 * 257 * 65537
    assert (1 << 32) * 3 * 5 * 17 * 257 * 65537 == P - 1
    for q in (2, 3, 5, 17, 257, 65537):
        assert pow(N.MULTIPLICATIVE_GROUP_GENERATOR, (P - 1) // q, P) != 1
    assert N.POWER_OF_TWO_GENERATOR == pow(N.MULTIPLICATIVE_GROUP_GENERATOR, (P - 1) >> 32, P)
    assert pow(N.POWER_OF_TWO_GENERATOR, 1 << 31, P) == P - 1
    w = N.primitive_root_of_unity(10)
    assert pow(w, 1024, P) == 1 and pow(w, 512, P) == P - 1


def test_oracle_fft_is_the_dft():
    rng = random.Random(1)
    for log_n in range(0, 7):
        x = [rng.randrange(P) for _ in range(1 << log_n)]
        assert N.fft(x) == N.dft_naive(x, N.primitive_root_of_unity(log_n))
        assert N.ifft(N.fft(x)) == x


def test_oracle_lde_evaluates_the_interpolant_on_the_coset():
    rng = random.Random(2)
    n, rb = 16, 3
    coeffs = [rng.randrange(P) for _ in range(n)]
    values = N.fft(coeffs)
    lde = N.lde_onto_coset(values, rb)
    w = N.primitive_root_of_unity(7)
    for j in (0, 1, 5, 77, 127):
        assert lde[j] == N.eval_poly(coeffs, 7 * pow(w, j, P) % P)


def test_ntt_cabi_rejects_bad_sizes():
    lib = pkg.load_library()
    assert lib.p2e_ntt(1, 0, 1, None, None) < 0 and lib.p2e_ntt(25, 0, 1, None, None) < 0
    assert lib.p2e_coset_lde(10, 0, 1, None, None) < 0 and lib.p2e_coset_lde(20, 5, 1, None, None) < 0
    assert lib.p2e_ntt_scratch_bytes(17, 3, 2) == 2 * (2 * (1 << 17) + (1 << 20)) * 8


# ------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14])
def test_fft_and_ifft_vs_oracle(engine, log_n):
    rng = random.Random(100 + log_n)
    n, polys = 1 << log_n, 3
    x = [[rng.randrange(P) for _ in range(n)] for _ in range(polys)]
    x[1][0] = P - 1
    x[2] = [0] * n
    x[2][1] = 1                                              # delta at 1 -> the row of powers of omega
    a = np.array(x, dtype=np.uint64)
    f = engine.fft(a)
    for p in range(polys):
        assert [int(v) for v in f[p]] == N.fft(x[p])
    w = N.primitive_root_of_unity(log_n)
    assert int(f[2][3]) == pow(w, 3, P)
    g = engine.fft(f, inverse=True)
    assert np.array_equal(g, a)


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,rate_bits", [(2, 1), (4, 3), (7, 3), (10, 1), (10, 3), (9, 4)])
def test_lde_vs_oracle(engine, log_n, rate_bits):
    rng = random.Random(200 + 10 * log_n + rate_bits)
    n, polys = 1 << log_n, 2
    x = [[rng.randrange(P) for _ in range(n)] for _ in range(polys)]
    out = engine.lde_onto_coset(np.array(x, dtype=np.uint64), rate_bits)
    assert out.shape == (polys, n << rate_bits)
    for p in range(polys):
        assert [int(v) for v in out[p]] == N.lde_onto_coset(x[p], rate_bits)


@pytest.mark.gpu
def test_full_size_lde_properties(engine):
    """BASELINE config 5's shape: trace 2^17, rate 8 -> 2^20 LDE rows.  Size-independent checks: the LDE
    values equal the interpolant evaluated directly at sampled coset points; constants extend to
    constants; linearity; ifft(fft(x)) == x at 2^20."""
    rng = np.random.default_rng(7)
    log_n, rb = 17, 3
    n = 1 << log_n
    vals = rng.integers(0, P, size=(4, n), dtype=np.uint64)
    vals[2] = 12345                                          # a constant polynomial
    vals[3] = (vals[0].astype(object) + vals[1].astype(object)) % P      # row 3 = row 0 + row 1
    lde = engine.lde_onto_coset(vals, rb)
    assert (lde[2] == 12345).all()
    assert np.array_equal(lde[3].astype(object), (lde[0].astype(object) + lde[1].astype(object)) % P)
    coeffs = [int(v) for v in engine.fft(vals[:1], inverse=True)[0]]
    w = N.primitive_root_of_unity(log_n + rb)
    for j in (0, 1, 524287, 777777, (1 << 20) - 1):
        assert int(lde[0][j]) == N.eval_poly(coeffs, 7 * pow(w, j, P) % P)
    # the original values sit on the subgroup: P(omega_n^i) from the coefficients
    wn = N.primitive_root_of_unity(log_n)
    for i in (0, 3, n - 1):
        assert int(vals[0][i]) == N.eval_poly(coeffs, pow(wn, i, P))
    big = rng.integers(0, P, size=(2, 1 << 20), dtype=np.uint64)
    assert np.array_equal(engine.fft(engine.fft(big), inverse=True), big)


@pytest.mark.gpu
def test_lde_feeds_the_gate_evaluator(engine):
    """Valid MulNonnative rows on the trace domain -> LDE -> every constraint polynomial vanishes ON the
    subgroup; on the coset the evaluator applied to the LDE equals the LDE of ... nothing to compare row-wise,
    but degree-2 constraints of the interpolants agree with evaluating them from the coefficient form."""
    from plonky2_ecdsa_b200 import workload as W
    log_n, rb = 8, 3
    n = 1 << log_n
    x, y = W.mulmod_workload(n, 0, seed=77)
    x[:, :16] = 5                                            # make sure the error-free rows dominate
    y[:, :16] = 9
    out, err = engine.mul_nonnative(0, x, y)
    ok = err == 0
    rows = np.concatenate([x, y, out[:35]], axis=0)          # 53 wires x n trace rows
    rows[:, ~ok] = 0                                         # all-zero rows satisfy the gate too
    assert not engine.eval_mul_nonnative_gate(0, rows).any()
    lde = engine.lde_onto_coset(rows, rb)                    # [53][8 n] wire-major LDE matrix
    c = engine.eval_mul_nonnative_gate(0, lde)               # constraints on the coset: degree < 2 n polynomials
    # each constraint polynomial vanishes on the trace subgroup, so it is Z_H(x) * t(x) with deg t < n:
    # c[j] / (x_j^n - 1) must itself be a polynomial of degree < n on the coset.  Check through the ifft:
    g, w = 7, N.primitive_root_of_unity(log_n + rb)
    zh_inv = [pow((pow(g * pow(w, j, P) % P, n, P) - 1) % P, P - 2, P) for j in range(n << rb)]
    t_vals = np.array([[int(c[k][j]) * zh_inv[j] % P for j in range(n << rb)] for k in (0, 8, 16)], dtype=np.uint64)
    # undo the coset: values of t on g*H_m -> coefficients of t(g x) -> high coefficients must be zero
    t_coeffs = engine.fft(t_vals, inverse=True)
    assert not t_coeffs[:, n:].any()


@pytest.mark.gpu
def test_lde_after_device_context_is_rebuilt(engine):
    """Per-device library state (twiddle tables, the > 48 KB dynamic shared memory opt-in of the NTT kernels, chain arenas)
    must follow the context: release the device's context, create it again (on the LAST visible device when there are
    several -- a device the process has not used yet), run an LDE that needs the opt-in and a chain witness."""
    import torch
    from plonky2_ecdsa_b200 import workload as W
    from tests import util as U
    dev = torch.cuda.device_count() - 1
    vals = [(i * i * 131 + 7) % P for i in range(1 << 13)]          # 2^13 -> 2^16: lines of 2^8, > 48 KB tiles
    want = N.lde_onto_coset(vals, 3)
    if dev == 0:
        assert engine.lib.p2e_shutdown_device(0) == 0
    e2 = pkg.Engine(dev)
    try:
        lde = e2.lde_onto_coset(np.array([vals], dtype=np.uint64), 3)
        assert [int(v) for v in lde[0]] == want
        x, y = W.mulmod_workload(64, 0, seed=5)
        out, err = e2.mul_nonnative(0, x, y)
        eo, ee = U.expect_mul(x, y, 0)
        assert np.array_equal(out, eo) and np.array_equal(err, ee)
    finally:
        if dev != 0:
            e2.close()
        engine.select()
