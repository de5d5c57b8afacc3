"""GPU tier: the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs
(bit-exact), plus size-independent properties at BASELINE.json's full sizes."""
import random

import numpy as np
import pytest

from oracle import spec as S
from plonky2_ecdsa_b200 import workload as W
from tests import util as U

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- A8 + A9
@pytest.mark.parametrize("modulus", [0, 1])
@pytest.mark.parametrize("n", [1, 127, 129, 2500])
def test_mul_nonnative_witness_vs_oracle(engine, modulus, n):
    x, y = W.mulmod_workload(n, modulus, seed=W.SEED + n)
    out, err = engine.mul_nonnative(modulus, x, y)
    eo, ee = U.expect_mul(x, y, modulus)
    assert np.array_equal(out, eo)
    assert np.array_equal(err, ee)


def test_mul_nonnative_wire_range_errors(engine):
    x = np.zeros((9, 4), dtype=np.uint64)
    y = np.ones((9, 4), dtype=np.uint64)
    x[0, 0] = 1 << 29            # assert limb < 2^29
    x[3, 1] = 1 << 32            # "Value doesn't fit in u32"
    x[8, 2] = S.GL_P - 1
    out, err = engine.mul_nonnative(0, x, y)
    assert list(err) == [2, 1, 1, 0]
    assert not out[:, :3].any()
    assert not out[:35, 3].any() and (out[35:, 3] == np.uint64(1 << 33)).all()     # 0 * 1: carries all zero


def test_empty_batch(engine):
    z = np.zeros((9, 0), dtype=np.uint64)
    out, err = engine.mul_nonnative(0, z, z)
    assert out.shape == (51, 0) and err.shape == (0,)
    assert engine.eval_check_sum_gate(np.zeros((33, 0), dtype=np.uint64)).shape == (17, 0)


def test_mul_full_size_round_trip(engine):
    """Config 2 at full size (2^20): witness -> both gates' constraints must vanish on every valid
    row (mul_nonnative.rs:72-80, 364-374), and a seeded sample matches the oracle bit for bit."""
    import torch
    n = 1 << 20
    x, y = W.mulmod_workload(n, 0)
    out, err = engine.mul_nonnative(0, x, y)
    ok = err == 0
    assert ok.sum() > n // 2
    # rows flagged with the carry-range error are exactly the reference's panics: all of them
    # have an operand above 2^256-ish (unreduced) -- never a reduced/reduced pair
    red = n - n // 4
    assert not (err[36:red] != 0).any()
    wires = np.concatenate([x, y, out[:35]], axis=0)                 # x y r q check_sum
    c = engine.eval_mul_nonnative_gate(0, wires)
    assert not c.any()                                               # holds on every row (q dropped limb included)
    cw = np.concatenate([out[18:35], out[35:51]], axis=0)            # a b
    c2 = engine.eval_check_sum_gate(cw)
    assert not c2[:, ok].any()
    rng = random.Random(1)
    idx = sorted(rng.sample(range(n), 1500)) + list(range(36))
    eo, ee = U.expect_mul(x[:, idx], y[:, idx], 0)
    assert np.array_equal(out[:, idx], eo) and np.array_equal(err[idx], ee)
    del torch


def test_mul_dev_api_with_stride(engine):
    """Device-pointer entry with stride > n on torch memory (plumbing only)."""
    import torch
    n, stride = 1000, 1024
    x, y = W.mulmod_workload(n, 1, seed=5)
    xd = torch.zeros((9, stride), dtype=torch.int64, device="cuda")
    yd = torch.zeros((9, stride), dtype=torch.int64, device="cuda")
    xd[:, :n] = torch.from_numpy(x.view(np.int64)).cuda()
    yd[:, :n] = torch.from_numpy(y.view(np.int64)).cuda()
    od = torch.full((51, stride), -1, dtype=torch.int64, device="cuda")
    ed = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = engine.lib.p2e_mul_nonnative_witness_dev(1, n, xd.data_ptr(), yd.data_ptr(), stride, od.data_ptr(), stride,
                                                  ed.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    out = od.cpu().numpy().view(np.uint64)
    eo, ee = U.expect_mul(x, y, 1)
    assert np.array_equal(out[:, :n], eo)
    assert (out[:, n:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()      # padding untouched
    assert np.array_equal(ed.cpu().numpy().view(np.uint32), ee)


def test_check_sum_witness(engine):
    x, y = W.mulmod_workload(700, 0, seed=77)
    eo, _ = U.expect_mul(x, y, 0)
    a = np.ascontiguousarray(eo[18:35])
    a[:, :50] = W.random_goldilocks(5, 17, 50)
    b, err = engine.check_sum(a)
    for i in range(a.shape[1]):
        eb, panic = S.check_sum_gen_nopanic(U.col(a, i))
        assert U.col(b, i) == eb and bool(err[i] & 16) == panic


# ---------------------------------------------------------------- A4 / A5 / A6 / A7 / A10 / A3
def _edge_pairs(M, rng, k):
    vals = [0, 1, M - 1, M, (1 << 256) - 1, M + 1, 2, M // 2, M // 2 + 1]
    pairs = [(a, b) for a in vals for b in vals] + [(M - 5, 5), (5, M - 5), ((1 << 256) + 3, 5), (7, (1 << 261) - 1)]
    pairs += [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(k)]
    return pairs


@pytest.mark.parametrize("modulus", [0, 1])
def test_add_sub_witness(engine, modulus):
    M = S.MODULI[modulus]
    pairs = _edge_pairs(M, random.Random(40 + modulus), 3000)
    a = W.ints_to_limbs([p[0] for p in pairs])
    b = W.ints_to_limbs([p[1] for p in pairs])
    for is_sub, fn in ((0, engine.add_nonnative), (1, engine.sub_nonnative)):
        out, err = fn(modulus, a, b)
        eo, ee = U.expect_addsub(a, b, modulus, is_sub)
        assert np.array_equal(out, eo) and np.array_equal(err, ee)


def test_add_sub_ragged_limb_counts(engine):
    """Operands with fewer than 9 limbs (constants: zero has 0 limbs, 7 has 1; glv k1 has 5)."""
    rng = random.Random(41)
    n = 300
    a5 = W.ints_to_limbs([rng.getrandbits(140) for _ in range(n)])[:5]
    b1 = W.ints_to_limbs([7] * n)[:1]
    z0 = np.zeros((0, n), dtype=np.uint64)
    out, err = engine.add_nonnative(1, a5, b1)
    eo, ee = U.expect_addsub(a5, b1, 1, 0)
    assert np.array_equal(out, eo) and not err.any()
    out, err = engine.sub_nonnative(1, z0, a5)                 # neg_nonnative: 0 - x
    eo, ee = U.expect_addsub(z0, a5, 1, 1)
    assert np.array_equal(out, eo) and not err.any()


@pytest.mark.parametrize("modulus", [0, 1])
def test_add_many_witness(engine, modulus):
    M = S.MODULI[modulus]
    rng = random.Random(50 + modulus)
    n = 1200
    s = np.stack([W.ints_to_limbs([rng.randrange(1 << 256) for _ in range(n)]) for _ in range(4)])
    s[3] = 0                                                    # curve_double: [xx, xx, xx, A = 0]
    s[:, :, 0] = W.ints_to_limbs([M - 1] * 4).T
    out, err = engine.add_many_nonnative(modulus, s)
    eo, ee = U.expect_add_many(s, modulus)
    assert np.array_equal(out, eo) and np.array_equal(err, ee)
    assert int(out[9, 0]) == 3


@pytest.mark.parametrize("modulus", [0, 1])
@pytest.mark.parametrize("n", [1, 9, 1031])
def test_inverse_witness(engine, modulus, n):
    M = S.MODULI[modulus]
    rng = random.Random(60 + modulus + n)
    xs = ([1, 2, M - 1, M + 1, (1 << 256) - 1, 0, M, 1 << 256] + [rng.randrange(1 << 256) for _ in range(n)])[:max(n, 1)]
    if n == 1:
        xs = [rng.randrange(1, M)]
    x = W.ints_to_limbs(xs)
    out, err = engine.inv_nonnative(modulus, x)
    eo, ee = U.expect_inv(x, modulus)
    assert np.array_equal(err, ee)
    assert np.array_equal(out, eo)


def test_glv_decompose_witness(engine):
    rng = random.Random(70)
    n_ = S.SECP_N
    ks = [0, 1, n_ - 1, n_, n_ // 2, n_ // 2 + 1, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(4000)]
    k = W.ints_to_limbs(ks)
    out, err = engine.decompose_secp256k1_scalar(k)
    eo, ee = U.expect_glv(k)
    assert np.array_equal(out, eo) and np.array_equal(err, ee)
    # k1 + S*k2 == k  (curve/glv.rs:60)
    for i in range(0, len(ks), 97):
        k1 = S.get_biguint(U.col(out[0:5], i))
        k2 = S.get_biguint(U.col(out[5:10], i))
        assert ((-k1 if out[10, i] else k1) + S.GLV_S * (-k2 if out[11, i] else k2)) % n_ == ks[i] % n_


@pytest.mark.parametrize("al,bl", [(9, 9), (18, 9), (9, 1), (5, 9), (17, 2), (3, 5), (1, 1), (0, 1), (18, 18)])
def test_biguint_div_rem_witness(engine, al, bl):
    rng = random.Random(80 + al * 31 + bl)
    n = 400
    a = np.zeros((al, n), dtype=np.uint64)
    b = np.zeros((bl, n), dtype=np.uint64)
    for i in range(n):
        av = rng.getrandbits(rng.randrange(0, 29 * al + 1)) if al else 0
        bv = rng.getrandbits(rng.randrange(0, 29 * bl + 1))
        if i % 50 == 7:
            bv = 0                                              # division by zero -> ERR_DIV_ZERO
        for l in range(al):
            a[l, i] = (av >> (29 * l)) & S.MASK29
        for l in range(bl):
            b[l, i] = (bv >> (29 * l)) & S.MASK29
    div, rem, err = engine.div_rem_biguint(a, b)
    eo, ee = U.expect_divrem(a, b)
    assert np.array_equal(err, ee)
    assert np.array_equal(np.concatenate([div, rem], axis=0), eo)


# ---------------------------------------------------------------- A11 / A12 constraint evaluation
@pytest.mark.parametrize("modulus", [0, 1])
@pytest.mark.parametrize("n", [1, 130, 3000])
def test_eval_mul_nonnative_gate_random_wires(engine, modulus, n):
    w = W.random_goldilocks(900 + n + modulus, 53, n)
    if n >= 130:
        w[:, 0] = S.GL_P - 1                                     # all-max row
        w[:, 1] = 0
        w[:, 2] = np.uint64(0xFFFFFFFF)
        w[:, 3] = np.uint64(0xFFFFFFFF00000000)
    out = engine.eval_mul_nonnative_gate(modulus, w)
    assert np.array_equal(out, U.expect_eval_mul(w, modulus))


@pytest.mark.parametrize("n", [1, 130, 3000])
def test_eval_check_sum_gate_random_wires(engine, n):
    w = W.random_goldilocks(950 + n, 33, n)
    if n >= 130:
        w[:, 0] = S.GL_P - 1
        w[:, 1] = 0
    out = engine.eval_check_sum_gate(w)
    assert np.array_equal(out, U.expect_eval_check_sum(w))


def test_eval_full_size_linearity(engine):
    """Config 5 at full size (2^20 rows): CheckSumGate is affine in its wires, so
    C(w1) + C(w2) - C(w1 + w2) == C(0) row-wise; MulNonnativeGate is checked on a seeded sample."""
    n = 1 << 20
    p = S.GL_P
    w1 = W.random_goldilocks(1, 33, n)
    w2 = W.random_goldilocks(2, 33, n)

    def gl_add(a, b):
        s = a.astype(object) + b.astype(object)
        return np.array(s % p, dtype=np.uint64)
    idx = np.arange(0, n, 4099)
    c1 = engine.eval_check_sum_gate(w1)
    c2 = engine.eval_check_sum_gate(w2)
    w12 = w1.copy()
    w12[:, idx] = gl_add(w1[:, idx], w2[:, idx])
    c12 = engine.eval_check_sum_gate(w12)
    c0 = engine.eval_check_sum_gate(np.zeros((33, 1), dtype=np.uint64))
    lhs = (c1[:, idx].astype(object) + c2[:, idx].astype(object) - c12[:, idx].astype(object)) % p
    assert (lhs == c0.astype(object)).all()
    wm = W.random_goldilocks(3, 53, n)
    cm = engine.eval_mul_nonnative_gate(0, wm)
    assert np.array_equal(cm[:, idx], U.expect_eval_mul(wm[:, idx], 0))
