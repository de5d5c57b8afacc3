"""Shared helpers: oracle expectations for wire-major batches (test infrastructure)."""
import numpy as np

from oracle import spec as S

ERR_NONCANONICAL, ERR_INVERSE_ZERO, ERR_CARRY_RANGE, ERR_DIV_ZERO, ERR_LIMBS_OVERFLOW = 4, 8, 16, 32, 64


def col(m, i):
    return [int(v) for v in m[:, i]]


def expect_mul(x, y, modulus):
    """x, y: [9][n] -> (out [51][n], err [n]) per the oracle (A8 + A9, no-panic carry variant)."""
    M = S.MODULI[modulus]
    n = x.shape[1]
    out = np.zeros((51, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        xs, ys = col(x, i), col(y, i)
        try:
            q, r, cs = S.mul_nonnative_gen(xs, ys, M)
        except S.RefPanic as e:
            err[i] = 1 if "u32" in str(e) else 2
            continue
        b, panic = S.check_sum_gen_nopanic(cs)
        out[:, i] = r + q + cs + b
        err[i] = ERR_CARRY_RANGE if panic else 0
    return out, err


def expect_addsub(a, b, modulus, is_sub):
    M = S.MODULI[modulus]
    n = a.shape[1] if a.shape[0] else b.shape[1]
    out = np.zeros((10, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    fn = S.nonnative_sub_gen if is_sub else S.nonnative_add_gen
    for i in range(n):
        try:
            r, ov = fn(col(a, i), col(b, i), M)
            out[:, i] = r + [ov]
        except S.RefPanic:
            err[i] = ERR_NONCANONICAL
    return out, err


def expect_add_many(s, modulus):
    M = S.MODULI[modulus]
    k, _, n = s.shape
    out = np.zeros((10, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        try:
            r, ov = S.nonnative_add_many_gen([[int(v) for v in s[t, :, i]] for t in range(k)], M)
            out[:, i] = r + [ov]
        except S.RefPanic:
            err[i] = ERR_NONCANONICAL
    return out, err


def expect_inv(x, modulus):
    M = S.MODULI[modulus]
    n = x.shape[1]
    out = np.zeros((18, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        try:
            d, iv = S.nonnative_inv_gen(col(x, i), M)
            out[:, i] = d + iv
        except S.RefPanic as e:
            err[i] = ERR_INVERSE_ZERO if "zero" in str(e) else ERR_NONCANONICAL
    return out, err


def expect_glv(k):
    n = k.shape[1]
    out = np.zeros((12, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        try:
            k1, k2, n1, n2 = S.glv_decomposition_gen(col(k, i))
            out[:, i] = k1 + k2 + [n1, n2]
        except S.RefPanic:
            err[i] = ERR_NONCANONICAL
    return out, err


def expect_divrem(a, b):
    al, bl = a.shape[0], b.shape[0]
    n = b.shape[1]
    dl = 0 if bl > al + 1 else al - bl + 1
    out = np.zeros((dl + bl, n), dtype=np.uint64)
    err = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        try:
            d, r = S.biguint_divrem_gen(col(a, i), col(b, i))
            out[:, i] = d + r
        except S.RefPanic as e:
            err[i] = ERR_DIV_ZERO if "zero" in str(e) else ERR_LIMBS_OVERFLOW
    return out, err


def expect_eval_mul(w, modulus):
    M = S.MODULI[modulus]
    n = w.shape[1]
    out = np.zeros((17, n), dtype=np.uint64)
    for i in range(n):
        out[:, i] = S.mul_nonnative_gate_eval(col(w, i), M)
    return out


def expect_eval_check_sum(w):
    n = w.shape[1]
    out = np.zeros((17, n), dtype=np.uint64)
    for i in range(n):
        out[:, i] = S.check_sum_gate_eval(col(w, i))
    return out
