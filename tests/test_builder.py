"""The builder C-API (p2e_builder_*): mirrors of the reference's gadget-level circuit tests, built
through plonky2_ecdsa_b200.CircuitBuilder.  CPU tier: tape layouts vs the oracle's gadget call
order; GPU tier: witnesses vs the oracle."""
import random

import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from oracle import spec as S
from plonky2_ecdsa_b200 import workload as W

P_, N_ = pkg.SECP256K1_BASE, pkg.SECP256K1_SCALAR


def _layout(lib, tape):
    cnt = lib.p2e_tape_layout(tape, None, None)
    kinds = np.empty(cnt, dtype=np.uint32)
    offs = np.empty(cnt, dtype=np.uint32)
    lib.p2e_tape_layout(tape, kinds.ctypes.data, offs.ctypes.data)
    return list(kinds), list(offs), int(lib.p2e_tape_slots(tape)), lib.p2e_tape_inputs(tape)


def _oracle_layout(tape):
    kinds, offs, off = [], [], 0
    for k, vals in tape:
        kinds.append(k)
        offs.append(off)
        off += len(vals)
    return kinds, offs, off


def build_ecdsa(cb):
    """verify_secp256k1_message_circuit (gadgets/ecdsa.rs:30-53) spelled out with the builder."""
    msg, r, s, pkx, pky = (cb.add_virtual_nonnative_target() for _ in range(5))
    pk = cb.affine_point(pkx, pky)
    cb.curve_assert_valid(pk)
    c = cb.inv_nonnative(N_, s)
    u1 = cb.mul_nonnative(N_, msg, c)
    u2 = cb.mul_nonnative(N_, r, c)
    point1 = cb.fixed_base_curve_mul_circuit(u1)
    point2 = cb.glv_mul(pk, u2)
    point = cb.curve_add(point1, point2)
    cb.connect_nonnative(r, cb.point_x(point))
    return cb.build()


def build_nonnative_chain(cb, n_mul=11):
    """test_nonnative_mul (gadgets/nonnative.rs:1007-1038): x * y_0 * ... * y_10, plus add/sub/neg/inv uses."""
    x = cb.add_virtual_nonnative_target()
    ys = [cb.add_virtual_nonnative_target() for _ in range(n_mul)]
    acc = x
    for y in ys:
        acc = cb.mul_nonnative(P_, acc, y)
    s = cb.add_nonnative(P_, acc, x)
    d = cb.sub_nonnative(P_, s, ys[0])
    ng = cb.neg_nonnative(P_, d)
    iv = cb.inv_nonnative(P_, ng)
    one = cb.mul_nonnative(P_, iv, ng)
    cb.connect_nonnative(one, cb.constant_nonnative(1))
    return cb.build()


def oracle_nonnative_chain(vals):
    m = S.SECP_P
    wb = S.WitnessBuilder()
    x = S.NN(S.limbs9(vals[0]))
    ys = [S.NN(S.limbs9(v)) for v in vals[1:]]
    acc = x
    for y in ys:
        acc = wb.mul_nonnative(acc, y, m)
    s = wb.add_nonnative(acc, x, m)
    d = wb.sub_nonnative(s, ys[0], m)
    ng = wb.neg_nonnative(d, m)
    iv = wb.inv_nonnative(ng, m)
    one = wb.mul_nonnative(iv, ng, m)
    assert S.get_biguint(one.limbs) == 1
    return wb


def build_curve_circuit(cb):
    x1, y1, x2, y2, b = (cb.add_virtual_nonnative_target() for _ in range(5))
    p1, p2 = cb.affine_point(x1, y1), cb.affine_point(x2, y2)
    p3 = cb.curve_add(p1, p2)
    p4 = cb.curve_double(p3)
    p5 = cb.curve_conditional_add(p4, p1, b)
    g = cb.constant_affine_point(S.G)
    p6 = cb.curve_add(p5, g)
    ex, ey = cb.add_virtual_nonnative_target(), cb.add_virtual_nonnative_target()
    cb.connect_nonnative(cb.point_x(p6), ex)
    cb.connect_nonnative(cb.point_y(p6), ey)
    return cb.build()


def oracle_curve_circuit(P1, P2, b):
    wb = S.WitnessBuilder()
    p1, p2 = wb.constant_affine_point(P1), wb.constant_affine_point(P2)
    p3 = wb.curve_add(p1, p2)
    p4 = wb.curve_double(p3)
    p5 = wb.curve_conditional_add(p4, p1, b)
    p6 = wb.curve_add(p5, wb.constant_affine_point(S.G))
    return wb, (S.get_biguint(p6[0].limbs), S.get_biguint(p6[1].limbs))


# ------------------------------------------------------------------------------------------ CPU tier
def test_builder_ecdsa_layout_equals_builtin_tape():
    lib = pkg.load_library()
    tid = build_ecdsa(pkg.CircuitBuilder(lib))
    assert tid >= 16
    assert _layout(lib, tid) == _layout(lib, 0)


def test_builder_nonnative_chain_layout_matches_oracle():
    lib = pkg.load_library()
    tid = build_nonnative_chain(pkg.CircuitBuilder(lib))
    rng = random.Random(1)
    wb = oracle_nonnative_chain([rng.randrange(1, S.SECP_P) for _ in range(12)])
    assert _layout(lib, tid) == (*_oracle_layout(wb.tape), 12)


def test_builder_curve_layout_matches_oracle():
    lib = pkg.load_library()
    tid = build_curve_circuit(pkg.CircuitBuilder(lib))
    wb, _ = oracle_curve_circuit(S.pt_mul(3, S.G), S.pt_mul(5, S.G), 1)
    assert _layout(lib, tid) == (*_oracle_layout(wb.tape), 7)


def test_builder_rejects_unsupported_dataflow():
    lib = pkg.load_library()
    cb2 = pkg.CircuitBuilder(lib)
    k = cb2.add_virtual_nonnative_target()
    cb2.fixed_base_curve_mul_circuit(k)
    with pytest.raises(pkg.P2EError):
        cb2.fixed_base_curve_mul_circuit(k)             # one per tape in this version
    with pytest.raises(pkg.P2EError):
        cb2.add_nonnative(7, k, k)                      # unknown modulus


def build_multi_round(cb):
    """Nonnative generators fed by curve-op outputs, and curve ops fed by those (three dependency
    rounds): the pattern of curve_assert_valid(result) / chained gadgets in the reference's tests."""
    x1, y1, x2, y2, k = (cb.add_virtual_nonnative_target() for _ in range(5))
    p, q = cb.affine_point(x1, y1), cb.affine_point(x2, y2)
    d = cb.curve_double(p)                                   # round 0
    s = cb.curve_add(d, q)                                   # round 0
    m = cb.mul_nonnative(P_, cb.point_x(s), x1)              # round 1 generator on a curve-op output
    a = cb.add_nonnative(P_, m, cb.point_y(d))
    cb.curve_assert_valid(s)                                 # round 1
    w = cb.affine_point(cb.point_x(s), cb.point_y(s))        # round 1 point made of round-0 coordinates
    u = cb.curve_double(w)                                   # round 1
    v = cb.curve_add(u, d)                                   # round 1, operand from round 0
    g = cb.glv_mul(v, k)                                     # round 2 (works on v's coordinate values)
    z = cb.sub_nonnative(P_, cb.point_x(g), a)               # round 3
    cb.inv_nonnative(P_, z)
    return cb.build()


def oracle_multi_round(P1, P2, kk):
    m = S.SECP_P
    wb = S.WitnessBuilder()
    p, q = wb.constant_affine_point(P1), wb.constant_affine_point(P2)
    d = wb.curve_double(p)
    s = wb.curve_add(d, q)
    mm = wb.mul_nonnative(s[0], p[0], m)
    a = wb.add_nonnative(mm, d[1], m)
    wb.curve_assert_valid(s)
    u = wb.curve_double(s)
    v = wb.curve_add(u, d)
    g = wb.glv_mul(v, S.NN(S.limbs9(kk)))
    z = wb.sub_nonnative(g[0], a, m)
    wb.inv_nonnative(z, m)
    return wb


def test_builder_multi_round_layout_matches_oracle():
    lib = pkg.load_library()
    tid = build_multi_round(pkg.CircuitBuilder(lib))
    wb = oracle_multi_round(S.pt_mul(3, S.G), S.pt_mul(5, S.G), 0x1234567)
    assert _layout(lib, tid) == (*_oracle_layout(wb.tape), 5)


# ------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
def test_builder_nonnative_chain_witness(engine):
    tid = build_nonnative_chain(pkg.CircuitBuilder(engine.lib))
    rng = random.Random(2)
    n = 50
    cols = [[rng.randrange(1, S.SECP_P) for _ in range(n)] for _ in range(12)]
    wit, err = engine.tape_witness(tid, [W.ints_to_limbs(c) for c in cols])
    assert not err.any()
    for j in (0, 17, 49):
        wb = oracle_nonnative_chain([c[j] for c in cols])
        assert [int(v) for v in wit[:, j]] == wb.flat()


@pytest.mark.gpu
def test_builder_curve_circuit_witness(engine):
    tid = build_curve_circuit(pkg.CircuitBuilder(engine.lib))
    rng = random.Random(3)
    n = 24
    P1 = [S.pt_mul(rng.randrange(1, S.SECP_N), S.G) for _ in range(n)]
    P2 = [S.pt_mul(rng.randrange(1, S.SECP_N), S.G) for _ in range(n)]
    bs = [j % 2 for j in range(n)]
    refs = [oracle_curve_circuit(P1[j], P2[j], bs[j]) for j in range(n)]
    exp = [r[1] for r in refs]
    exp_bad = list(exp)
    exp_bad[5] = (exp[5][0] ^ 1, exp[5][1])                 # wrong expected x for instance 5 -> ERR_CONNECT
    bm = np.zeros((9, n), dtype=np.uint64)
    bm[0] = bs
    ins = [W.ints_to_limbs([p[0] for p in P1]), W.ints_to_limbs([p[1] for p in P1]),
           W.ints_to_limbs([p[0] for p in P2]), W.ints_to_limbs([p[1] for p in P2]), bm,
           W.ints_to_limbs([e[0] for e in exp_bad]), W.ints_to_limbs([e[1] for e in exp_bad])]
    wit, err = engine.tape_witness(tid, ins)
    assert err[5] & 128 and not np.delete(err, 5).any()
    for j in range(n):
        assert [int(v) for v in wit[:, j]] == refs[j][0].flat()


@pytest.mark.gpu
def test_builder_ecdsa_witness_equals_builtin(engine):
    tid = build_ecdsa(pkg.CircuitBuilder(engine.lib))
    wl = W.ecdsa_workload(9, seed=123)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    w1, e1 = engine.verify_secp256k1_message_witness(*ins)
    w2, e2 = engine.tape_witness(tid, ins)
    assert np.array_equal(w1, w2) and np.array_equal(e1, e2) and not e1.any()


def build_misc_nonnative(cb, modulus):
    """test_nonnative_many_adds / conditional neg / if (gadgets/nonnative.rs:943-977, 584-596, 297-308)."""
    a, b, c, d, flag = (cb.add_virtual_nonnative_target() for _ in range(5))
    s = cb.add_many_nonnative(modulus, [a, b, c, d])
    t = cb.nonnative_conditional_neg(modulus, s, flag)
    u = cb.if_nonnative(modulus, flag, a, b)
    v = cb.mul_nonnative(modulus, t, u)
    one = cb.add_many_nonnative(modulus, [v])          # single summand: returns its input, no generator
    assert one == v
    return cb.build()


def oracle_misc_nonnative(vals, flag, m):
    wb = S.WitnessBuilder()
    a, b, c, d = (S.NN(S.limbs9(v)) for v in vals)
    s = wb.add_many_nonnative([a, b, c, d], m)
    t = wb.nonnative_conditional_neg(s, flag, m)
    u = wb.if_nonnative(flag, a, b, m)
    wb.mul_nonnative(t, u, m)
    return wb


@pytest.mark.parametrize("modulus", [P_, N_])
def test_builder_misc_nonnative_layout(modulus):
    lib = pkg.load_library()
    tid = build_misc_nonnative(pkg.CircuitBuilder(lib), modulus)
    wb = oracle_misc_nonnative([5, 6, 7, 8], 1, S.MODULI[modulus])
    assert _layout(lib, tid) == (*_oracle_layout(wb.tape), 5)


@pytest.mark.gpu
@pytest.mark.parametrize("modulus", [P_, N_])
def test_builder_misc_nonnative_witness(engine, modulus):
    m = S.MODULI[modulus]
    tid = build_misc_nonnative(pkg.CircuitBuilder(engine.lib), modulus)
    rng = random.Random(4 + modulus)
    n = 40
    cols = [[rng.randrange(m) for _ in range(n)] for _ in range(4)]
    cols[0][0] = cols[1][0] = cols[2][0] = cols[3][0] = m - 1          # overflow count 3
    flags = [j % 2 for j in range(n)]
    fm = np.zeros((9, n), dtype=np.uint64)
    fm[0] = flags
    wit, err = engine.tape_witness(tid, [W.ints_to_limbs(c) for c in cols] + [fm])
    assert not err.any()
    for j in range(n):
        wb = oracle_misc_nonnative([c[j] for c in cols], flags[j], m)
        assert [int(v) for v in wit[:, j]] == wb.flat()


@pytest.mark.gpu
def test_builder_multi_round_witness(engine):
    tid = build_multi_round(pkg.CircuitBuilder(engine.lib))
    rng = random.Random(11)
    n = 20
    P1 = [S.pt_mul(rng.randrange(1, S.SECP_N), S.G) for _ in range(n)]
    P2 = [S.pt_mul(rng.randrange(1, S.SECP_N), S.G) for _ in range(n)]
    ks = [rng.randrange(1, S.SECP_N) for _ in range(n)]
    ins = [W.ints_to_limbs([p[0] for p in P1]), W.ints_to_limbs([p[1] for p in P1]),
           W.ints_to_limbs([p[0] for p in P2]), W.ints_to_limbs([p[1] for p in P2]), W.ints_to_limbs(ks)]
    wit, err = engine.tape_witness(tid, ins)
    assert not err.any()
    for j in range(n):
        wb = oracle_multi_round(P1[j], P2[j], ks[j])
        assert [int(v) for v in wit[:, j]] == wb.flat()
