"""Selector filtering + alpha-combination around the gate evaluators (include/p2e.h
p2e_eval_filtered_combined).  The wrapped code is plonky2's (out of tree, restated in
oracle/vanishing_spec.py -- parity unpinned against the dependency); the gate evaluators inside are the
pinned in-tree ones (oracle/spec.py) and the plonky2_ux restatement (oracle/ux_spec.py)."""
import random

import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from oracle import spec as S
from oracle import ux_spec as U
from oracle import vanishing_spec as V

GL_P = V.GL_P


# ------------------------------------------------------------------------------------------ CPU tier
def test_filter_vanishes_on_the_other_gates_of_the_group_and_on_unused_rows():
    group = (3, 7)
    for me in range(*group):
        for other in range(*group):
            f = V.compute_filter(me, group, other, True)
            assert (f == 0) == (other != me)
        assert V.compute_filter(me, group, V.UNUSED_SELECTOR, True) == 0
        assert V.compute_filter(me, group, V.UNUSED_SELECTOR, False) != 0
    # a single-gate group with one selector polynomial: the empty product
    assert V.compute_filter(0, (0, 1), 0, False) == 1


def test_reduce_with_powers_multi_is_a_power_sum():
    rng = random.Random(1)
    terms = [rng.randrange(GL_P) for _ in range(23)]
    alphas = [rng.randrange(GL_P) for _ in range(2)]
    got = V.reduce_with_powers_multi(terms, alphas)
    for a, g in zip(alphas, got):
        assert g == sum(t * pow(a, k, GL_P) for k, t in enumerate(terms)) % GL_P
    # the gates of a circuit are summed position by position before the reduction: linear, so per-gate
    # reductions add up
    t2 = [rng.randrange(GL_P) for _ in range(17)]
    both = V.evaluate_gate_constraints([terms, t2])
    assert V.reduce_with_powers_multi(both, alphas) == [
        (x + y) % GL_P for x, y in zip(V.reduce_with_powers_multi(terms, alphas), V.reduce_with_powers_multi(t2, alphas))]


def test_fold_of_the_in_tree_gates_is_a_product_of_evaluations():
    """The identities the fused kernels compute instead of evaluate-then-fold, against the gate
    evaluators of the oracle on arbitrary field elements, both moduli:
      sum_c a^c MulNonnative_c = M(a) Q(a) - X(a) Y(a) + R(a) - CS(a)      (gates/mul_nonnative.rs:101-130)
      sum_c a^c CheckSum_c     = A(a) + (a - 2^29) (B(a) - 2^33 sum_{i<16} a^i)   (gates/mul_nonnative.rs:411-427)"""
    rng = random.Random(11)
    ev = lambda v, a: sum(x * pow(a, j, GL_P) for j, x in enumerate(v)) % GL_P
    for m in (S.SECP_P, S.SECP_N):
        m29 = S.convert_base(S.to_u32_digits(m), 32, 29)
        for _ in range(20):
            w = [rng.randrange(GL_P) for _ in range(53)]
            a = rng.choice([0, 1, GL_P - 1, rng.randrange(GL_P)])
            lhs = V.reduce_with_powers_multi(S.mul_nonnative_gate_eval(w, m), [a])[0]
            rhs = (ev(m29, a) * ev(w[27:36], a) - ev(w[0:9], a) * ev(w[9:18], a) + ev(w[18:27], a) - ev(w[36:53], a)) % GL_P
            assert lhs == rhs
    for _ in range(20):
        w = [rng.randrange(GL_P) for _ in range(33)]
        a = rng.choice([0, 1, 1 << 29, rng.randrange(GL_P)])
        lhs = V.reduce_with_powers_multi(S.check_sum_gate_eval(w), [a])[0]
        g16 = sum(pow(a, i, GL_P) for i in range(16))
        rhs = (ev(w[:17], a) + (a - (1 << 29)) * (ev(w[17:33], a) - (1 << 33) * g16)) % GL_P
        assert lhs == rhs


def test_gate_shape_cabi():
    import ctypes
    lib = pkg.load_library()
    nw, nc = ctypes.c_int(), ctypes.c_int()
    prm = (ctypes.c_int * 5)(pkg.SECP256K1_BASE, 0, 0, 0, 0)
    assert lib.p2e_gate_shape(pkg.GATE_MUL_NONNATIVE, ctypes.addressof(prm), ctypes.addressof(nw), ctypes.addressof(nc)) == 0
    assert (nw.value, nc.value) == (53, 17)                    # gates/mul_nonnative.rs:88-99
    assert lib.p2e_gate_shape(pkg.GATE_CHECK_SUM, ctypes.addressof(prm), ctypes.addressof(nw), ctypes.addressof(nc)) == 0
    assert (nw.value, nc.value) == (33, 17)                    # gates/mul_nonnative.rs:398-409
    prm = (ctypes.c_int * 5)(pkg.UX_RANGE_CHECK, 29, 3, 9, 0)
    assert lib.p2e_gate_shape(pkg.GATE_UX, ctypes.addressof(prm), ctypes.addressof(nw), ctypes.addressof(nc)) == 0
    assert (nw.value, nc.value) == (99, 99)
    assert lib.p2e_eval_filtered_scratch_bytes(pkg.GATE_UX, ctypes.addressof(prm), 1000) == 4 * 99 * 8       # fused: powers only
    prm = (ctypes.c_int * 5)(pkg.UX_COMPARISON, 29, 2, 1, 0)
    assert lib.p2e_eval_filtered_scratch_bytes(pkg.GATE_UX, ctypes.addressof(prm), 1000) == (83 * 1000 + 4 * 83) * 8
    assert lib.p2e_gate_shape(7, ctypes.addressof(prm), None, None) < 0
    # no device: the compute entry fails loudly
    import torch
    if not torch.cuda.is_available():
        assert lib.p2e_eval_filtered_combined(pkg.GATE_CHECK_SUM, ctypes.addressof(prm), 0, 0, 1, 0, 1, None, None, None, 1, 0, 0, None) != 0


# ------------------------------------------------------------------------------------------ GPU tier
def _rand_wires(rng, nw, n):
    w = np.array([[rng.randrange(GL_P) for _ in range(n)] for _ in range(nw)], dtype=np.uint64)
    w[:, 0] = 0
    w[:, 1] = GL_P - 1
    return w


GATES = [
    ("mul_nonnative_p", pkg.GATE_MUL_NONNATIVE, (pkg.SECP256K1_BASE,), lambda w: S.mul_nonnative_gate_eval(w, S.SECP_P)),
    ("mul_nonnative_n", pkg.GATE_MUL_NONNATIVE, (pkg.SECP256K1_SCALAR,), lambda w: S.mul_nonnative_gate_eval(w, S.SECP_N)),
    ("check_sum", pkg.GATE_CHECK_SUM, (), S.check_sum_gate_eval),
    ("ux_range_check", pkg.GATE_UX, (pkg.UX_RANGE_CHECK, 29, 3, 9, 0), U.UXRangeCheckGate(29, 3, 9).eval_unfiltered),
    ("ux_range_check_34", pkg.GATE_UX, (pkg.UX_RANGE_CHECK, 34, 3, 8, 0), U.UXRangeCheckGate(34, 3, 8).eval_unfiltered),
    ("ux_range_check_2bit", pkg.GATE_UX, (pkg.UX_RANGE_CHECK, 32, 2, 8, 0), U.UXRangeCheckGate(32, 2, 8).eval_unfiltered),
    ("ux_range_check_1bit", pkg.GATE_UX, (pkg.UX_RANGE_CHECK, 7, 1, 3, 0), U.UXRangeCheckGate(7, 1, 3).eval_unfiltered),
    ("ux_range_check_4bit", pkg.GATE_UX, (pkg.UX_RANGE_CHECK, 29, 4, 5, 0), U.UXRangeCheckGate(29, 4, 5).eval_unfiltered),
    ("ux_comparison", pkg.GATE_UX, (pkg.UX_COMPARISON, 29, 2, 1, 0), U.UXComparisonGate(29, chunk_bits=2).eval_unfiltered),
    ("ux_add_many", pkg.GATE_UX, (pkg.UX_ADD_MANY, 29, 2, 5, 3), U.UXAddManyGate(29, 2, 3, 5).eval_unfiltered),
    ("ux_subtraction", pkg.GATE_UX, (pkg.UX_SUBTRACTION, 29, 2, 6, 0), U.UXSubtractionGate(29, 2, 6).eval_unfiltered),
    ("ux_arithmetic", pkg.GATE_UX, (pkg.UX_ARITHMETIC, 29, 2, 3, 0), U.UXArithmeticGate(29, 2, 3).eval_unfiltered),
]


@pytest.mark.gpu
@pytest.mark.parametrize("g", GATES, ids=[g[0] for g in GATES])
def test_filtered_combined_vs_spec(engine, g):
    _, gate, params, ref_eval = g
    rng = random.Random(77 + len(params))
    nw, nc = engine.gate_shape(gate, params)
    n = 45                                                      # ragged
    w = _rand_wires(rng, nw, n)
    group, me = (2, 6), 4
    # selector constants: this gate's rows, other gates' rows, unused rows, arbitrary field elements (LDE points)
    sel = [me, 2, 5, V.UNUSED_SELECTOR, 0, GL_P - 1] + [rng.randrange(GL_P) for _ in range(n - 6)]
    alphas = [rng.randrange(GL_P) for _ in range(2)]
    first_power = 11
    got = engine.eval_filtered_combined(gate, params, w, alphas, selector=sel, gate_index=me, group_range=group,
                                        many_selectors=True, first_power=first_power)
    assert got.shape == (2, n)
    for j in range(n):
        cons = ref_eval([int(v) for v in w[:, j]])
        assert len(cons) == nc
        exp = V.gate_terms_combined(cons, me, group, sel[j], True, alphas, first_power)
        assert [int(v) for v in got[:, j]] == exp, f"row {j}"
    assert not got[:, [1, 2, 3]].any()                          # filtered out
    # no selector column: plain alpha-combination of the unfiltered constraints, 4 challenges, power 0
    al4 = [rng.randrange(GL_P) for _ in range(4)]
    got4 = engine.eval_filtered_combined(gate, params, w, al4)
    for j in (0, 1, 7, n - 1):
        cons = ref_eval([int(v) for v in w[:, j]])
        assert [int(v) for v in got4[:, j]] == V.reduce_with_powers_multi(cons, al4)
    # 1 and 3 challenges (3 runs through the 4-challenge instantiation), accumulating onto a previous result
    for k in (1, 3):
        prev = np.array([[rng.randrange(GL_P) for _ in range(n)] for _ in range(k)], dtype=np.uint64)
        gotk = engine.eval_filtered_combined(gate, params, w, al4[:k], selector=sel, gate_index=me, group_range=group,
                                             many_selectors=False, first_power=1, accumulate_into=prev)
        assert gotk.shape == (k, n)
        for j in (0, 3, 6, n - 1):
            cons = ref_eval([int(v) for v in w[:, j]])
            add = V.gate_terms_combined(cons, me, group, sel[j], False, al4[:k], 1)
            assert [int(v) for v in gotk[:, j]] == [(int(prev[a, j]) + add[a]) % GL_P for a in range(k)]


@pytest.mark.gpu
def test_gates_of_one_circuit_accumulate(engine):
    """Two gate types sharing the LDE rows: the second call adds to the first (evaluate_gate_constraints
    sums the filtered vectors position by position; reduce_with_powers_multi is linear)."""
    rng = random.Random(5)
    n = 33
    w = _rand_wires(rng, 53, n)
    sel = [rng.choice([0, 1, V.UNUSED_SELECTOR]) for _ in range(n)]
    alphas = [rng.randrange(GL_P) for _ in range(2)]
    acc = engine.eval_filtered_combined(pkg.GATE_MUL_NONNATIVE, (pkg.SECP256K1_BASE,), w, alphas, selector=sel,
                                        gate_index=0, group_range=(0, 2), many_selectors=True, first_power=3)
    acc = engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), w[:33], alphas, selector=sel, gate_index=1,
                                        group_range=(0, 2), many_selectors=True, first_power=3, accumulate_into=acc)
    for j in range(n):
        row = [int(v) for v in w[:, j]]
        a = V.eval_filtered(S.mul_nonnative_gate_eval(row, S.SECP_P), 0, (0, 2), sel[j], True)
        b = V.eval_filtered(S.check_sum_gate_eval(row[:33]), 1, (0, 2), sel[j], True)
        exp = V.reduce_with_powers_multi([0, 0, 0] + V.evaluate_gate_constraints([a, b]), alphas)
        assert [int(v) for v in acc[:, j]] == exp


@pytest.mark.gpu
def test_filtered_combined_vanishes_on_generated_rows_full_size(engine):
    """Config-5 size (2^20 LDE rows), device-resident: rows that hold real MulNonnative witnesses give an
    all-zero combination whatever the selector says; corrupting one wire of one row shows up in exactly
    that row when the selector selects the gate and nowhere when it does not."""
    import ctypes
    import torch
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randint(0, 1 << 29, (9, n), dtype=torch.int64, device="cuda", generator=g)
    y = torch.randint(0, 1 << 29, (9, n), dtype=torch.int64, device="cuda", generator=g)
    rows = torch.empty((18 + pkg.MUL_OUT_ROWS, n), dtype=torch.int64, device="cuda")
    rows[:9], rows[9:18] = x, y
    err = torch.empty(n, dtype=torch.int32, device="cuda")
    lib = engine.lib
    torch.cuda.synchronize()      # torch's stream and the engine's stream are unordered: finish the fills first
    assert lib.p2e_mul_nonnative_witness_dev(pkg.SECP256K1_BASE, n, x.data_ptr(), y.data_ptr(), n, rows[18:].data_ptr(), n,
                                             err.data_ptr(), None) == 0
    sel = torch.zeros(n, dtype=torch.int64, device="cuda")
    sel[1::2] = 1                                                # odd rows belong to the other gate of the group
    alphas = np.array([0x123456789ABCDEF, GL_P - 2], dtype=np.uint64)
    out = torch.full((2, n), 7, dtype=torch.int64, device="cuda")
    prm = (ctypes.c_int * 5)(pkg.SECP256K1_BASE, 0, 0, 0, 0)

    def run():
        torch.cuda.synchronize()
        assert lib.p2e_eval_filtered_combined_dev(pkg.GATE_MUL_NONNATIVE, ctypes.addressof(prm), 0, 0, 2, 1, n, rows.data_ptr(), n,
                                                  sel.data_ptr(), alphas.ctypes.data, 2, 5, 0, out.data_ptr(), n, None, None) == 0
        engine.synchronize()
        torch.cuda.synchronize()

    run()
    assert not bool(out.any())
    rows[20, 1000] += 1          # r[2] of an even (selected) row
    rows[20, 1001] += 1          # and of an odd (filtered-out) row
    run()
    nz = torch.nonzero(out.any(dim=0)).flatten().tolist()
    assert nz == [1000]


@pytest.mark.gpu
def test_filtered_combined_argument_errors(engine):
    w = np.zeros((33, 4), dtype=np.uint64)
    with pytest.raises(pkg.P2EError):
        engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), w, [1, 2, 3, 4, 5])              # > 4 challenges
    with pytest.raises(pkg.P2EError):
        engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), w, [GL_P])                        # non-canonical alpha
    with pytest.raises(pkg.P2EError):
        engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), w, [3], selector=[0, 0, 0, 0], gate_index=5, group_range=(0, 2))
    assert engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), np.zeros((33, 0), dtype=np.uint64), [3]).shape == (1, 0)
