"""plonky2_ux gates (PARITY UNPINNED -- checked against the restated spec
oracle/ux_spec.py, not against upstream).  CPU tier: the spec's own consistency (generated rows
satisfy the gate, corrupted rows do not) and the C-ABI's geometry; GPU tier: generators and
evaluators vs the spec, bit-exact."""
import random

import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from oracle import ux_spec as U

GL_P = U.GL_P

# (gate id, spec object, bits, limb_bits, count, num_addends)
CONFIGS = [
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(29, 3, 9), 29, 3, 9, 0),       # x / y / r / q limbs, nonnative.rs:454-458
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(34, 3, 8), 34, 3, 8, 0),       # carries b[0..8], nonnative.rs:459-460
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(29, 2, 1), 29, 2, 1, 0),       # single overflow limb, nonnative.rs:331
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(32, 2, 8), 32, 2, 8, 0),       # upstream plonky2_u32 geometry
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(29, 4, 5), 29, 4, 5, 0),
    (pkg.UX_RANGE_CHECK, U.UXRangeCheckGate(7, 1, 3), 7, 1, 3, 0),
    (pkg.UX_COMPARISON, U.UXComparisonGate(29, chunk_bits=2), 29, 2, 1, 0),  # list_le_ux_circuit, biguint.rs:228
    (pkg.UX_COMPARISON, U.UXComparisonGate(32, chunk_bits=2), 32, 2, 1, 0),
    (pkg.UX_COMPARISON, U.UXComparisonGate(34, chunk_bits=2), 34, 2, 1, 0),
    (pkg.UX_COMPARISON, U.UXComparisonGate(30, chunk_bits=3), 30, 3, 1, 0),
    (pkg.UX_ADD_MANY, U.UXAddManyGate(29, 2, 3, 5), 29, 2, 5, 3),          # add_biguint, biguint.rs:261
    (pkg.UX_ADD_MANY, U.UXAddManyGate(32, 2, 3, 5), 32, 2, 5, 3),
    (pkg.UX_ADD_MANY, U.UXAddManyGate(29, 3, 16, 1), 29, 3, 1, 16),
    (pkg.UX_SUBTRACTION, U.UXSubtractionGate(29, 2, 6), 29, 2, 6, 0),      # sub_biguint, biguint.rs:284
    (pkg.UX_SUBTRACTION, U.UXSubtractionGate(32, 2, 6), 32, 2, 6, 0),
    (pkg.UX_ARITHMETIC, U.UXArithmeticGate(29, 2, 3), 29, 2, 3, 0),        # mul_biguint, biguint.rs:305
    (pkg.UX_ARITHMETIC, U.UXArithmeticGate(32, 2, 3), 32, 2, 3, 0),
]
IDS = [f"{type(c[1]).__name__}-{c[2]}-{c[3]}-{c[4]}" for c in CONFIGS]


def _inputs(gate, spec, bits, rng, edge=None):
    """One instance's generator inputs (flat, in the C-ABI's input order) and the spec call arguments."""
    top = (1 << bits) - 1

    def val():
        if edge == "zero":
            return 0
        if edge == "max":
            return top
        return rng.randrange(1 << bits)

    if gate == pkg.UX_RANGE_CHECK:
        v = [val() for _ in range(spec.num_input_limbs)]
        return v, (v,)
    if gate == pkg.UX_COMPARISON:
        a, b = val(), val()
        r = rng.random()
        if r < 0.2:
            b = a
        elif r < 0.4:
            b = a ^ (1 << rng.randrange(bits))
        return [a, b], (a, b)
    if gate == pkg.UX_ADD_MANY:
        ops = [([val() for _ in range(spec.num_addends)], rng.randrange(2) if edge != "max" else 1) for _ in range(spec.num_ops)]
        flat = [x for a, c in ops for x in (*a, c)]
        return flat, (ops,)
    if gate == pkg.UX_SUBTRACTION:
        ops = [(val(), val(), rng.randrange(2)) for _ in range(spec.num_ops)]
        return [x for op in ops for x in op], (ops,)
    ops = [(val(), val(), val()) for _ in range(spec.num_ops)]
    return [x for op in ops for x in op], (ops,)


# ------------------------------------------------------------------------------------------ CPU tier
@pytest.mark.parametrize("cfg", CONFIGS, ids=IDS)
def test_ux_spec_generated_rows_satisfy_gate(cfg):
    gate, spec, bits, lb, count, na = cfg
    rng = random.Random(77 + CONFIGS.index(cfg))
    # wires the gate leaves free on some rows: a chunk's equality dummy when the chunks are equal
    # (diff * dummy - (1 - eq) holds for any dummy), the arithmetic gate's inverse when low == 0
    free = set()
    if gate == pkg.UX_COMPARISON:
        free = set(range(4 + 2 * spec.num_chunks, 4 + 3 * spec.num_chunks))
    if gate == pkg.UX_ARITHMETIC:
        free = {6 * i + 5 for i in range(spec.num_ops)}
    for edge in (None, None, None, "zero", "max"):
        _, args = _inputs(gate, spec, bits, rng, edge)
        row, err = spec.generate(*args)
        c = spec.eval_unfiltered(row)
        assert err == 0 and len(row) == spec.num_wires and len(c) == spec.num_constraints
        assert not any(c)
        k = rng.choice([k for k in range(spec.num_wires) if k not in free])
        bad = list(row)
        bad[k] = (bad[k] + 1 + rng.randrange(5)) % GL_P
        assert any(spec.eval_unfiltered(bad)), f"corrupting wire {k} went unnoticed"


def test_ux_spec_comparison_result():
    g = U.UXComparisonGate(29, chunk_bits=2)
    for a, b in [(0, 0), (0, 1), (1, 0), ((1 << 29) - 1, (1 << 29) - 1), ((1 << 29) - 1, 0), (5 << 20, 5 << 20 | 1)]:
        row, _ = g.generate(a, b)
        assert row[2] == int(a <= b) and not any(g.eval_unfiltered(row))


def test_ux_reference_rows_fit_the_wire_budget():
    # nonnative.rs:454-460: 9 x 29-bit limbs and 8 x 34-bit limbs are range-checked by ONE gate row each;
    # with the default 3-bit aux limbs both fit plonky2's 135-wire standard config, with upstream's
    # 2-bit aux limbs they would not (the structural hint behind the default)
    assert U.UXRangeCheckGate(29, 3, 9).num_wires == 99 <= 135
    assert U.UXRangeCheckGate(34, 3, 8).num_wires == 104 <= 135
    assert U.UXRangeCheckGate(29, 2, 9).num_wires == 144 > 135
    assert U.UXAddManyGate.num_ops_for(3) == 5


@pytest.mark.parametrize("cfg", CONFIGS, ids=IDS)
def test_ux_cabi_shape_matches_spec(cfg):
    gate, spec, bits, lb, count, na = cfg
    import ctypes
    lib = pkg.load_library()
    nw, nc, ni = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.p2e_ux_gate_shape(gate, bits, lb, count, na, ctypes.addressof(nw), ctypes.addressof(nc), ctypes.addressof(ni)) == 0
    flat, _ = _inputs(gate, spec, bits, random.Random(0))
    assert (nw.value, nc.value, ni.value) == (spec.num_wires, spec.num_constraints, len(flat))


def test_ux_cabi_rejects_bad_geometry():
    lib = pkg.load_library()
    assert lib.p2e_ux_gate_shape(pkg.UX_RANGE_CHECK, 29, 5, 9, 0, None, None, None) < 0     # aux bits > 4
    assert lib.p2e_ux_gate_shape(pkg.UX_COMPARISON, 29, 2, 2, 0, None, None, None) < 0      # one comparison per row
    assert lib.p2e_ux_gate_shape(pkg.UX_COMPARISON, 9, 4, 1, 0, None, None, None) < 0       # chunk_bits != ceil(9 / 3)
    assert lib.p2e_ux_gate_shape(pkg.UX_ADD_MANY, 29, 2, 5, 17, None, None, None) < 0       # > 16 addends
    assert lib.p2e_ux_gate_shape(9, 29, 2, 5, 0, None, None, None) < 0                      # unknown gate
    assert b"ux" in lib.p2e_last_error()


# ------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=IDS)
def test_ux_witness_and_eval_vs_spec(engine, cfg):
    gate, spec, bits, lb, count, na = cfg
    rng = random.Random(1000 + CONFIGS.index(cfg))
    n = 67                                                   # ragged: not a multiple of the block size
    cols, rows = [], []
    for j in range(n):
        flat, args = _inputs(gate, spec, bits, rng, {0: "zero", 1: "max"}.get(j))
        cols.append(flat)
        rows.append(spec.generate(*args)[0])
    ins = np.array(cols, dtype=np.uint64).T.copy()
    got, err = engine.ux_gate_witness(gate, ins, bits, lb, count, na)
    assert not err.any()
    exp = np.array(rows, dtype=np.uint64).T
    assert np.array_equal(got, exp)
    # generated rows satisfy the gate on the device as well
    assert not engine.eval_ux_gate(gate, got, bits, lb, count, na).any()
    # arbitrary field elements (what an LDE row holds) and edge values
    w = np.array([[rng.randrange(GL_P) for _ in range(n)] for _ in range(spec.num_wires)], dtype=np.uint64)
    w[:, 0] = 0
    w[:, 1] = GL_P - 1
    w[:, 2] = 1
    c = engine.eval_ux_gate(gate, w, bits, lb, count, na)
    for j in range(n):
        assert [int(v) for v in c[:, j]] == spec.eval_unfiltered([int(v) for v in w[:, j]]), f"row {j}"


@pytest.mark.gpu
def test_ux_witness_flags_values_that_do_not_fit(engine):
    limbs = np.zeros((9, 5), dtype=np.uint64)
    limbs[3, 2] = 1 << 30                       # does not fit 10 x 3 bits
    limbs[8, 4] = (1 << 30) - 1                 # fits the aux limbs (30 bits), although it is not a 29-bit value
    row, err = engine.range_check_ux_witness(limbs, bits=29, aux_bits=3)
    assert list(err) == [0, 0, pkg.ERR_LIMB_RANGE, 0, 0]
    c = engine.eval_range_check_ux_gate(row, 9)
    assert c[:, 2].any() and not c[:, [0, 1, 3, 4]].any()
    row, err = engine.comparison_ux_witness([1 << 31, 3], [5, 4], num_bits=29, chunk_bits=2)
    assert list(err) == [pkg.ERR_LIMB_RANGE, 0] and row[2, 1] == 1


@pytest.mark.gpu
def test_ux_empty_batch(engine):
    row, err = engine.range_check_ux_witness(np.zeros((9, 0), dtype=np.uint64))
    assert row.shape == (99, 0) and err.shape == (0,)
    assert engine.eval_comparison_ux_gate(np.zeros((82, 0), dtype=np.uint64)).shape == (83, 0)


@pytest.mark.gpu
def test_ux_full_size_rows_vanish(engine):
    """Config-5 size (2^20 LDE rows): valid 9 x 29-bit range-check rows evaluate to all-zero
    constraints (size-independent property), device-resident."""
    import torch
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(5)
    limbs = torch.randint(0, 1 << 29, (9, n), dtype=torch.int64, device="cuda", generator=g)
    row = torch.empty((99, n), dtype=torch.int64, device="cuda")
    out = torch.empty((99, n), dtype=torch.int64, device="cuda")
    err = torch.empty(n, dtype=torch.int32, device="cuda")
    lib = engine.lib
    torch.cuda.synchronize()      # torch's stream and the engine's stream are unordered
    assert lib.p2e_ux_witness_dev(pkg.UX_RANGE_CHECK, 29, 3, 9, 0, n, limbs.data_ptr(), n, row.data_ptr(), n, err.data_ptr(), None) == 0
    assert lib.p2e_eval_ux_dev(pkg.UX_RANGE_CHECK, 29, 3, 9, 0, n, row.data_ptr(), n, out.data_ptr(), n, None) == 0
    engine.synchronize()
    torch.cuda.synchronize()
    assert not bool(out.any()) and not bool(err.any())
    # the aux limbs recompose to the inputs
    acc = torch.zeros(n, dtype=torch.int64, device="cuda")
    for j in reversed(range(10)):
        acc = acc * 8 + row[9 + j]
    assert torch.equal(acc, limbs[0])
