"""GPU tier: the plonky2_ux range-check rows UNDER the in-tree generators.  Every range_check_ux_circuit call of the crate is
in gadgets/nonnative.rs (add_many :330-331, sub :373, mul :454-460) and each is one gate row; the engine produces them from
the finished device witness for every call of the ECDSA circuit.  Checked against the oracle's call graph (pinned: the call
list and its inputs are the reference's own) and the restated gate spec (oracle/ux_spec.py: the row layout is PARITY UNPINNED)."""
import numpy as np
import pytest

from oracle import spec as S
from oracle import ux_spec as UX
from plonky2_ecdsa_b200 import workload as W

pytestmark = pytest.mark.gpu


def expected_calls(wb):
    """The range_check_ux_circuit calls of one circuit instance in gadget order: (generator index, limbs, bits)."""
    calls = []
    mi = 0
    for g, (kind, vals) in enumerate(wb.tape):
        if kind == S.K_ADDMANY:                      # nonnative.rs:330-331
            calls.append((g, vals[:9], 29))
            calls.append((g, vals[9:10], 29))
        elif kind == S.K_SUB:                        # nonnative.rs:373
            calls.append((g, vals[:9], 29))
        elif kind == S.K_MUL:                        # nonnative.rs:454-460
            x, y = wb.mul_ops[mi]
            mi += 1
            r, q = vals[:9], vals[9:18]
            b = wb.tape[g + 1][1]
            assert wb.tape[g + 1][0] == S.K_CHECKSUM and len(b) == 16
            calls += [(g, x, 29), (g, y, 29), (g, r, 29), (g, q, 29), (g, r, 29), (g, b[:8], 34), (g, b[8:], 34)]
    assert mi == len(wb.mul_ops)
    return calls


@pytest.mark.parametrize("limb_bits", [3, 2])
def test_range_check_rows_of_the_ecdsa_circuit(engine, limb_bits):
    n = 4
    wl = W.ecdsa_workload(n, seed=97)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    wit, err = engine.verify_secp256k1_message_witness(*ins)
    assert not err.any()
    narrow, _, err2 = engine.tape_witness_compact(engine.TAPE_ECDSA, ins, derive_check_sum=True)
    assert not err2.any()
    gen, cnt, bits, off = engine.tape_range_checks(engine.TAPE_ECDSA, limb_bits)
    nc = len(gen)
    rows_p, e_p = engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, wit, 0, nc, limb_bits)
    rows_c, e_c = engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, narrow, 0, nc, limb_bits, compact=True)
    assert not e_p.any() and not e_c.any()
    assert rows_p.shape[0] == int(off[nc]) and np.array_equal(rows_p, rows_c), "plain and compact witness give different rows"
    i = wl["ints"]
    for j in range(n):
        wb = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j]))
        exp = expected_calls(wb)
        assert len(exp) == nc == 9168
        for c, (g, limbs, b) in enumerate(exp):
            assert (g, len(limbs), b) == (int(gen[c]), int(cnt[c]), int(bits[c])), f"call {c}"
            row, bad = UX.UXRangeCheckGate(b, limb_bits, len(limbs)).generate(limbs)
            assert not bad and len(row) == int(off[c + 1] - off[c])
            got = rows_p[int(off[c]): int(off[c + 1]), j]
            assert [int(v) for v in got] == row, f"signature {j} call {c} (generator {g})"


def test_range_check_rows_sub_range_and_errors(engine):
    n = 3
    wl = W.ecdsa_workload(n, seed=5)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    wit, _ = engine.verify_secp256k1_message_witness(*ins)
    gen, cnt, bits, off = engine.tape_range_checks(engine.TAPE_ECDSA)
    nc = len(gen)
    full, _ = engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, wit, 0, nc)
    part, _ = engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, wit, 1000, 1400)
    assert np.array_equal(part, full[int(off[1000]): int(off[1400])])
    with pytest.raises(Exception):
        engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, wit, 0, nc + 1)
    # a carry that does not fit its aux limbs makes its gate row unsatisfiable: flagged for that signature only
    c = int(np.nonzero(bits == 34)[0][10])
    kinds, offs, _ = engine.ecdsa_tape_layout()
    bad = wit.copy()
    bad[int(offs[gen[c]]) + 35, 1] = np.uint64(1 << 40)
    _, e = engine.tape_range_check_rows(engine.TAPE_ECDSA, ins, bad, c, c + 1)
    assert e[1] != 0 and e[0] == 0 and e[2] == 0


def test_range_check_rows_of_a_glv_mul_chain(engine):
    """A second tape (glv_mul: GLV decomposition + msm, gadgets/glv.rs:87-104): factors with 5 limbs and constants."""
    n = 3
    ks = [0x1357 << 240 | 12345, S.SECP_N - 2, 77]
    P = S.pt_mul(7, S.G)
    k = W.ints_to_limbs(ks)
    px, py = W.ints_to_limbs([P[0]] * n), W.ints_to_limbs([P[1]] * n)
    wit, err = engine.glv_mul_witness(k, px, py)
    assert not err.any()
    gen, cnt, bits, off = engine.tape_range_checks(engine.TAPE_GLV_MUL)
    rows, e = engine.tape_range_check_rows(engine.TAPE_GLV_MUL, [k, px, py], wit, 0, len(gen))
    assert not e.any()
    for j in range(n):
        wb = S.WitnessBuilder()
        wb.glv_mul(wb.constant_affine_point(P), wb.constant_nonnative(ks[j], S.SECP_N))
        exp = expected_calls(wb)
        assert len(exp) == len(gen)
        for c, (g, limbs, b) in enumerate(exp):
            assert (g, b) == (int(gen[c]), int(bits[c]))
            # the oracle builds k and P as constants (leading zero limbs dropped), the tape takes 9-limb targets: compare the
            # limbs the oracle has, the tape's extra ones must be zero
            row, bad = UX.UXRangeCheckGate(b, 3, int(cnt[c])).generate(list(limbs) + [0] * (int(cnt[c]) - len(limbs)))
            assert not bad and int(cnt[c]) >= len(limbs)
            assert [int(v) for v in rows[int(off[c]): int(off[c + 1]), j]] == row, f"scalar {j} call {c}"
