"""GPU tier: the plonky2_ux range-check rows UNDER the in-tree generators -- the seven range_check_ux_circuit rows every
mul_nonnative adds (gadgets/nonnative.rs:453-460) -- produced from the finished device witness for every mul_nonnative
instance of the ECDSA circuit, against the restated gate spec (oracle/ux_spec.py, PARITY UNPINNED for the row layout; the
call list and its inputs are the reference's own)."""
import numpy as np
import pytest

from oracle import spec as S
from oracle import ux_spec as UX
from plonky2_ecdsa_b200 import workload as W

pytestmark = pytest.mark.gpu


def _expected_rows(wb, limb_bits):
    """[instances][7] -> (wire list, count) from the oracle's witness builder."""
    muls = [v for k, v in wb.tape if k == S.K_MUL]
    sums = [v for k, v in wb.tape if k == S.K_CHECKSUM]
    assert len(muls) == len(sums) == len(wb.mul_ops)
    out = []
    for (x, y), mt, b in zip(wb.mul_ops, muls, sums):
        r, q = mt[:9], mt[9:18]
        calls = [(x, 29), (y, 29), (r, 29), (q, 29), (r, 29), (b[:8], 34), (b[8:], 34)]
        rows = []
        for limbs, bits in calls:
            row, err = UX.UXRangeCheckGate(bits, limb_bits, len(limbs)).generate(limbs)
            assert err == 0
            rows.append(row)
        out.append(rows)
    return out


@pytest.mark.parametrize("limb_bits", [3, 2])
def test_mul_range_check_rows_of_the_ecdsa_circuit(engine, limb_bits):
    n = 4
    wl = W.ecdsa_workload(n, seed=97)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    wit, err = engine.verify_secp256k1_message_witness(*ins)
    assert not err.any()
    narrow, _, err2 = engine.tape_witness_compact(engine.TAPE_ECDSA, ins, derive_check_sum=True)
    assert not err2.any()
    info = engine.tape_mul_info(engine.TAPE_ECDSA)
    per, offs, wid = engine.tape_mul_range_rows_shape(limb_bits)
    a29, a34 = -(-29 // limb_bits), -(-34 // limb_bits)
    assert per == 5 * 9 * (1 + a29) + 2 * 8 * (1 + a34) and wid == [9 * (1 + a29)] * 5 + [8 * (1 + a34)] * 2
    nm = len(info)
    rows_p, e_p = engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, wit, 0, nm, limb_bits)
    rows_c, e_c = engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, narrow, 0, nm, limb_bits, compact=True)
    assert not e_p.any() and not e_c.any()
    assert np.array_equal(rows_p, rows_c), "plain and compact witness give different rows"
    i = wl["ints"]
    for j in range(n):
        wb = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j]))
        exp = _expected_rows(wb, limb_bits)
        assert len(exp) == nm
        for m, rows in enumerate(exp):
            assert (len(wb.mul_ops[m][0]), len(wb.mul_ops[m][1])) == info[m][:2], f"instance {m}: factor limb counts"
            for c, row in enumerate(rows):
                got = rows_p[m * per + offs[c]: m * per + offs[c] + wid[c], j]
                assert [int(v) for v in got[:len(row)]] == row, f"signature {j} instance {m} call {c}"
                assert not got[len(row):].any()


def test_mul_range_check_rows_sub_range_and_errors(engine):
    n = 3
    wl = W.ecdsa_workload(n, seed=5)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    wit, _ = engine.verify_secp256k1_message_witness(*ins)
    nm = len(engine.tape_mul_info(engine.TAPE_ECDSA))
    per, _, _ = engine.tape_mul_range_rows_shape(3)
    full, _ = engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, wit, 0, nm)
    part, _ = engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, wit, 100, 140)
    assert np.array_equal(part, full[100 * per: 140 * per])
    with pytest.raises(Exception):
        engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, wit, 0, nm + 1)
    # a carry that does not fit 34 + limb padding bits makes its gate row unsatisfiable: flagged
    bad = wit.copy()
    slot = engine.tape_mul_info(engine.TAPE_ECDSA)[7][2]
    bad[slot + 35, 1] = np.uint64(1 << 40)
    _, e = engine.tape_mul_range_rows(engine.TAPE_ECDSA, ins, bad, 7, 8)
    assert e[1] != 0 and e[0] == 0 and e[2] == 0
