"""CPU tier: the host-built tape (C++ mirror of the gadget call graph) lays generators out exactly
in the oracle's gadget call order (same kinds, same tuple widths, same slot offsets)."""
import numpy as np

from oracle import spec as S
import plonky2_ecdsa_b200 as pkg


def _layout(fn_layout, fn_slots):
    cnt = fn_layout(None, None)
    kinds = np.empty(cnt, dtype=np.uint32)
    offs = np.empty(cnt, dtype=np.uint32)
    fn_layout(kinds.ctypes.data, offs.ctypes.data)
    return kinds, offs, int(fn_slots())


def _oracle_layout(tape):
    kinds, offs, off = [], [], 0
    for k, vals in tape:
        kinds.append(k)
        offs.append(off)
        off += len(vals)
    return kinds, offs, off


def test_ecdsa_tape_layout_matches_oracle():
    lib = pkg.load_library()
    kinds, offs, slots = _layout(lib.p2e_ecdsa_tape_layout, lib.p2e_ecdsa_tape_slots)
    sk, msg, k = 0x1234567 << 200 | 99, 0xabcdef << 180 | 5, 0x777 << 222 | 1
    pk = S.pt_mul(sk, S.G)
    r, s = S.sign_message(msg, sk, k)
    ok, oo, oslots = _oracle_layout(S.ecdsa_witness(msg, r, s, pk).tape)
    assert slots == oslots == 82615
    assert list(kinds) == ok
    assert list(offs) == oo


def test_glv_mul_tape_layout_matches_oracle():
    lib = pkg.load_library()
    kinds, offs, slots = _layout(lib.p2e_glv_mul_tape_layout, lib.p2e_glv_mul_tape_slots)
    wb = S.WitnessBuilder()
    P = S.pt_mul(7, S.G)
    kk = 0x1357 << 240 | 12345
    wb.glv_mul(wb.constant_affine_point(P), wb.constant_nonnative(kk, S.SECP_N))
    ok, oo, oslots = _oracle_layout(wb.tape)
    assert slots == oslots
    assert list(kinds) == ok and list(offs) == oo


def _tape(lib, tape):
    cnt = lib.p2e_tape_layout(tape, None, None)
    kinds = np.empty(cnt, dtype=np.uint32)
    offs = np.empty(cnt, dtype=np.uint32)
    lib.p2e_tape_layout(tape, kinds.ctypes.data, offs.ctypes.data)
    return list(kinds), list(offs), int(lib.p2e_tape_slots(tape)), lib.p2e_tape_inputs(tape)


def test_curve_op_tapes_match_oracle():
    """curve_add / curve_double / curve_conditional_add (gadgets/curve.rs:160-243)."""
    lib = pkg.load_library()
    P, Q = S.pt_mul(5, S.G), S.pt_mul(9, S.G)
    wb = S.WitnessBuilder()
    wb.curve_add(wb.constant_affine_point(P), wb.constant_affine_point(Q))
    assert _tape(lib, 2) == (*_oracle_layout(wb.tape), 4) and _tape(lib, 2)[2] == 231
    wb = S.WitnessBuilder()
    wb.curve_double(wb.constant_affine_point(P))
    assert _tape(lib, 3) == (*_oracle_layout(wb.tape), 2) and _tape(lib, 3)[2] == 282
    wb = S.WitnessBuilder()
    wb.curve_conditional_add(wb.constant_affine_point(P), wb.constant_affine_point(Q), 1)
    assert _tape(lib, 4) == (*_oracle_layout(wb.tape), 5) and _tape(lib, 4)[2] == 251
    assert lib.p2e_tape_slots(99) == 0 and lib.p2e_tape_inputs(-1) == -1


def test_fixed_base_and_msm_tapes_match_oracle():
    """fixed_base_curve_mul_circuit (curve_fixed_base.rs:18-66) and curve_msm_circuit with 9-limb
    scalars (curve_msm.rs:21-79: 131 two-bit digits)."""
    lib = pkg.load_library()
    wb = S.WitnessBuilder()
    wb.fixed_base_curve_mul(S.G, S.NN(S.limbs9(0x1234 << 230 | 77)))
    assert _tape(lib, 5) == (*_oracle_layout(wb.tape), 1)
    P, Q = S.pt_mul(5, S.G), S.pt_mul(9, S.G)
    wb = S.WitnessBuilder()
    wb.curve_msm(wb.constant_affine_point(P), wb.constant_affine_point(Q), S.NN(S.limbs9(3 << 250 | 1)), S.NN(S.limbs9(5 << 249 | 2)))
    assert _tape(lib, 6) == (*_oracle_layout(wb.tape), 6)


def test_compact_layout_map():
    """Compact host format: narrow rows for <= 32-bit slots, wide rows for check_sum / carry slots."""
    lib = pkg.load_library()
    slots = int(lib.p2e_tape_slots(0))
    m = np.empty(slots, dtype=np.uint32)
    import ctypes
    nn, nw = ctypes.c_uint32(0), ctypes.c_uint32(0)
    assert lib.p2e_tape_compact_layout(0, m.ctypes.data, ctypes.addressof(nn), ctypes.addressof(nw)) == slots
    assert nw.value == 1087 * 17 and nn.value + nw.value == slots + 1087      # one row of carry tops per CheckSum tuple
    wide = (m >> np.uint32(31)).astype(bool)
    carry = ((m >> np.uint32(30)) & np.uint32(1)).astype(bool) & ~wide
    first = ((m >> np.uint32(29)) & np.uint32(1)).astype(bool) & carry
    assert wide.sum() == nw.value and carry.sum() == 1087 * 16 and first.sum() == 1087
    # rows are dense and in slot order within each class; every CheckSum tuple is followed by its row of tops
    assert np.array_equal(m[wide] & np.uint32(0x7FFFFFFF), np.arange(nw.value, dtype=np.uint32))
    rows = np.where(carry, m & np.uint32(0x1FFFFFFF), m)[~wide].astype(np.int64)
    gaps = np.diff(rows)
    assert rows[0] == 0 and set(gaps.tolist()) <= {1, 2} and (gaps == 2).sum() == 1087 and rows[-1] < nn.value
    # oracle: exactly the check_sum / carry slots exceed 32 bits in a real witness
    sk, msg, k = 0x1234567 << 200 | 99, 0xabcdef << 180 | 5, 0x777 << 222 | 1
    r, s = S.sign_message(msg, sk, k)
    flat = np.array(S.ecdsa_witness(msg, r, s, S.pt_mul(sk, S.G)).flat(), dtype=np.uint64)
    assert not (flat[~wide & ~carry] >> np.uint64(32)).any()
    assert (flat[wide] >> np.uint64(32)).any()
    assert not (flat[carry] >> np.uint64(34)).any() and (flat[carry] >> np.uint64(32)).any()


def test_derived_compact_form_rebuilds_check_sum_slots_of_an_oracle_witness():
    """Derived compact form (include/p2e.h, wide == NULL): the check_sum columns a MulNonnativeGenerator writes
    (gates/mul_nonnative.rs:305-322) are exactly c_j 2^29 - c_{j-1} for the carries c_j = b_j - 2^33 the
    CheckSumGenerator derives from them (:513-531); checked on a witness computed by the oracle with the
    reference's literal arithmetic (Goldilocks column sums, field division per carry)."""
    import ctypes
    lib = pkg.load_library()
    slots = int(lib.p2e_tape_slots(0))
    m = np.empty(slots, dtype=np.uint32)
    nn, nw = ctypes.c_uint32(0), ctypes.c_uint32(0)
    lib.p2e_tape_compact_layout(0, m.ctypes.data, ctypes.addressof(nn), ctypes.addressof(nw))
    sk, msg, k = 0x7654321 << 190 | 3, 0xfedcba << 170 | 11, 0x999 << 211 | 7
    r, s = S.sign_message(msg, sk, k)
    flat = np.array(S.ecdsa_witness(msg, r, s, S.pt_mul(sk, S.G)).flat(), dtype=np.uint64)
    # pack what the device ships in the derived form: narrow rows only
    narrow = np.zeros((nn.value, 1), dtype=np.uint32)
    wide = (m >> np.uint32(31)).astype(bool)
    carry = ((m >> np.uint32(30)) & np.uint32(1)).astype(bool) & ~wide
    plain = ~wide & ~carry
    narrow[m[plain], 0] = flat[plain].astype(np.uint32)
    rows = (m[carry] & np.uint32(0x1FFFFFFF)).astype(np.int64)
    narrow[rows, 0] = (flat[carry] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    j = np.arange(rows.shape[0]) % 16
    tops = ((flat[carry] >> np.uint64(32)) & np.uint64(3)).astype(np.uint32) << (2 * j).astype(np.uint32)
    np.bitwise_or.at(narrow[:, 0], rows - j + 16, tops)
    got = pkg.Engine.expand_compact(m, narrow)
    assert np.array_equal(got[:, 0], flat)
    assert (flat[wide] >> np.uint64(62)).any()              # negative columns (P - |v|) are among them
    # the C decoder of the library (p2e_tape_compact_expand: host code, no device) agrees, in both forms, for n > 1
    n = 3
    narrow3 = np.ascontiguousarray(np.repeat(narrow, n, axis=1))
    narrow3[:, 1] = 0                                       # an all-zero instance next to the real ones
    out = np.empty((slots, n), dtype=np.uint64)
    assert lib.p2e_tape_compact_expand(0, n, narrow3.ctypes.data, None, out.ctypes.data) == 0
    assert np.array_equal(out[:, 0], flat) and np.array_equal(out[:, 2], flat)
    assert np.array_equal(out, pkg.Engine.expand_compact(m, narrow3))
    wide3 = np.zeros((nw.value, n), dtype=np.uint64)
    wide3[:, 0] = wide3[:, 2] = flat[wide]
    assert lib.p2e_tape_compact_expand(0, n, narrow3.ctypes.data, wide3.ctypes.data, out.ctypes.data) == 0
    assert np.array_equal(out[:, 0], flat) and np.array_equal(out, pkg.Engine.expand_compact(m, narrow3, wide3))
    assert lib.p2e_tape_compact_expand(99, n, narrow3.ctypes.data, None, out.ctypes.data) != 0


def test_mul_nonnative_instances_of_the_ecdsa_tape():
    """p2e_tape_num_mul / p2e_tape_mul_info: every mul_nonnative instance in gadget order, the first slot of its tuple and the
    limb counts of its factors (what nonnative.rs:453-455 range-checks) -- against the oracle's call graph."""
    import ctypes
    lib = pkg.load_library()
    kinds, offs, slots = _layout(lib.p2e_ecdsa_tape_layout, lib.p2e_ecdsa_tape_slots)
    # the oracle builds its inputs with constant_nonnative (leading zero limbs dropped), the tape takes 9-limb targets: use
    # values whose top limb is not zero
    sk, msg, k = 0x1234567 << 200 | 99, 0xabcdef << 233 | 5, 0x777 << 222 | 1
    pk = S.pt_mul(sk, S.G)
    r, s = S.sign_message(msg, sk, k)
    wb = S.ecdsa_witness(msg, r, s, pk)
    lib.p2e_tape_num_mul.restype = ctypes.c_int
    nm = lib.p2e_tape_num_mul(0)
    mul_offs = [int(o) for kd, o in zip(kinds, offs) if kd == S.K_MUL]
    assert nm == len(mul_offs) == len(wb.mul_ops) == 1087
    xl, yl, md = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    sl = ctypes.c_uint32(0)
    for j in range(nm):
        assert lib.p2e_tape_mul_info(0, j, ctypes.byref(xl), ctypes.byref(yl), ctypes.byref(sl), ctypes.byref(md)) == 0
        assert sl.value == mul_offs[j]
        assert (xl.value, yl.value) == (len(wb.mul_ops[j][0]), len(wb.mul_ops[j][1])), j
    assert lib.p2e_tape_mul_info(0, nm, None, None, None, None) != 0
    # every range_check_ux_circuit call: 2 per add_many, 1 per sub, 7 per mul (nonnative.rs:330-331, 373, 454-460)
    lib.p2e_tape_num_range_checks.restype = ctypes.c_int
    nc = lib.p2e_tape_num_range_checks(0)
    want = sum({S.K_ADDMANY: 2, S.K_SUB: 1, S.K_MUL: 7}.get(int(kd), 0) for kd in kinds)
    assert nc == want == 9168
    gen = np.empty(nc, dtype=np.int32)
    cnt = np.empty(nc, dtype=np.int32)
    bits = np.empty(nc, dtype=np.int32)
    assert lib.p2e_tape_range_check_info(0, 0, nc, gen.ctypes.data, cnt.ctypes.data, bits.ctypes.data) == 0
    off = np.empty(nc + 1, dtype=np.uint64)
    assert lib.p2e_tape_range_check_layout(0, 3, off.ctypes.data) == 0
    widths = cnt * (1 + (bits + 2) // 3)
    assert off[0] == 0 and np.array_equal(np.diff(off.astype(np.int64)), widths)
    assert set(bits.tolist()) == {29, 34} and (gen[1:] >= gen[:-1]).all()
    assert lib.p2e_tape_range_check_info(0, 0, nc + 1, None, None, None) != 0


def test_mul_factor_references_resolve_to_the_oracles_factors():
    """Tape::muls on the CPU tier: the two factor references of every mul_nonnative instance (witness slots, inputs,
    constants, constant-point coordinates, bool factors), resolved on the host against the oracle's witness of one
    signature, are the factors the oracle multiplied -- and x * y = q * m + r holds for the tuple at the recorded slot."""
    import ctypes
    lib = pkg.load_library()
    sk, msg, k = 0x7654321 << 200 | 3, 0x7dcba9 << 232 | 9, 0x555 << 222 | 7       # all below the group order, top limb not zero
    pk = S.pt_mul(sk, S.G)
    r, s = S.sign_message(msg, sk, k)
    wb = S.ecdsa_witness(msg, r, s, pk)
    flat = np.array(wb.flat(), dtype=np.uint64)
    ins = [np.array(S.convert_base(S.to_u32_digits(v), 32, 29) + [0] * 9, dtype=np.uint64)[:9].copy() for v in (msg, r, s, pk[0], pk[1])]
    arr = (ctypes.c_void_p * 5)(*[a.ctypes.data for a in ins])
    x9, y9 = np.zeros(9, dtype=np.uint64), np.zeros(9, dtype=np.uint64)
    xl, yl, md = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    sl = ctypes.c_uint32(0)
    lib.p2e_tape_mul_factors_host.restype = ctypes.c_int
    lib.p2e_tape_mul_factors_host.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
    val = lambda limbs: sum(int(v) << (29 * i) for i, v in enumerate(limbs))
    for j in range(lib.p2e_tape_num_mul(0)):
        assert lib.p2e_tape_mul_factors_host(0, j, arr, flat.ctypes.data, x9.ctypes.data, y9.ctypes.data) == 0
        ox, oy = wb.mul_ops[j]
        assert [int(v) for v in x9] == list(ox) + [0] * (9 - len(ox)), j
        assert [int(v) for v in y9] == list(oy) + [0] * (9 - len(oy)), j
        assert lib.p2e_tape_mul_info(0, j, ctypes.byref(xl), ctypes.byref(yl), ctypes.byref(sl), ctypes.byref(md)) == 0
        m = S.MODULI[md.value]
        rr, qq = val(flat[sl.value: sl.value + 9]), val(flat[sl.value + 9: sl.value + 18])
        assert val(x9) * val(y9) == qq * m + rr and rr < m, j
    assert lib.p2e_tape_mul_factors_host(0, 99999, arr, flat.ctypes.data, x9.ctypes.data, y9.ctypes.data) != 0


def test_range_check_call_lists_of_every_builtin_tape():
    """2 calls per add_many, 1 per sub, 7 per mul_nonnative (nonnative.rs:330-331, 373, 454-460) on all seven built-in tapes;
    generator indices point at those kinds; wire offsets are the running sum of inputs * (1 + aux limbs)."""
    import ctypes
    lib = pkg.load_library()
    lib.p2e_tape_num_range_checks.restype = ctypes.c_int
    lib.p2e_tape_num_mul.restype = ctypes.c_int
    per_kind = {S.K_ADDMANY: 2, S.K_SUB: 1, S.K_MUL: 7}
    for tape in range(7):
        kinds, offs, slots, _ = _tape(lib, tape)
        nc = lib.p2e_tape_num_range_checks(tape)
        assert nc == sum(per_kind.get(int(k), 0) for k in kinds), tape
        assert lib.p2e_tape_num_mul(tape) == sum(1 for k in kinds if k == S.K_MUL)
        if nc == 0:
            continue
        gen = np.empty(nc, dtype=np.int32)
        cnt = np.empty(nc, dtype=np.int32)
        bits = np.empty(nc, dtype=np.int32)
        assert lib.p2e_tape_range_check_info(tape, 0, nc, gen.ctypes.data, cnt.ctypes.data, bits.ctypes.data) == 0
        seen = {}
        for g in gen:
            seen[int(g)] = seen.get(int(g), 0) + 1
        for g, c in seen.items():
            assert per_kind[int(kinds[g])] == c, (tape, g)
        assert ((bits == 29) | (bits == 34)).all() and (cnt >= 0).all() and (cnt <= 9).all() and (cnt[bits == 34] == 8).all()
        for lb in (1, 2, 3, 4):
            off = np.empty(nc + 1, dtype=np.uint64)
            assert lib.p2e_tape_range_check_layout(tape, lb, off.ctypes.data) == 0
            assert np.array_equal(np.diff(off.astype(np.int64)), cnt * (1 + -(-bits // lb)))
        assert lib.p2e_tape_range_check_layout(tape, 0, None) != 0 and lib.p2e_tape_range_check_layout(tape, 9, None) != 0
