"""GPU tier: whole-chain witnesses (verify_secp256k1_message_circuit, glv_mul) vs the oracle tape."""
import ctypes

import numpy as np
import pytest

from oracle import spec as S
from plonky2_ecdsa_b200 import workload as W

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2])
def replay_lanes(request, engine):
    """Both forms of the Jacobian replay kernel (p2e_set_option("chain_replay_lanes")): one lane per (signature, chain) --
    what calls of 4096 signatures or more run -- and the two-lane form of smaller calls."""
    old = engine.get_option("chain_replay_lanes")
    engine.set_option("chain_replay_lanes", request.param)
    yield request.param
    engine.set_option("chain_replay_lanes", old)


def _first_diff(col, ref, kinds, offs):
    for idx, (a, b) in enumerate(zip(col, ref)):
        if int(a) != b:
            g = int(np.searchsorted(offs, idx, side="right")) - 1
            return f"slot {idx} (generator #{g}, kind {kinds[g]}, +{idx - int(offs[g])}): got {int(a)} want {b}"
    return None


def test_ecdsa_witness_vs_oracle(engine, replay_lanes):
    n = 6
    wl = W.ecdsa_workload(n, seed=4242)
    wit, err = engine.verify_secp256k1_message_witness(wl["msg"], wl["r"], wl["s"], wl["pkx"], wl["pky"])
    kinds, offs, slots = engine.ecdsa_tape_layout()
    assert wit.shape == (slots, n)
    i = wl["ints"]
    for j in range(n):
        ref = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j])).flat()
        d = _first_diff(wit[:, j], ref, kinds, offs)
        assert d is None, f"signature {j}: {d}"
    assert not err.any()


def test_ecdsa_bad_signature_flags_connect(engine):
    wl = W.ecdsa_workload(3, seed=7)
    s_bad = wl["s"].copy()
    s_bad[0, 1] ^= np.uint64(1)                        # tamper signature 1
    wit, err = engine.verify_secp256k1_message_witness(wl["msg"], wl["r"], s_bad, wl["pkx"], wl["pky"])
    assert err[0] == 0 and err[2] == 0
    assert err[1] & 128                                 # ERR_CONNECT: connect_nonnative(r, point.x) cannot hold
    pk_bad = wl["pky"].copy()
    pk_bad[0, 0] ^= np.uint64(1)                       # pk not on the curve -> curve_assert_valid fails
    wit, err = engine.verify_secp256k1_message_witness(wl["msg"], wl["r"], wl["s"], wl["pkx"], pk_bad)
    assert err[0] & 128 and err[1] == 0


def test_glv_mul_witness_vs_oracle(engine, replay_lanes):
    n = 5
    ks = [(0x1357 << 240 | 12345), S.SECP_N - 2, 3, 1 << 200, 0xdeadbeef << 128]
    pts = [S.pt_mul(7 + j, S.G) for j in range(n)]
    k = W.ints_to_limbs(ks)
    px = W.ints_to_limbs([p[0] for p in pts])
    py = W.ints_to_limbs([p[1] for p in pts])
    wit, err = engine.glv_mul_witness(k, px, py)
    kinds, offs, slots = engine.glv_mul_tape_layout()
    for j in range(n):
        wb = S.WitnessBuilder()
        res = wb.glv_mul(wb.constant_affine_point(pts[j]), wb.constant_nonnative(ks[j], S.SECP_N))
        d = _first_diff(wit[:, j], wb.flat(), kinds, offs)
        assert d is None, f"scalar {j}: {d}"
        # the chain really computes k * P
        assert (S.get_biguint(res[0].limbs), S.get_biguint(res[1].limbs)) == S.pt_mul(ks[j], pts[j])
    assert not err.any()


def test_ecdsa_full_batch_properties(engine):
    """Config 3 at full size (4096 signatures): no error rows (every connect_* holds, i.e. the
    circuit's final point.x == r), chunked execution is consistent, sampled tapes match the oracle."""
    n = 4096
    wl = W.ecdsa_workload(n)
    wit, err = engine.verify_secp256k1_message_witness(wl["msg"], wl["r"], wl["s"], wl["pkx"], wl["pky"])
    assert not err.any()
    kinds, offs, slots = engine.ecdsa_tape_layout()
    # last generator = y3 of the final curve_add; x3 sits 71 slots before it: must equal r limb-wise
    x3 = wit[slots - 10 - 51 - 10 - 10: slots - 10 - 51 - 10 - 1]
    assert np.array_equal(x3, wl["r"])
    i = wl["ints"]
    for j in (0, 1777, 4095):
        ref = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j])).flat()
        assert _first_diff(wit[:, j], ref, kinds, offs) is None


def _rand_points(n, seed):
    import random
    rng = random.Random(seed)
    return [S.pt_mul(rng.randrange(1, S.SECP_N), S.G) for _ in range(n)]


def test_curve_add_double_conditional_add_vs_oracle(engine, replay_lanes):
    """Mirrors the reference's test_curve_add / test_curve_double / test_curve_conditional_add
    (gadgets/curve.rs:359-456): random points, result must equal the native sum, and here the whole
    generator tape must equal the oracle's."""
    n = 40
    P, Q = _rand_points(n, 1), _rand_points(n, 2)
    x1, y1 = W.ints_to_limbs([p[0] for p in P]), W.ints_to_limbs([p[1] for p in P])
    x2, y2 = W.ints_to_limbs([p[0] for p in Q]), W.ints_to_limbs([p[1] for p in Q])
    b = [j % 2 for j in range(n)]
    wa, ea = engine.curve_add(x1, y1, x2, y2)
    wd, ed = engine.curve_double(x1, y1)
    wc, ec = engine.curve_conditional_add(x1, y1, x2, y2, b)
    assert not ea.any() and not ed.any() and not ec.any()
    for j in range(n):
        wb = S.WitnessBuilder()
        r = wb.curve_add(wb.constant_affine_point(P[j]), wb.constant_affine_point(Q[j]))
        assert [int(v) for v in wa[:, j]] == wb.flat()
        assert (S.get_biguint(r[0].limbs), S.get_biguint(r[1].limbs)) == S.pt_add(P[j], Q[j])
        wb = S.WitnessBuilder()
        r = wb.curve_double(wb.constant_affine_point(P[j]))
        assert [int(v) for v in wd[:, j]] == wb.flat()
        assert (S.get_biguint(r[0].limbs), S.get_biguint(r[1].limbs)) == S.pt_add(P[j], P[j])
        wb = S.WitnessBuilder()
        r = wb.curve_conditional_add(wb.constant_affine_point(P[j]), wb.constant_affine_point(Q[j]), b[j])
        assert [int(v) for v in wc[:, j]] == wb.flat()
        want = S.pt_add(P[j], Q[j]) if b[j] else P[j]
        assert (S.get_biguint(r[0].limbs), S.get_biguint(r[1].limbs)) == want


def test_curve_add_equal_x_flags_inverse_zero(engine, replay_lanes):
    """Adding a point to itself with the incomplete formulas: x2 - x1 = 0, the reference's inverse() panics."""
    P = _rand_points(3, 3)
    x, y = W.ints_to_limbs([p[0] for p in P]), W.ints_to_limbs([p[1] for p in P])
    x2, y2 = x.copy(), y.copy()
    Q = S.pt_mul(77, S.G)
    x2[:, 1] = W.ints_to_limbs([Q[0]])[:, 0]
    y2[:, 1] = W.ints_to_limbs([Q[1]])[:, 0]
    _, err = engine.curve_add(x, y, x2, y2)
    assert err[0] & 8 and err[2] & 8 and err[1] == 0


def test_chain_chunking_is_transparent(engine):
    """Batches above the scratch chunk capacity are processed chunk by chunk: same bits."""
    n = 300
    ks = [(j * 0x9E3779B97F4A7C15 + 12345) % S.SECP_N for j in range(n)]
    pts = _rand_points(4, 9)
    k = W.ints_to_limbs(ks)
    px = W.ints_to_limbs([pts[j % 4][0] for j in range(n)])
    py = W.ints_to_limbs([pts[j % 4][1] for j in range(n)])
    w1, e1 = engine.glv_mul_witness(k, px, py)
    old = engine.get_option("chain_chunk")
    engine.set_option("chain_chunk", 128)
    try:
        w2, e2 = engine.glv_mul_witness(k, px, py)
    finally:
        engine.set_option("chain_chunk", old)
    assert np.array_equal(w1, w2) and np.array_equal(e1, e2) and not e1.any()
    wb = S.WitnessBuilder()
    wb.glv_mul(wb.constant_affine_point(pts[299 % 4]), wb.constant_nonnative(ks[299], S.SECP_N))
    assert [int(v) for v in w2[:, 299]] == wb.flat()


def test_replay_kernel_odd_batch_is_bit_exact(engine, replay_lanes):
    """The two-lane replay kernel pairs lanes 2j / 2j+1: an odd batch size leaves a half-filled warp."""
    wl = W.ecdsa_workload(71, seed=31)
    args = (wl["msg"], wl["r"], wl["s"], wl["pkx"], wl["pky"])
    w2, e2 = engine.verify_secp256k1_message_witness(*args)
    assert not e2.any()
    i = wl["ints"]
    for j in (0, 35, 70):
        ref = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j])).flat()
        assert [int(v) for v in w2[:, j]] == ref


def test_ecdsa_edge_inputs_match_reference_panics(engine, replay_lanes):
    """Rows where the reference would panic get an error bit; neighbours in the same batch are untouched.
    s = 0 -> FF::inverse() of zero (nonnative.rs:863); s >= 2^256 -> from_noncanonical_biguint panics;
    msg limb >= 2^29 -> MulNonnativeGenerator's limb assert (mul_nonnative.rs:275-276); a non-canonical
    but < 2^256 message (msg + n) is legal for the generators and must reproduce the oracle's tape."""
    n = 6
    wl = W.ecdsa_workload(n, seed=77)
    msg, r, s, pkx, pky = (wl[k].copy() for k in ("msg", "r", "s", "pkx", "pky"))
    i = wl["ints"]
    s[:, 1] = 0                                                     # inverse of zero
    s[:, 2] = W.ints_to_limbs([(1 << 256) + 5])[:, 0]               # >= 2^256
    msg[0, 3] = np.uint64(1 << 29)                                  # limb out of range
    big = i["msg"][4] + S.SECP_N
    if big < (1 << 256):
        msg[:, 4] = W.ints_to_limbs([big])[:, 0]                    # non-canonical message
    wit, err = engine.verify_secp256k1_message_witness(msg, r, s, pkx, pky)
    assert err[0] == 0 and err[5] == 0
    assert err[1] & 8, "s = 0 must raise ERR_INVERSE_ZERO"
    assert err[2] & 4, "s >= 2^256 must raise ERR_NONCANONICAL"
    assert err[3] & 2, "limb >= 2^29 must raise ERR_LIMB_RANGE"
    for j in (0, 5):
        ref = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j])).flat()
        assert [int(v) for v in wit[:, j]] == ref
    # oracle agrees on which rows panic
    for j, exc in ((1, True), (2, True), (3, True)):
        with pytest.raises(S.RefPanic):
            wb = S.WitnessBuilder()
            wb.verify_secp256k1_message(S.NN([int(v) for v in msg[:, j]]), S.NN([int(v) for v in r[:, j]]),
                                        S.NN([int(v) for v in s[:, j]]),
                                        (S.NN([int(v) for v in pkx[:, j]]), S.NN([int(v) for v in pky[:, j]])))
    if big < (1 << 256):
        wb = S.WitnessBuilder()
        # u1 = (msg + n) * c mod n has the same r but a different quotient: the tape differs from row 4's canonical one
        try:
            wb.verify_secp256k1_message(S.NN(S.limbs9(big)), S.NN(S.limbs9(i["r"][4])), S.NN(S.limbs9(i["s"][4])),
                                        (S.NN(S.limbs9(i["pkx"][4])), S.NN(S.limbs9(i["pky"][4]))))
            assert [int(v) for v in wit[:, 4]] == wb.flat() and err[4] == 0
        except S.RefPanic:
            assert err[4] != 0


def test_fixed_base_mul_vs_oracle(engine, replay_lanes):
    """Mirrors test_fixed_base (gadgets/curve_fixed_base.rs:86-116): result == scalar * G."""
    import random
    rng = random.Random(55)
    ks = [rng.randrange(1, S.SECP_N) for _ in range(6)] + [1, S.SECP_N - 1, 16, (1 << 255) + 12345]
    wit, err = engine.fixed_base_curve_mul(W.ints_to_limbs(ks))
    assert not err.any()
    for j, k in enumerate(ks):
        wb = S.WitnessBuilder()
        res = wb.fixed_base_curve_mul(S.G, S.NN(S.limbs9(k)))
        assert [int(v) for v in wit[:, j]] == wb.flat()
        assert (S.get_biguint(res[0].limbs), S.get_biguint(res[1].limbs)) == S.pt_mul(k, S.G)


def test_curve_msm_vs_oracle(engine, replay_lanes):
    """Mirrors test_curve_msm (gadgets/curve_msm.rs:98-136): result == n*p + m*q with full-width scalars."""
    import random
    rng = random.Random(56)
    cnt = 5
    P, Q = _rand_points(cnt, 57), _rand_points(cnt, 58)
    ns = [rng.randrange(1, S.SECP_N) for _ in range(cnt)]
    ms = [rng.randrange(1, S.SECP_N) for _ in range(cnt)]
    wit, err = engine.curve_msm(W.ints_to_limbs([p[0] for p in P]), W.ints_to_limbs([p[1] for p in P]),
                                W.ints_to_limbs([p[0] for p in Q]), W.ints_to_limbs([p[1] for p in Q]),
                                W.ints_to_limbs(ns), W.ints_to_limbs(ms))
    assert not err.any()
    for j in range(cnt):
        wb = S.WitnessBuilder()
        res = wb.curve_msm(wb.constant_affine_point(P[j]), wb.constant_affine_point(Q[j]), S.NN(S.limbs9(ns[j])), S.NN(S.limbs9(ms[j])))
        assert [int(v) for v in wit[:, j]] == wb.flat()
        want = S.pt_add(S.pt_mul(ns[j], P[j]), S.pt_mul(ms[j], Q[j]))
        assert (S.get_biguint(res[0].limbs), S.get_biguint(res[1].limbs)) == want


def test_compact_host_format_round_trips(engine):
    """The compact host format (u32 narrow slots, u64 check_sum slots, packed 34-bit carries) carries exactly the plain witness."""
    wl = W.ecdsa_workload(37, seed=5)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    plain, e1 = engine.verify_secp256k1_message_witness(*ins)
    cmap, nn, nw = engine.tape_compact_layout(engine.TAPE_ECDSA)
    assert nn + nw == plain.shape[0] + 1087 and plain.shape[0] == 82615 and nw == 1087 * 17
    narrow, wide, e2 = engine.tape_witness_compact(engine.TAPE_ECDSA, ins)
    assert narrow.shape == (nn, 37) and wide.shape == (nw, 37)
    assert np.array_equal(engine.expand_compact(cmap, narrow, wide), plain) and np.array_equal(e1, e2)
    # every narrow slot really fits 32 bits in the plain layout
    assert not (plain[(cmap >> np.uint32(30)) == 0] >> np.uint64(32)).any()
    # ... and every carry slot fits 34 bits (its two top bits travel in the tuple's packed row)
    carry = (cmap >> np.uint32(30)) == 1
    assert carry.sum() == 1087 * 16 and not (plain[carry] >> np.uint64(34)).any()


@pytest.mark.gpu
def test_compact_host_format_chunked_copy(engine):
    """Batches of >= 1024 signatures go through the chunked path (pitched D2H copies overlapping the next
    chunk's kernels); ragged last chunk included."""
    n = 1100 + 3
    wl = W.ecdsa_workload(n, seed=8)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    ins[2] = ins[2].copy()
    ins[2][:, 777] = 0                                      # s = 0 in the third chunk -> ERR_INVERSE_ZERO there only
    plain, e1 = engine.verify_secp256k1_message_witness(*ins)
    cmap, nn, nw = engine.tape_compact_layout(engine.TAPE_ECDSA)
    narrow, wide, e2 = engine.tape_witness_compact(engine.TAPE_ECDSA, ins)
    assert np.array_equal(e1, e2) and e1[777] & 8 and not np.delete(e1, 777).any()
    ok = np.ones(n, dtype=bool)
    ok[777] = False
    assert np.array_equal(engine.expand_compact(cmap, narrow, wide)[:, ok], plain[:, ok])


@pytest.mark.gpu
def test_compact_derived_form_rebuilds_check_sum_from_carries(engine):
    """Derived compact form (wide == NULL): the check_sum slots are not shipped; expand_compact rebuilds them from
    the CheckSum carries and the result is the plain witness, bit for bit (chunked path, ragged last chunk)."""
    n = 1100 + 3
    wl = W.ecdsa_workload(n, seed=18)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    plain, e1 = engine.verify_secp256k1_message_witness(*ins)
    cmap, nn, nw = engine.tape_compact_layout(engine.TAPE_ECDSA)
    narrow, wide, e2 = engine.tape_witness_compact(engine.TAPE_ECDSA, ins, derive_check_sum=True)
    assert wide is None and narrow.shape == (nn, n)
    assert not e1.any() and not e2.any()                    # the device verified the identity for every row
    assert np.array_equal(engine.expand_compact(cmap, narrow), plain)
    out = np.empty_like(plain)
    assert engine.lib.p2e_tape_compact_expand(engine.TAPE_ECDSA, n, narrow.ctypes.data, None, out.ctypes.data) == 0
    assert np.array_equal(out, plain)                       # the library's own host-side decoder
    # the generic tapes as well (curve_add: 3 products)
    wl = W.ecdsa_workload(5, seed=3)
    pts = [wl["pkx"], wl["pky"], np.roll(wl["pkx"], 1, axis=1), np.roll(wl["pky"], 1, axis=1)]
    plain, _ = engine.tape_witness(engine.TAPE_CURVE_ADD, pts)
    cmap, nn, nw = engine.tape_compact_layout(engine.TAPE_CURVE_ADD)
    narrow, _, e = engine.tape_witness_compact(engine.TAPE_CURVE_ADD, pts, derive_check_sum=True)
    assert not e.any() and np.array_equal(engine.expand_compact(cmap, narrow), plain)


@pytest.mark.gpu
def test_ecdsa_batch_vs_c_oracle(engine, replay_lanes):
    """Every slot of 768 complete witnesses against the C restatement of the reference's generators
    (oracle/p2e_oracle.c, itself pinned to the Python spec by tests/test_oracle_c.py): one Fermat ladder
    per curve op on the CPU side, batched inversion on the GPU side, identical bits."""
    import os
    from oracle import cpu_baseline as CB
    lib = CB.load()
    n = 768
    wl = W.ecdsa_workload(n, seed=31337)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    wit, err = engine.verify_secp256k1_message_witness(*ins)
    ref, ref_err = CB.ecdsa_witness(lib, *ins, threads=os.cpu_count() or 1)
    assert not err.any() and not ref_err.any()
    assert np.array_equal(wit, ref)


def test_compact_device_format_is_written_directly(engine):
    """p2e_tape_witness_compact_dev: the chain kernels write the compact rows themselves (with and without the
    check_sum columns, padded stride); expanding them gives the plain device witness bit for bit."""
    import torch
    n, stride = 157, 192
    wl = W.ecdsa_workload(n, seed=1234)
    ins = [wl[k] for k in ("msg", "r", "s", "pkx", "pky")]
    plain, e1 = engine.tape_witness(engine.TAPE_ECDSA, ins)
    cmap, nn, nw = engine.tape_compact_layout(engine.TAPE_ECDSA)
    dins = [torch.from_numpy(a.view(np.int64)).cuda() for a in ins]
    arr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in dins])
    for with_wide in (True, False):
        nar = torch.zeros((nn, stride), dtype=torch.int32, device="cuda")
        wide = torch.zeros((nw, stride), dtype=torch.int64, device="cuda") if with_wide else None
        err = torch.zeros(n, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        rc = engine.lib.p2e_tape_witness_compact_dev(engine.TAPE_ECDSA, n, ctypes.cast(arr, ctypes.c_void_p), n, nar.data_ptr(),
                                                     wide.data_ptr() if with_wide else None, stride, err.data_ptr(), None)
        assert rc == 0, engine.last_error()
        engine.synchronize()
        nh = nar.cpu().numpy().view(np.uint32)[:, :n]
        wh = wide.cpu().numpy().view(np.uint64)[:, :n] if with_wide else None
        assert not err.cpu().numpy().any()
        assert np.array_equal(engine.expand_compact(cmap, np.ascontiguousarray(nh), None if wh is None else np.ascontiguousarray(wh)), plain)
        assert not nar.cpu().numpy()[:, n:].any()          # nothing is written past n


def test_glv_mul_full_size_sampled_columns(engine):
    """BASELINE config 4 at full size and across the scratch-chunk boundary: 2^16 + 96 glv_mul chains (chunks of 65 536) in the
    derived compact device format; sampled columns on both sides of the boundary against the oracle tape."""
    import torch
    n = (1 << 16) + 96
    tape = engine.TAPE_GLV_MUL
    ks = W.random_field_limbs(W.SEED + 7, n, 1)
    gx, gy = W._ec_mul_g(7)
    px = np.repeat(W.ints_to_limbs([gx]), n, axis=1)
    py = np.repeat(W.ints_to_limbs([gy]), n, axis=1)
    cmap, nn, nw = engine.tape_compact_layout(tape)
    dins = [torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda() for a in (ks, px, py)]
    arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in dins])
    nar = torch.empty((nn, n), dtype=torch.int32, device="cuda")
    err = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = engine.lib.p2e_tape_witness_compact_dev(tape, n, ctypes.cast(arr, ctypes.c_void_p), n, nar.data_ptr(), None, n, err.data_ptr(), None)
    assert rc == 0, engine.last_error()
    engine.synchronize()
    assert not bool(err.any())
    cols = [0, 1, 4097, 65535, 65536, 65537, n - 1]
    sub = np.ascontiguousarray(nar[:, cols].cpu().numpy().view(np.uint32))
    wit = engine.expand_compact(cmap, sub)
    for j, c in enumerate(cols):
        k = sum(int(ks[l, c]) << (29 * l) for l in range(9))
        wb = S.WitnessBuilder()
        wb.glv_mul(wb.constant_affine_point((gx, gy)), wb.constant_nonnative(k, S.SECP_N))
        assert [int(v) for v in wit[:, j]] == wb.flat(), f"column {c}"


def test_repeated_call_replays_a_cuda_graph_with_the_same_bits(engine):
    """A call repeated with the same tape, sizes, buffers and stream is captured on its second occurrence and replayed as
    one CUDA graph launch from the third on (option "chain_graph"): every occurrence must write the same witness."""
    import torch
    n = 96
    wl = W.ecdsa_workload(n, seed=123)
    ins = [torch.from_numpy(wl[k].view(np.int64)).cuda() for k in ("msg", "r", "s", "pkx", "pky")]
    _, nn, _ = engine.tape_compact_layout(engine.TAPE_ECDSA)
    nar = torch.zeros((nn, n), dtype=torch.int32, device="cuda")
    err = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    outs = []
    for rep in range(5):
        nar.zero_()
        err.fill_(7)
        torch.cuda.synchronize()
        rc = engine.lib.p2e_ecdsa_witness_compact_dev(n, *[t.data_ptr() for t in ins], n, nar.data_ptr(), None, n, err.data_ptr(), st.cuda_stream)
        assert rc == 0, engine.last_error()
        st.synchronize()
        assert not bool(err.any())
        outs.append(nar.cpu().numpy().view(np.uint32).copy())
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    old = engine.get_option("chain_graph")
    engine.set_option("chain_graph", 0)
    nar.zero_()
    assert engine.lib.p2e_ecdsa_witness_compact_dev(n, *[t.data_ptr() for t in ins], n, nar.data_ptr(), None, n, err.data_ptr(), st.cuda_stream) == 0
    st.synchronize()
    engine.set_option("chain_graph", old)
    assert np.array_equal(nar.cpu().numpy().view(np.uint32), outs[0])
    narrow, _, e2 = engine.tape_witness_compact(engine.TAPE_ECDSA, [wl[k] for k in ("msg", "r", "s", "pkx", "pky")], derive_check_sum=True)
    assert np.array_equal(narrow, outs[0]) and not e2.any()
