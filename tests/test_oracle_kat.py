"""Pins the oracle (oracle/spec.py) against the reference's own in-tree vectors/identities and
against independent implementations in this image (OpenSSL BN_*, python `cryptography`)."""
import ctypes
import ctypes.util
import random

import pytest

from oracle import spec as S


# ---------------------------------------------------------------- reference literal vectors
def test_to_digits_literal_vector():
    """curve/curve_msm.rs:199-233 (the repo's only literal known-answer vector)."""
    x_canonical = [
        0b10101010101010101010101010101010, 0b10101010101010101010101010101010,
        0b11001100110011001100110011001100, 0b11001100110011001100110011001100,
        0b11110000111100001111000011110000, 0b11110000111100001111000011110000,
        0b00001111111111111111111111111111, 0b11111111111111111111111111111111]
    x = sum(v << (32 * i) for i, v in enumerate(x_canonical))
    assert S.to_u32_digits(S.ff_canon(x, S.SECP_N)) == x_canonical
    assert S.to_digits(x, 17) == [
        0b01010101010101010, 0b10101010101010101, 0b01010101010101010, 0b11001010101010101,
        0b01100110011001100, 0b00110011001100110, 0b10011001100110011, 0b11110000110011001,
        0b01111000011110000, 0b00111100001111000, 0b00011110000111100, 0b11111111111111110,
        0b01111111111111111, 0b11111111111111000, 0b11111111111111111, 0b1]


def test_generator_decimal_comments():
    """curve/secp256k1.rs:24,32 decimal comments next to the hex limbs."""
    assert S.GX == 55066263022277343669578718895168534326250603453777594175500187360389116729240
    assert S.GY == 32670510020758816978083085130507043184471273380659243275938904335757337482424
    assert S.pt_is_valid(S.G) and S.pt_is_valid(S.pt_neg(S.G))          # secp256k1.rs:49-61
    assert S.pt_mul(S.SECP_N, S.G) is None


def test_survey_limb_constants():
    """SURVEY.md section 0 'derived constants' (base-2^29 limbs of p and n)."""
    assert S.convert_base(S.to_u32_digits(S.SECP_P), 32, 29) == \
        [0x1ffffc2f, 0x1ffffff7] + [0x1fffffff] * 6 + [0x00ffffff]
    assert S.convert_base(S.to_u32_digits(S.SECP_N), 32, 29) == \
        [0x10364141, 0x1e92f466, 0x12280eef, 0x1db9cd5e, 0x1fffebaa, 0x1fffffff, 0x1fffffff, 0x1fffffff, 0x00ffffff]
    assert S.GL_P == 0xFFFFFFFF00000001


def test_glv_constants_identities():
    """curve/glv.rs:11-32 constants; identities asserted at :60 and in tests :114-141."""
    p, n = S.SECP_P, S.SECP_N
    assert pow(S.GLV_BETA, 3, p) == 1 and S.GLV_BETA != 1
    assert (S.GLV_S * S.GLV_S + S.GLV_S + 1) % n == 0
    assert (S.GLV_A1 - S.GLV_MINUS_B1 * S.GLV_S) % n == 0      # a1 + b1*s = 0
    assert (S.GLV_A2 + S.GLV_B2 * S.GLV_S) % n == 0
    # psi(P) = (beta x, y) == s * P
    P = S.pt_mul(123456789, S.G)
    assert S.pt_mul(S.GLV_S, P) == (S.GLV_BETA * P[0] % p, P[1])
    rng = random.Random(5)
    for _ in range(200):
        k = rng.randrange(n)
        k1, k2, n1, n2 = S.decompose_secp256k1_scalar(k)
        assert ((-k1 if n1 else k1) + S.GLV_S * (-k2 if n2 else k2)) % n == k     # glv.rs:60,122
        assert k1 < (1 << 129) and k2 < (1 << 129)


def test_fixed_scalar_mul_vectors():
    """secp256k1.rs:63-81 (10*G vs repeated add; fixed scalar [1111..8888] vs naive)."""
    acc = None
    for _ in range(10):
        acc = S.pt_add(acc, S.G)
    assert S.pt_mul(10, S.G) == acc
    lhs = sum(v << (32 * i) for i, v in enumerate([1111, 2222, 3333, 4444, 5555, 6666, 7777, 8888]))
    naive, g = None, S.G
    for j in range(256):
        if (lhs >> j) & 1:
            naive = S.pt_add(naive, g)
        g = S.pt_add(g, g)
    assert S.pt_mul(lhs, S.G) == naive


def test_convert_base_semantics():
    """gadgets/biguint.rs:27-51: zero -> no limbs; extra limb iff the remainder is non-zero."""
    assert S.convert_base([], 32, 29) == []
    assert S.convert_base([1], 32, 29) == [1]
    assert S.convert_base([0xFFFFFFFF], 32, 29) == [0x1FFFFFFF, 7]
    rng = random.Random(3)
    for _ in range(200):
        v = rng.getrandbits(rng.randrange(1, 300))
        l29 = S.convert_base(S.to_u32_digits(v), 32, 29)
        assert S.get_biguint(l29) == v
        back = S.convert_base(l29, 29, 32)
        assert sum(d << (32 * i) for i, d in enumerate(back)) == v


# ---------------------------------------------------------------- Keccak-256 (original padding) KATs
def test_keccak256_known_answers():
    assert S.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert S.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # multi-block input (> 136 bytes): cross-check with the sponge on a repeated pattern
    assert len(S.keccak256(b"\x00" * 200)) == 32
    assert S.blinding_scalar() < S.SECP_N


# ---------------------------------------------------------------- independent checkers
def _libcrypto():
    name = ctypes.util.find_library("crypto")
    if not name:
        pytest.skip("libcrypto not found")
    lib = ctypes.CDLL(name)
    lib.BN_new.restype = ctypes.c_void_p
    lib.BN_CTX_new.restype = ctypes.c_void_p
    lib.BN_bin2bn.restype = ctypes.c_void_p
    lib.BN_bin2bn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p]
    lib.BN_mod_mul.argtypes = [ctypes.c_void_p] * 5
    lib.BN_mod_inverse.restype = ctypes.c_void_p
    lib.BN_mod_inverse.argtypes = [ctypes.c_void_p] * 4
    lib.BN_div.argtypes = [ctypes.c_void_p] * 5
    lib.BN_mul.argtypes = [ctypes.c_void_p] * 4
    lib.BN_bn2bin.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.BN_num_bits.argtypes = [ctypes.c_void_p]
    lib.BN_free.argtypes = [ctypes.c_void_p]
    return lib


def _bn(lib, v):
    b = v.to_bytes((v.bit_length() + 7) // 8 or 1, "big")
    return lib.BN_bin2bn(b, len(b), None)


def _bn_int(lib, bn):
    nb = (lib.BN_num_bits(bn) + 7) // 8
    buf = ctypes.create_string_buffer(nb or 1)
    lib.BN_bn2bin(bn, buf)
    return int.from_bytes(buf.raw[:nb], "big")


def test_mul_and_inverse_generators_against_openssl():
    """A7/A8 against OpenSSL BN_mul/BN_div/BN_mod_inverse (independent of any of our code)."""
    lib = _libcrypto()
    ctx = lib.BN_CTX_new()
    rng = random.Random(11)
    for m in (S.SECP_P, S.SECP_N):
        bm = _bn(lib, m)
        for _ in range(60):
            x, y = rng.getrandbits(261), rng.getrandbits(261)
            bx, by = _bn(lib, x), _bn(lib, y)
            prod, q, r = lib.BN_new(), lib.BN_new(), lib.BN_new()
            assert lib.BN_mul(prod, bx, by, ctx) == 1
            assert lib.BN_div(q, r, prod, bm, ctx) == 1
            qq, rr, _ = S.mul_nonnative_gen(S.limbs9(x), S.limbs9(y), m)
            assert S.get_biguint(rr) == _bn_int(lib, r)
            assert S.get_biguint(qq) == _bn_int(lib, q) % (1 << 261)
            v = rng.randrange(1, m)
            inv = lib.BN_mod_inverse(None, _bn(lib, v), bm, ctx)
            d, iv = S.nonnative_inv_gen(S.limbs9(v), m)
            assert S.get_biguint(iv) == _bn_int(lib, inv)
            assert S.get_biguint(d) == (v * S.get_biguint(iv) - 1) // m


def test_sign_message_verifies_under_cryptography():
    """curve/ecdsa.rs:25-40 restatement vs OpenSSL ECDSA verify (SECP256K1, prehashed digest)."""
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import utils
    rng = random.Random(21)
    for _ in range(8):
        sk = rng.randrange(1, S.SECP_N)
        msg = rng.randrange(S.SECP_N)
        k = rng.randrange(1, S.SECP_N)
        r, s = S.sign_message(msg, sk, k)
        pk = S.pt_mul(sk, S.G)
        pn = ec.derive_private_key(sk, ec.SECP256K1()).public_key().public_numbers()
        assert (pn.x, pn.y) == pk
        pub = ec.EllipticCurvePublicNumbers(pk[0], pk[1], ec.SECP256K1()).public_key()
        pub.verify(utils.encode_dss_signature(r, s), msg.to_bytes(32, "big"),
                   ec.ECDSA(utils.Prehashed(hashes.SHA256())))
        assert S.verify_message(msg, r, s, pk)


# ---------------------------------------------------------------- identities each gadget enforces
@pytest.mark.parametrize("m", [S.SECP_P, S.SECP_N])
def test_generator_identities(m):
    rng = random.Random(31)
    n9 = S.NUM_LIMBS
    for _ in range(300):
        x, y = rng.randrange(m), rng.randrange(m)
        q, r, cs = S.mul_nonnative_gen(S.limbs9(x), S.limbs9(y), m)
        assert x * y == S.get_biguint(q) * m + S.get_biguint(r) and S.get_biguint(r) < m   # mul_nonnative.rs:72-76
        b = S.check_sum_gen(cs)
        # CheckSumGate constraints all vanish (mul_nonnative.rs:364-374, 411-427)
        assert S.check_sum_gate_eval(cs + b) == [0] * 17
        # MulNonnativeGate constraints all vanish on the generated row
        assert S.mul_nonnative_gate_eval(S.limbs9(x) + S.limbs9(y) + r + q + cs, m) == [0] * 17
        for c in b:
            assert c < (1 << 34)
        s, ov = S.nonnative_add_gen(S.limbs9(x), S.limbs9(y), m)
        assert x + y == S.get_biguint(s) + ov * m                                            # nonnative.rs:262-267
        d, ov = S.nonnative_sub_gen(S.limbs9(x), S.limbs9(y), m)
        assert S.get_biguint(d) + y - ov * m == x                                            # nonnative.rs:376-380
        if x:
            dv, iv = S.nonnative_inv_gen(S.limbs9(x), m)
            assert x * S.get_biguint(iv) == S.get_biguint(dv) * m + 1                        # nonnative.rs:518-524
        z = rng.randrange(m)
        sm, ov = S.nonnative_add_many_gen([S.limbs9(x), S.limbs9(y), S.limbs9(z), []], m)
        assert x + y + z == S.get_biguint(sm) + ov * m                                       # nonnative.rs:333-343
    assert len(S.limbs9(0)) == n9


def test_quirks():
    m = S.SECP_P
    # A4: strict '>' -- a + b == M leaves sum = M, overflow = 0 (nonnative.rs:637)
    s, ov = S.nonnative_add_gen(S.limbs9(m - 5), S.limbs9(5), m)
    assert S.get_biguint(s) == m and ov == 0
    # A6: '>=' (nonnative.rs:803)
    d, ov = S.nonnative_sub_gen(S.limbs9(7), S.limbs9(7), m)
    assert S.get_biguint(d) == 0 and ov == 0
    # non-canonical inputs below 2^256 are canonicalised by ONE subtraction, above panic
    s, ov = S.nonnative_add_gen(S.limbs9(m + 1), S.limbs9(1), m)
    assert S.get_biguint(s) == 2 and ov == 0
    with pytest.raises(S.RefPanic):
        S.nonnative_add_gen(S.limbs9(1 << 256), S.limbs9(1), m)
    with pytest.raises(S.RefPanic):
        S.nonnative_inv_gen(S.limbs9(0), m)
    with pytest.raises(S.RefPanic):
        S.nonnative_inv_gen(S.limbs9(m), m)
    # A8: limb checks
    with pytest.raises(S.RefPanic):
        S.mul_nonnative_gen([1 << 29] + [0] * 8, [1] + [0] * 8, m)
    with pytest.raises(S.RefPanic):
        S.mul_nonnative_gen([1 << 32] + [0] * 8, [1] + [0] * 8, m)
    # A8: q's 10th limb is dropped for unreduced inputs, the carry assert then fires in A9
    x = (1 << 261) - 1
    q, r, cs = S.mul_nonnative_gen(S.limbs9(x), S.limbs9(x), m)
    assert S.get_biguint(q) == (x * x // m) % (1 << 261) and S.get_biguint(r) == x * x % m
    with pytest.raises(S.RefPanic):
        S.check_sum_gen(cs)
    # A3: output sizes and division by zero
    d, rm = S.biguint_divrem_gen(S.limbs9(12345678901234567890123), [7, 1])
    assert len(d) == 8 and len(rm) == 2
    with pytest.raises(S.RefPanic):
        S.biguint_divrem_gen([1, 2, 3], [0, 0])


def test_ecdsa_tape_matches_survey_counts():
    """SURVEY.md section 3.1 per-circuit op counts, replayed through the gadget call graph."""
    sk, msg, k = 0x1234567 << 200 | 99, 0xabcdef << 180 | 5, 0x777 << 222 | 1
    pk = S.pt_mul(sk, S.G)
    r, s = S.sign_message(msg, sk, k)
    wb = S.ecdsa_witness(msg, r, s, pk)
    c = wb.counts
    assert c[S.K_MUL] == 1087 and c[S.K_CHECKSUM] == 1087
    assert c[S.K_INV] == 312 and c[S.K_SUB] == 1267 and c[S.K_ADD] == 742
    assert c[S.K_ADDMANY] == 146 and c[S.K_GLV] == 1
    assert len(wb.flat()) == 1087 * 51 + 312 * 18 + (1267 + 742 + 146) * 10 + 12
    # tampered signature -> the circuit's final connect cannot hold
    with pytest.raises(S.RefPanic):
        S.ecdsa_witness(msg, r, (s + 1) % S.SECP_N, pk)
