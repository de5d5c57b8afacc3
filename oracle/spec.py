"""CPU ORACLE (test infrastructure, NOT product code) -- Python big-int restatement of the
reference's hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this module.  The product path (plonky2_ecdsa_b200/) never does.

Every function cites the file:line of /root/reference (Weobe/plonky2-ecdsa) it restates.

Parity status
  * In-tree generators / gates: restated literally, quirks
    included.  Pinned against the reference's own in-tree vectors and identities by
    tests/test_oracle_kat.py (test_to_digits literal vector curve/curve_msm.rs:199-233,
    generator coordinates curve/secp256k1.rs:25-38, GLV constants + identities
    curve/glv.rs:11-32,60,122, the algebraic identities each generator's gadget enforces) and
    against independent implementations present in this image (OpenSSL BN_*, python
    `cryptography` SECP256K1 sign/verify).  The Rust reference itself cannot be built here
    (no cargo/rustc; git dependencies un-vendored, Cargo.toml:17-19).
  * Third-party pieces not under /root/reference, restated from their published algorithms:
      - plonky2 (git 0xPolygonZero/plonky2, NO rev pinned, Cargo.toml:17): GoldilocksField,
        Secp256K1Base/Scalar (from_noncanonical_biguint / to_canonical_biguint / inverse),
        KeccakHash::<32>::hash_no_pad.  Semantics follow the in-tree clone
        field/p256_base.rs:112-130,162-168 (same pattern).
      - num 0.4 BigUint / Ratio::round (Cargo.toml:16): exact integer semantics.
      - plonky2_ux (git Weobe/plonky2-ux, NO rev pinned, Cargo.toml:19): NOT restated here;
        "parity unpinned" (SURVEY.md section 8c).
"""
from __future__ import annotations

BITS = 29                      # gadgets/nonnative.rs:32
MASK29 = (1 << BITS) - 1
NUM_LIMBS = 9                  # gates/mul_nonnative.rs:37-39
GL_P = (1 << 64) - (1 << 32) + 1          # Goldilocks order (plonky2 GoldilocksField)
SECP_P = (1 << 256) - (1 << 32) - 977      # Secp256K1Base order
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141  # Secp256K1Scalar order

MOD_BASE = 0
MOD_SCALAR = 1
MODULI = {MOD_BASE: SECP_P, MOD_SCALAR: SECP_N}


class RefPanic(Exception):
    """The reference would panic / assert here (SURVEY.md section 8b 'Ownership / errors')."""


# ----------------------------------------------------------------------------- A1 convert_base
def convert_base(x, frm, to):
    """gadgets/biguint.rs:27-51."""
    res = []
    rem = 0
    offset = 0
    mask = (1 << to) - 1
    for i in x:
        rem += i << offset
        offset += frm
        while offset >= to:
            res.append(rem & mask)
            rem >>= to
            offset -= to
    if rem != 0:
        if rem >= (1 << 32):
            raise RefPanic("Value doesn't fit in u32")
        res.append(rem)
    for v in res:
        if v >= (1 << to):
            raise RefPanic("convert_base limb assert")
    return res


def to_u32_digits(v):
    """num BigUint::to_u32_digits: little-endian, zero -> empty vec."""
    out = []
    while v:
        out.append(v & 0xFFFFFFFF)
        v >>= 32
    return out


# ----------------------------------------------------------------------------- A2 get/set biguint
def get_biguint(limbs):
    """WitnessBigUint::get_biguint_target, gadgets/biguint.rs:444-452 (Horner, no range check)."""
    acc = 0
    for l in reversed(limbs):
        acc = (acc << BITS) + l
    return acc


def set_biguint(value, num_limbs):
    """GeneratedValuesBigUint::set_biguint_target, gadgets/biguint.rs:471-480."""
    limbs = convert_base(to_u32_digits(value), 32, BITS)
    if num_limbs < len(limbs):
        raise RefPanic("assert!(target.num_limbs() >= limbs.len())")
    return limbs + [0] * (num_limbs - len(limbs))


def ff_from_noncanonical(v):
    """[OOT plonky2] Secp256K1*::from_noncanonical_biguint: stores 4 x u64, panics above 2^256
    (pattern field/p256_base.rs:121-130)."""
    if v >= (1 << 256):
        raise RefPanic("from_noncanonical_biguint: value does not fit 4 x u64")
    return v


def ff_canon(v, m):
    """[OOT plonky2] to_canonical_biguint: ONE conditional subtract (field/p256_base.rs:162-168)."""
    return v - m if v >= m else v


def ff_inverse(v, m):
    """[OOT plonky2] try_inverse = Fermat exp_biguint(order-2); inverse() panics on zero
    (pattern field/p256_base.rs:112-119)."""
    v = ff_canon(v, m)
    if v == 0:
        raise RefPanic("Tried to invert zero")
    return pow(v, m - 2, m)


# ----------------------------------------------------------------------------- A3 div-rem
def biguint_divrem_gen(a_limbs, b_limbs):
    """BigUintDivRemGenerator::run_once gadgets/biguint.rs:508-518; output sizes :391-399."""
    a_len, b_len = len(a_limbs), len(b_limbs)
    div_n = 0 if b_len > a_len + 1 else a_len - b_len + 1
    a = get_biguint(a_limbs)
    b = get_biguint(b_limbs)
    if b == 0:
        raise RefPanic("division by zero")
    d, r = divmod(a, b)
    return set_biguint(d, div_n), set_biguint(r, b_len)


# ----------------------------------------------------------------------------- A4 addition
def nonnative_add_gen(a_limbs, b_limbs, m):
    """NonNativeAdditionGenerator::run_once gadgets/nonnative.rs:626-645 (strict `>` quirk :637)."""
    a = ff_canon(ff_from_noncanonical(get_biguint(a_limbs)), m)
    b = ff_canon(ff_from_noncanonical(get_biguint(b_limbs)), m)
    s = a + b
    if s > m:
        ov, s = 1, s - m
    else:
        ov = 0
    return set_biguint(s, NUM_LIMBS), ov


# ----------------------------------------------------------------------------- A5 multiple adds
def nonnative_add_many_gen(summands, m):
    """NonNativeMultipleAddsGenerator::run_once gadgets/nonnative.rs:696-728."""
    tot = 0
    for limbs in summands:
        tot += ff_canon(ff_from_noncanonical(get_biguint(limbs)), m)
    ov, s = divmod(tot, m)
    ov32 = ov & 0xFFFFFFFF          # to_u64_digits()[0] as u32  (:724)
    return set_biguint(s, NUM_LIMBS), ov32


# ----------------------------------------------------------------------------- A6 subtraction
def nonnative_sub_gen(a_limbs, b_limbs, m):
    """NonNativeSubtractionGenerator::run_once gadgets/nonnative.rs:792-810 (`>=` :803)."""
    a = ff_canon(ff_from_noncanonical(get_biguint(a_limbs)), m)
    b = ff_canon(ff_from_noncanonical(get_biguint(b_limbs)), m)
    if a >= b:
        d, ov = a - b, 0
    else:
        d, ov = m + a - b, 1
    return set_biguint(d, NUM_LIMBS), ov


# ----------------------------------------------------------------------------- A7 inverse
def nonnative_inv_gen(x_limbs, m):
    """NonNativeInverseGenerator::run_once gadgets/nonnative.rs:857-872.
    Outputs sized x.num_limbs() each (:507-509).  Returns (div, inv) in set_* call order."""
    n = len(x_limbs)
    x = ff_from_noncanonical(get_biguint(x_limbs))
    inv = ff_inverse(x, m)
    xc = ff_canon(x, m)
    div = (xc * inv) // m
    return set_biguint(div, n), set_biguint(inv, n)


# ----------------------------------------------------------------------------- A8 mul gate generator
def gl_canon_signed(v):
    """Canonical Goldilocks representative of a (possibly negative) integer."""
    return v % GL_P


def mul_nonnative_gen(x_wires, y_wires, m):
    """MulNonnativeGenerator::run_once gates/mul_nonnative.rs:249-324.
    x_wires / y_wires are 9 canonical Goldilocks values.  Returns (q[9], r[9], check_sum[17])."""
    assert len(x_wires) == NUM_LIMBS and len(y_wires) == NUM_LIMBS
    for w in list(x_wires) + list(y_wires):
        if w >= (1 << 32):
            raise RefPanic("Value doesn't fit in u32")          # :262,271
    for w in list(x_wires) + list(y_wires):
        if w >= (1 << BITS):
            raise RefPanic("assert limb < 2^29")                # :275-276
    m29 = convert_base(to_u32_digits(m), 32, BITS)
    x = get_biguint(x_wires)      # == BigUint::from_slice(convert_base(29->32))  :278-282
    y = get_biguint(y_wires)
    q, r = divmod(x * y, m)                                      # :284-285 (no pre-reduction)
    q29 = convert_base(to_u32_digits(q), 32, BITS)
    r29 = convert_base(to_u32_digits(r), 32, BITS)
    q29 += [0] * (NUM_LIMBS - len(q29))                          # :289-294 (a 10th limb is NOT written)
    r29 += [0] * (NUM_LIMBS - len(r29))
    m29 += [0] * (NUM_LIMBS - len(m29))
    cs = []
    for i in range(2 * NUM_LIMBS - 1):                           # :305-322
        lo = max(i - NUM_LIMBS + 1, 0)
        hi = min(i + 1, NUM_LIMBS)
        acc = 0
        for j in range(lo, hi):
            acc = (acc + q29[i - j] * m29[j] - x_wires[j] * y_wires[i - j]) % GL_P
        if i < NUM_LIMBS:
            acc = (acc + r29[i]) % GL_P
        cs.append(acc)
    return q29[:NUM_LIMBS], r29, cs


# ----------------------------------------------------------------------------- A9 check-sum generator
_GL_INV_2_29 = pow(1 << BITS, GL_P - 2, GL_P)


def check_sum_gen(a_wires):
    """CheckSumGenerator::run_once gates/mul_nonnative.rs:513-531 (Goldilocks field division)."""
    assert len(a_wires) == 2 * NUM_LIMBS - 1
    last = 0
    out = []
    for i in range(2 * NUM_LIMBS - 2):
        b = ((a_wires[i] + last) % GL_P) * _GL_INV_2_29 % GL_P
        o = (b + (1 << 33)) % GL_P
        out.append(o)
        last = b
        if o >= (1 << 34):
            raise RefPanic("assert!((b_i + offset) < 2^34)")     # :527
    return out


def check_sum_gen_nopanic(a_wires):
    """Same arithmetic as check_sum_gen but returns (out, panicked) so that batch parity tests can
    compare the values the reference would have produced up to the failing assert."""
    last = 0
    out = []
    bad = False
    for i in range(2 * NUM_LIMBS - 2):
        b = ((a_wires[i] + last) % GL_P) * _GL_INV_2_29 % GL_P
        o = (b + (1 << 33)) % GL_P
        out.append(o)
        last = b
        if o >= (1 << 34):
            bad = True
    return out, bad


# ----------------------------------------------------------------------------- A11 / A12 gate evaluation
def mul_nonnative_gate_eval(w, m):
    """MulNonnativeGate::eval_unfiltered gates/mul_nonnative.rs:101-130, base-field component.
    w: 53 canonical Goldilocks wires (x 0..8, y 9..17, r 18..26, q 27..35, check_sum 36..52)."""
    assert len(w) == 53
    m29 = convert_base(to_u32_digits(m), 32, BITS)
    assert len(m29) == NUM_LIMBS
    n = NUM_LIMBS
    res = []
    for i in range(2 * n - 1):
        lo = max(i - n + 1, 0)
        hi = min(i + 1, n)
        acc = 0
        for j in range(lo, hi):
            acc = (acc + m29[j] * w[3 * n + (i - j)] - w[j] * w[n + (i - j)]) % GL_P
        if i < n:
            acc = (acc + w[2 * n + i]) % GL_P
        res.append((acc - w[4 * n + i]) % GL_P)
    return res


def check_sum_gate_eval(w):
    """CheckSumGate::eval_unfiltered gates/mul_nonnative.rs:411-427.  w: 33 wires (a 0..16, b 17..32)."""
    assert len(w) == 33
    n = NUM_LIMBS
    base = 1 << 29
    off = 1 << 33
    last = 0
    res = []
    for i in range(2 * n - 1):
        if i < 2 * n - 2:
            o = (w[2 * n - 1 + i] - off) % GL_P
            res.append((w[i] + last - base * o) % GL_P)
            last = o
        else:
            res.append((w[i] + last) % GL_P)
    return res


# ----------------------------------------------------------------------------- native curve (constants only)
# curve/curve_types.rs, curve/curve_adds.rs, curve/secp256k1.rs: out-of-circuit arithmetic the
# reference uses to compute circuit CONSTANTS.  Affine with None == point at infinity.
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798   # curve/secp256k1.rs:25-31
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8   # curve/secp256k1.rs:33-38
G = (GX, GY)


def pt_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % SECP_P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, SECP_P - 2, SECP_P) % SECP_P
    else:
        lam = (y2 - y1) * pow(x2 - x1, SECP_P - 2, SECP_P) % SECP_P
    x3 = (lam * lam - x1 - x2) % SECP_P
    y3 = (lam * (x1 - x3) - y1) % SECP_P
    return (x3, y3)


def pt_neg(P):
    return None if P is None else (P[0], (-P[1]) % SECP_P)


def pt_mul(k, P):
    R = None
    Q = P
    while k:
        if k & 1:
            R = pt_add(R, Q)
        Q = pt_add(Q, Q)
        k >>= 1
    return R


def pt_is_valid(P):
    """AffinePoint::is_valid curve/curve_types.rs:62-65 (A = 0, B = 7  curve/secp256k1.rs:15-16)."""
    return P is None or (P[1] * P[1] - P[0] ** 3 - 7) % SECP_P == 0


def to_digits(x, w):
    """curve/curve_msm.rs:160-186 (native MSM digit split; holds the repo's only literal KAT)."""
    scalar_bits = 256
    num_digits = -(-scalar_bits // w)
    x = ff_canon(x, SECP_N)
    out = []
    for i in range(num_digits):
        lo = i * w
        hi = min((i + 1) * w, scalar_bits)
        out.append((x >> lo) & ((1 << (hi - lo)) - 1))
    return out


# ----------------------------------------------------------------------------- Keccak-256 (original padding)
_KECCAK_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_KECCAK_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61],
               [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & 0xFFFFFFFFFFFFFFFF if n else v


def keccak_f1600(A):
    for rc in _KECCAK_RC:
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _KECCAK_ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= rc
    return A


def keccak256(data: bytes) -> bytes:
    """Keccak-256 with the ORIGINAL 0x01 padding (not SHA3-256), as used by
    [OOT plonky2] KeccakHash::hash_no_pad (keccak_hash::keccak)."""
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        blk = msg[off:off + rate]
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = keccak_f1600(A)
    out = b""
    for i in range(4):
        out += A[i % 5][i // 5].to_bytes(8, "little")
    return out


def blinding_scalar():
    """gadgets/curve_msm.rs:33-36, gadgets/curve_fixed_base.rs:34-37:
    KeccakHash::<32>::hash_no_pad(&[F::ZERO]) -> BigUint::from_bytes_le -> from_noncanonical_biguint."""
    h = keccak256((0).to_bytes(8, "little"))
    return ff_canon(ff_from_noncanonical(int.from_bytes(h, "little")), SECP_N)


# ----------------------------------------------------------------------------- A10 GLV decomposition
# curve/glv.rs:11-32 (u64 little-endian limbs)
def _u64s(ls):
    return sum(v << (64 * i) for i, v in enumerate(ls))


GLV_BETA = _u64s([13923278643952681454, 11308619431505398165, 7954561588662645993, 8856726876819556112])
GLV_S = _u64s([16069571880186789234, 1310022930574435960, 11900229862571533402, 6008836872998760672])
GLV_A1 = _u64s([16747920425669159701, 3496713202691238861, 0, 0])
GLV_MINUS_B1 = _u64s([8022177200260244675, 16448129721693014056, 0, 0])
GLV_A2 = _u64s([6323353552219852760, 1498098850674701302, 1, 0])
GLV_B2 = _u64s([16747920425669159701, 3496713202691238861, 0, 0])


def _ratio_round(num, den):
    """num::rational::Ratio::round for a non-negative ratio: half away from zero."""
    return (2 * num + den) // (2 * den)


def decompose_secp256k1_scalar(k):
    """curve/glv.rs:39-77.  Returns (|k1|, |k2|, k1_neg, k2_neg)."""
    n = SECP_N
    k = ff_canon(k, n)
    c1 = ff_canon(ff_from_noncanonical(_ratio_round(GLV_B2 * k, n)), n)
    c2 = ff_canon(ff_from_noncanonical(_ratio_round(GLV_MINUS_B1 * k, n)), n)
    k1_raw = (k - c1 * GLV_A1 - c2 * GLV_A2) % n
    k2_raw = (c1 * GLV_MINUS_B1 - c2 * GLV_B2) % n
    k1_neg = k1_raw > n // 2
    k1 = n - k1_raw if k1_neg else k1_raw
    k2_neg = k2_raw > n // 2
    k2 = n - k2_raw if k2_neg else k2_raw
    return k1, k2, int(k1_neg), int(k2_neg)


def glv_decomposition_gen(k_limbs):
    """GLVDecompositionGenerator::run_once gadgets/glv.rs:128-142; k1,k2 have ceil(128/29)=5 limbs (:62-63)."""
    k = ff_from_noncanonical(get_biguint(k_limbs))
    k1, k2, n1, n2 = decompose_secp256k1_scalar(k)
    return set_biguint(k1, 5), set_biguint(k2, 5), n1, n2


# ----------------------------------------------------------------------------- A13 gadget chains -> tape
# Generator kinds recorded on the tape, in gadget call order.  The tuple layout of each kind
# (this is the "parity unit" of SURVEY.md section 8b):
#   ADD      : sum[9], overflow                      (10)
#   ADDMANY  : sum[9], overflow                      (10)
#   SUB      : diff[9], overflow                     (10)
#   INV      : div[n], inv[n]                        (2n, n = limbs of the input)
#   MUL      : r[9], q[9], check_sum[17]             (35, gate wires 18..52)
#   CHECKSUM : b[16]                                 (16, gate wires 17..32)
#   GLV      : k1[5], k2[5], k1_neg, k2_neg          (12)
K_ADD, K_ADDMANY, K_SUB, K_INV, K_MUL, K_CHECKSUM, K_GLV = range(7)


class NN:
    """NonNativeTarget with a concrete value: limbs as held on the wires (gadgets/nonnative.rs:34-38)."""
    __slots__ = ("limbs",)

    def __init__(self, limbs):
        self.limbs = list(limbs)

    @property
    def value(self):
        return get_biguint(self.limbs)


class WitnessBuilder:
    """Executes the gadget call graph on concrete values and records every in-tree generator's
    output tuple in call order.  Method names mirror the reference's CircuitBuilder extension traits."""

    def __init__(self):
        self.tape = []          # list of (kind, [values])
        self.counts = {}
        self.mul_ops = []       # (x limbs, y limbs) of every mul_nonnative, call order (nonnative.rs:453-455 range-checks them)

    def _emit(self, kind, vals):
        self.tape.append((kind, list(vals)))
        self.counts[kind] = self.counts.get(kind, 0) + 1

    def flat(self):
        out = []
        for _, v in self.tape:
            out.extend(v)
        return out

    # -- gadgets/nonnative.rs
    def constant_nonnative(self, v, m):
        """nonnative.rs:205-208 + biguint.rs:165-175: zero has 0 limbs."""
        return NN(convert_base(to_u32_digits(ff_canon(v, m)), 32, BITS))

    def add_nonnative(self, a, b, m):
        s, ov = nonnative_add_gen(a.limbs, b.limbs, m)      # nonnative.rs:245-276
        self._emit(K_ADD, s + [ov])
        return NN(s)

    def add_many_nonnative(self, xs, m):
        if len(xs) == 1:                                    # nonnative.rs:315-317
            return xs[0]
        s, ov = nonnative_add_many_gen([x.limbs for x in xs], m)
        self._emit(K_ADDMANY, s + [ov])
        return NN(s)

    def sub_nonnative(self, a, b, m):
        d, ov = nonnative_sub_gen(a.limbs, b.limbs, m)      # nonnative.rs:356-388
        self._emit(K_SUB, d + [ov])
        return NN(d)

    def mul_nonnative(self, x, y, m):
        """nonnative.rs:390-464: wires zero-padded to 9 (:406-424); MulNonnative row then CheckSum row."""
        xw = x.limbs + [0] * (NUM_LIMBS - len(x.limbs))
        yw = y.limbs + [0] * (NUM_LIMBS - len(y.limbs))
        self.mul_ops.append((list(x.limbs), list(y.limbs)))
        q, r, cs = mul_nonnative_gen(xw, yw, m)
        self._emit(K_MUL, r + q + cs)
        self._emit(K_CHECKSUM, check_sum_gen(cs))
        return NN(r)

    def neg_nonnative(self, x, m):
        return self.sub_nonnative(NN([]), x, m)             # nonnative.rs:491-500 (zero has 0 limbs)

    def inv_nonnative(self, x, m):
        div, inv = nonnative_inv_gen(x.limbs, m)            # nonnative.rs:502-536
        self._emit(K_INV, div + inv)
        return NN(inv)

    def mul_nonnative_by_bool(self, a, b):
        return NN([l * b for l in a.limbs])                 # nonnative.rs:278-295, biguint.rs:360-374

    def if_nonnative(self, b, x, y, m):
        t = self.mul_nonnative_by_bool(x, b)                # nonnative.rs:297-308
        f = self.mul_nonnative_by_bool(y, 1 - b)
        return self.add_nonnative(t, f, m)

    def nonnative_conditional_neg(self, x, b, m):
        neg = self.neg_nonnative(x, m)                      # nonnative.rs:584-596
        t = self.mul_nonnative_by_bool(neg, b)
        f = self.mul_nonnative_by_bool(x, 1 - b)
        return self.add_nonnative(t, f, m)

    # -- gadgets/curve.rs
    def curve_assert_valid(self, p):
        m = SECP_P                                          # curve.rs:123-135
        a = self.constant_nonnative(0, m)
        b = self.constant_nonnative(7, m)
        y2 = self.mul_nonnative(p[1], p[1], m)
        x2 = self.mul_nonnative(p[0], p[0], m)
        x3 = self.mul_nonnative(x2, p[0], m)
        ax = self.mul_nonnative(a, p[0], m)
        axb = self.add_nonnative(ax, b, m)
        rhs = self.add_nonnative(x3, axb, m)
        if y2.limbs != rhs.limbs:
            raise RefPanic("connect_nonnative(y_squared, rhs) unsatisfied")

    def curve_conditional_neg(self, p, b):
        return (p[0], self.nonnative_conditional_neg(p[1], b, SECP_P))   # curve.rs:149-158

    def curve_double(self, p):
        m = SECP_P                                          # curve.rs:160-185
        x, y = p
        dy = self.add_nonnative(y, y, m)
        idy = self.inv_nonnative(dy, m)
        xx = self.mul_nonnative(x, x, m)
        a = self.constant_nonnative(0, m)
        t = self.add_many_nonnative([xx, xx, xx, a], m)
        lam = self.mul_nonnative(t, idy, m)
        lam2 = self.mul_nonnative(lam, lam, m)
        xd = self.add_nonnative(x, x, m)
        x3 = self.sub_nonnative(lam2, xd, m)
        xdiff = self.sub_nonnative(x, x3, m)
        lx = self.mul_nonnative(lam, xdiff, m)
        y3 = self.sub_nonnative(lx, y, m)
        return (x3, y3)

    def curve_repeated_double(self, p, n):
        for _ in range(n):                                  # curve.rs:187-200
            p = self.curve_double(p)
        return p

    def curve_add(self, p1, p2):
        m = SECP_P                                          # curve.rs:202-223
        x1, y1 = p1
        x2, y2 = p2
        u = self.sub_nonnative(y2, y1, m)
        v = self.sub_nonnative(x2, x1, m)
        vi = self.inv_nonnative(v, m)
        s = self.mul_nonnative(u, vi, m)
        s2 = self.mul_nonnative(s, s, m)
        xs = self.add_nonnative(x2, x1, m)
        x3 = self.sub_nonnative(s2, xs, m)
        xd = self.sub_nonnative(x1, x3, m)
        pr = self.mul_nonnative(s, xd, m)
        y3 = self.sub_nonnative(pr, y1, m)
        return (x3, y3)

    def curve_conditional_add(self, p1, p2, b):
        m = SECP_P                                          # curve.rs:225-243
        s = self.curve_add(p1, p2)
        xt = self.mul_nonnative_by_bool(s[0], b)
        yt = self.mul_nonnative_by_bool(s[1], b)
        xf = self.mul_nonnative_by_bool(p1[0], 1 - b)
        yf = self.mul_nonnative_by_bool(p1[1], 1 - b)
        return (self.add_nonnative(xt, xf, m), self.add_nonnative(yt, yf, m))

    def constant_affine_point(self, P):
        return (self.constant_nonnative(P[0], SECP_P), self.constant_nonnative(P[1], SECP_P))

    # -- gadgets/split_nonnative.rs
    @staticmethod
    def split_bits(x):
        bits = []
        for l in x.limbs:                                   # split_le_base::<2>(limb, 29) per limb
            if l >= (1 << BITS):
                raise RefPanic("split_le_base: limb does not fit 29 bits")
            bits += [(l >> i) & 1 for i in range(BITS)]
        return bits

    def split_nonnative_to_4_bit_limbs(self, x):
        bits = self.split_bits(x)                           # split_nonnative.rs:25-50
        while len(bits) % 4:
            bits.append(0)
        return [bits[i] + 2 * bits[i + 1] + 4 * (bits[i + 2] + 2 * bits[i + 3]) for i in range(0, len(bits), 4)]

    def split_nonnative_to_2_bit_limbs(self, x):
        bits = self.split_bits(x)                           # split_nonnative.rs:52-72
        while len(bits) % 2:
            bits.append(0)
        return [bits[i] + 2 * bits[i + 1] for i in range(0, len(bits), 2)]

    # -- gadgets/curve_fixed_base.rs:18-66
    def fixed_base_curve_mul(self, base, scalar):
        limbs = self.split_nonnative_to_4_bit_limbs(scalar)
        n_scaled = len(scalar.limbs) * 8                    # scaled_base length (:24); zip truncates
        rando = pt_mul(blinding_scalar(), G)
        result = self.constant_affine_point(rando)
        for i, limb in enumerate(limbs[:n_scaled]):
            muls = fixed_base_window(base, i)
            should_add = int(limb != 0)
            r = self.constant_affine_point(muls[limb])      # random_access_curve_points
            result = self.curve_conditional_add(result, r, should_add)
        to_add = self.constant_affine_point(pt_neg(rando))
        return self.curve_add(result, to_add)

    # -- gadgets/curve_msm.rs:21-79
    def curve_msm(self, p, q, n, m_):
        limbs_n = self.split_nonnative_to_2_bit_limbs(n)
        limbs_m = self.split_nonnative_to_2_bit_limbs(m_)
        assert len(limbs_n) == len(limbs_m)
        num_limbs = len(limbs_n)
        rando = pt_mul(blinding_scalar(), G)
        rando_t = self.constant_affine_point(rando)
        neg_rando = self.constant_affine_point(pt_neg(rando))
        pre = [p] * 16
        cur_p, cur_q = rando_t, rando_t
        for i in range(4):
            pre[i] = cur_p
            pre[4 * i] = cur_q
            cur_p = self.curve_add(cur_p, p)
            cur_q = self.curve_add(cur_q, q)
        for i in range(1, 4):
            pre[i] = self.curve_add(pre[i], neg_rando)
            pre[4 * i] = self.curve_add(pre[4 * i], neg_rando)
        for i in range(1, 4):
            for j in range(1, 4):
                pre[i + 4 * j] = self.curve_add(pre[i], pre[4 * j])
        result = rando_t
        for ln, lm in reversed(list(zip(limbs_n, limbs_m))):
            result = self.curve_repeated_double(result, 2)
            index = 4 * lm + ln
            r = pre[index]                                  # random_access_curve_points
            result = self.curve_conditional_add(result, r, int(index != 0))
        spm = rando
        for _ in range(2 * num_limbs):
            spm = pt_add(spm, spm)
        to_add = self.constant_affine_point(pt_neg(spm))
        return self.curve_add(result, to_add)

    # -- gadgets/glv.rs
    def decompose_secp256k1_scalar(self, k):
        n = SECP_N                                          # glv.rs:53-85
        k1l, k2l, n1, n2 = glv_decomposition_gen(k.limbs)
        self._emit(K_GLV, k1l + k2l + [n1, n2])
        k1, k2 = NN(k1l), NN(k2l)
        k1_raw = self.nonnative_conditional_neg(k1, n1, n)
        k2_raw = self.nonnative_conditional_neg(k2, n2, n)
        s = self.constant_nonnative(GLV_S, n)
        sbk = self.mul_nonnative(s, k2_raw, n)
        sbk = self.add_nonnative(sbk, k1_raw, n)
        if get_biguint(sbk.limbs) != get_biguint(k.limbs):
            raise RefPanic("connect_nonnative(should_be_k, k) unsatisfied")
        return k1, k2, n1, n2

    def glv_mul(self, p, k):
        k1, k2, n1, n2 = self.decompose_secp256k1_scalar(k)  # glv.rs:87-104
        beta = self.constant_nonnative(GLV_BETA, SECP_P)
        bpx = self.mul_nonnative(beta, p[0], SECP_P)
        sp = (bpx, p[1])
        p_neg = self.curve_conditional_neg(p, n1)
        sp_neg = self.curve_conditional_neg(sp, n2)
        return self.curve_msm(p_neg, sp_neg, k1, k2)

    # -- gadgets/ecdsa.rs:30-53
    def verify_secp256k1_message(self, msg, r, s, pk):
        n = SECP_N
        self.curve_assert_valid(pk)
        c = self.inv_nonnative(s, n)
        u1 = self.mul_nonnative(msg, c, n)
        u2 = self.mul_nonnative(r, c, n)
        p1 = self.fixed_base_curve_mul(G, u1)
        p2 = self.glv_mul(pk, u2)
        pt = self.curve_add(p1, p2)
        if get_biguint(pt[0].limbs) != get_biguint(r.limbs):
            raise RefPanic("connect_nonnative(r, point.x) unsatisfied")
        return pt


_FIXED_TABLES = {}


def fixed_base_window(base, i):
    """Constant tables of fixed_base_curve_mul_circuit (curve_fixed_base.rs:24-52): window i holds
    t * 16^i * base for t = 1..15, with entry 0 duplicating entry 1 (:52)."""
    tabs = _FIXED_TABLES.get(base)
    if tabs is None:
        tabs = []
        point = base
        for _ in range(72):
            acc = None
            muls = []
            for _ in range(16):
                muls.append(acc)
                acc = pt_add(point, acc)
            muls = muls[1:]
            muls.insert(0, muls[0])
            tabs.append(muls)
            for _ in range(4):
                point = pt_add(point, point)
        _FIXED_TABLES[base] = tabs
    return tabs[i]


def limbs9(v):
    """Value -> 9 little-endian 29-bit limbs (witness wire contents of a full-width nonnative)."""
    return [(v >> (BITS * i)) & MASK29 for i in range(NUM_LIMBS)]


def ecdsa_witness(msg, r, s, pk):
    """Full tape for one verify_secp256k1_message_circuit instance (gadgets/ecdsa.rs:91-125: all
    inputs are constants produced by constant_nonnative / constant_affine_point)."""
    wb = WitnessBuilder()
    n = SECP_N
    wb.verify_secp256k1_message(wb.constant_nonnative(msg, n), wb.constant_nonnative(r, n),
                                wb.constant_nonnative(s, n), wb.constant_affine_point(pk))
    return wb


# ----------------------------------------------------------------------------- native ECDSA (workload generator)
def sign_message(msg, sk, k):
    """curve/ecdsa.rs:25-40 with the nonce k supplied by the caller (the reference draws it at random)."""
    rr = pt_mul(k, G)
    r = ff_canon(rr[0], SECP_N)                 # base_to_scalar, curve/curve_types.rs:280-282
    s = pow(k, SECP_N - 2, SECP_N) * (msg + r * sk) % SECP_N
    return r, s


def verify_message(msg, r, s, pk):
    """curve/ecdsa.rs:42-65."""
    c = pow(s, SECP_N - 2, SECP_N)
    u1 = msg * c % SECP_N
    u2 = r * c % SECP_N
    pt = pt_add(pt_mul(u1, G), pt_mul(u2, pk))
    return pt is not None and ff_canon(pt[0], SECP_N) == r


def splitmix64(state):
    """SplitMix64 step: returns (new_state, output).  Workload PRNG of SURVEY.md section 8d."""
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)
