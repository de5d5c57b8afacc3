"""Seeded synthetic workloads for the BASELINE.json configs.

Host-side input generation only (numpy + Python ints); nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np

BITS = 29
MASK29 = (1 << BITS) - 1
SECP_P = (1 << 256) - (1 << 32) - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GL_P = (1 << 64) - (1 << 32) + 1
SEED = 0x5EC9256B1
MODULI = {0: SECP_P, 1: SECP_N}


def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """count outputs of SplitMix64 started at `seed` (vectorised: state_i = seed + (i+1)*gamma)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def ints_to_limbs(vals) -> np.ndarray:
    """list of Python ints (< 2^261) -> [9][n] uint64 wire-major limb matrix."""
    n = len(vals)
    out = np.empty((9, n), dtype=np.uint64)
    for i, v in enumerate(vals):
        for l in range(9):
            out[l, i] = (v >> (BITS * l)) & MASK29
    return out


def limbs_to_ints(m: np.ndarray):
    rows, n = m.shape
    return [sum(int(m[l, i]) << (BITS * l) for l in range(rows)) for i in range(n)]


def random_limbs_261(seed: int, n: int) -> np.ndarray:
    """[9][n]: every limb uniform < 2^29 (an 'unreduced' 261-bit operand, legal for MulNonnativeGenerator)."""
    r = splitmix64_stream(seed, 9 * n).reshape(9, n)
    return r & np.uint64(MASK29)


def random_field_limbs(seed: int, n: int, modulus: int) -> np.ndarray:
    """[9][n] limbs of values uniform in [0, modulus) by rejection from 256-bit SplitMix64 draws."""
    m = MODULI[modulus] if modulus in MODULI else modulus
    w = splitmix64_stream(seed, 4 * n + 64).reshape(-1, 4)
    # vectorised 256-bit -> limbs
    out = np.empty((9, n), dtype=np.uint64)
    vals = []
    k = 0
    j = 0
    # the secp256k1 moduli are within 2^-128 of 2^256: rejection is (practically) never taken,
    # handle it exactly anyway on the python-int slow path for the rare rows.
    top_ok = None
    lo = w[:, 0]; b1 = w[:, 1]; b2 = w[:, 2]; hi = w[:, 3]
    # rows whose top 64 bits are all ones might exceed m: send those through the exact path
    suspicious = np.nonzero(hi == np.uint64(0xFFFFFFFFFFFFFFFF))[0]
    bad = set()
    for s in suspicious:
        v = int(lo[s]) | int(b1[s]) << 64 | int(b2[s]) << 128 | int(hi[s]) << 192
        if v >= m:
            bad.add(int(s))
    keep = np.array([i for i in range(w.shape[0]) if i not in bad][:n], dtype=np.int64) if bad else np.arange(n)
    lo, b1, b2, hi = lo[keep], b1[keep], b2[keep], hi[keep]
    words = [lo, b1, b2, hi]
    for l in range(9):
        bit = BITS * l
        wi, sh = divmod(bit, 64)
        v = words[wi] >> np.uint64(sh)
        if sh + BITS > 64 and wi + 1 < 4:
            v = v | (words[wi + 1] << np.uint64(64 - sh))
        out[l] = v & np.uint64(MASK29)
    return out


def mulmod_workload(n: int, modulus: int = 0, seed: int = SEED):
    """Config 2: 75 % reduced operands, 25 % unreduced 261-bit operands, plus the fixed edge block
    {0, 1, M-1, M, 2^256-1, all-limbs-0x1FFFFFFF, x == y} at the front."""
    m = MODULI[modulus]
    x = random_field_limbs(seed, n, modulus)
    y = random_field_limbs(seed + 1, n, modulus)
    q = n // 4
    if q:
        x[:, -q:] = random_limbs_261(seed + 2, q)
        y[:, -q:] = random_limbs_261(seed + 3, q)
    edge = [0, 1, m - 1, m, (1 << 256) - 1, (1 << 261) - 1]
    pairs = [(a, b) for a in edge for b in edge]
    pairs = pairs[: min(len(pairs), n)]
    if pairs:
        x[:, : len(pairs)] = ints_to_limbs([p[0] for p in pairs])
        y[:, : len(pairs)] = ints_to_limbs([p[1] for p in pairs])
    return x, y


def random_goldilocks(seed: int, rows: int, n: int) -> np.ndarray:
    """[rows][n] uniform canonical Goldilocks elements (config 5 wire matrix)."""
    r = splitmix64_stream(seed, rows * n).reshape(rows, n)
    # fold the 2^32 - 1 non-canonical values back (bias 2^-32, irrelevant for a synthetic LDE)
    return np.where(r >= np.uint64(GL_P), r - np.uint64(GL_P), r)


# ------------------------------------------------------------------ ECDSA signatures (config 3)
def _ec_mul_g(k: int):
    """k*G on secp256k1 -> (x, y).  Uses OpenSSL through `cryptography` when importable, else a
    small Jacobian double-and-add in Python ints."""
    try:
        from cryptography.hazmat.primitives.asymmetric import ec
        pn = ec.derive_private_key(k, ec.SECP256K1()).public_key().public_numbers()
        return pn.x, pn.y
    except ImportError:
        pass
    p = SECP_P
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    X, Y, Z = 0, 1, 0
    for bit in bin(k)[2:]:
        if Z:
            S = 4 * X * Y * Y % p; M = 3 * X * X % p
            X2 = (M * M - 2 * S) % p
            Y, Z = (M * (S - X2) - 8 * pow(Y, 4, p)) % p, 2 * Y * Z % p
            X = X2
        if bit == "1":
            if not Z:
                X, Y, Z = gx, gy, 1
            else:
                Z2 = Z * Z % p
                U2 = gx * Z2 % p; S2 = gy * Z2 * Z % p
                H = (U2 - X) % p; R = (S2 - Y) % p
                H2 = H * H % p; H3 = H2 * H % p
                X3 = (R * R - H3 - 2 * X * H2) % p
                Y = (R * (X * H2 - X3) - Y * H3) % p
                Z = Z * H % p
                X = X3
    zi = pow(Z, p - 2, p)
    return X * zi * zi % p, Y * zi * zi * zi % p


def ecdsa_workload(n: int, seed: int = SEED):
    """n valid secp256k1 signatures: sk, msg, nonce from SplitMix64(seed); pk = sk*G,
    (r, s) as curve/ecdsa.rs:25-40 sign_message.  Returns dict of [9][n] limb matrices + ints."""
    w = splitmix64_stream(seed ^ 0xEC0DA, 12 * n).reshape(n, 12)
    msgs, rs, ss, pkx, pky = [], [], [], [], []
    for i in range(n):
        def take(o):
            return (int(w[i, o]) | int(w[i, o + 1]) << 64 | int(w[i, o + 2]) << 128 | int(w[i, o + 3]) << 192)
        sk = take(0) % (SECP_N - 1) + 1
        msg = take(4) % SECP_N
        k = take(8) % (SECP_N - 1) + 1
        px, py = _ec_mul_g(sk)
        rx, _ = _ec_mul_g(k)
        r = rx % SECP_N
        s = pow(k, SECP_N - 2, SECP_N) * (msg + r * sk) % SECP_N
        assert r != 0 and s != 0
        msgs.append(msg); rs.append(r); ss.append(s); pkx.append(px); pky.append(py)
    return {
        "msg": ints_to_limbs(msgs), "r": ints_to_limbs(rs), "s": ints_to_limbs(ss),
        "pkx": ints_to_limbs(pkx), "pky": ints_to_limbs(pky),
        "ints": {"msg": msgs, "r": rs, "s": ss, "pkx": pkx, "pky": pky},
    }
