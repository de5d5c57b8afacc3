"""GPU tier: the PTX carry-chain bodies of csrc/p2e_f8.cuh (saturated 8 x 32-bit secp256k1 base field) against
Python integers, and word for word against the portable bodies the host / CPU tier runs (tests/hostcheck)."""
import ctypes
import random

import numpy as np
import pytest

from oracle import spec as S
from tests.test_hostcheck import F8_EDGES, P, _from_words, _words

pytestmark = pytest.mark.gpu
P_SECP = S.MODULI[0]


@pytest.mark.parametrize("op", range(9))
def test_f8_ptx_ops_match_integers_and_the_portable_bodies(engine, hostcheck, op):
    rng = random.Random(5000 + op)
    pairs = [(a, b) for a in F8_EDGES for b in F8_EDGES] + [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(20000)]
    a, b = _words([p[0] for p in pairs]), _words([p[1] for p in pairs])
    out = np.zeros_like(a)
    assert engine.lib.p2e_debug_f8(op, len(pairs), a.ctypes.data, b.ctypes.data, out.ctypes.data) == 0, engine.last_error()
    ref = np.zeros_like(a)
    hostcheck.hc_f8(op, ctypes.c_size_t(len(pairs)), P(a), P(b), P(ref))
    assert np.array_equal(out, ref), "PTX body differs from the portable body"
    got = _from_words(out)
    for (x, y), g in zip(pairs[:2000], got[:2000]):
        want = [x * y, x * x, x + y, x - y, 2 * x, 3 * x, 4 * x, 8 * x, x][op] % P_SECP
        assert g % P_SECP == want and g < (1 << 256)


def test_f8_inverse_division_steps_and_ladder(engine):
    """op 11: division steps on a weak value (what chain K2 runs); op 14: the Fermat ladder; op 15: division steps modulo
    the group order (what K0 runs for s^-1).  0 -> 0."""
    rng = random.Random(11)
    vals = [0, P_SECP, 1, 2, P_SECP - 1, (1 << 256) - 1, P_SECP + 5] + [rng.randrange(1, 1 << 256) for _ in range(3000)]
    a = _words(vals)
    out, lad = np.zeros_like(a), np.zeros_like(a)
    assert engine.lib.p2e_debug_f8(11, len(vals), a.ctypes.data, a.ctypes.data, out.ctypes.data) == 0
    assert engine.lib.p2e_debug_f8(14, len(vals), a.ctypes.data, a.ctypes.data, lad.ctypes.data) == 0
    for x, g, l in zip(vals, _from_words(out), _from_words(lad)):
        assert g == (pow(x, -1, P_SECP) if x % P_SECP else 0), hex(x)
        assert l % P_SECP == g
    N_SECP = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    vals = [0, 1, 2, N_SECP - 1, (N_SECP + 1) // 2] + [rng.randrange(N_SECP) for _ in range(3000)]
    a = _words(vals)
    out = np.zeros_like(a)
    assert engine.lib.p2e_debug_f8(15, len(vals), a.ctypes.data, a.ctypes.data, out.ctypes.data) == 0
    for x, g in zip(vals, _from_words(out)):
        assert g == (pow(x, -1, N_SECP) if x else 0), hex(x)


def test_f8_divmod16_ptx_body(engine):
    """(q, r) = divmod(a * b, p) on the device (PTX body) against Python integers; factors below p so that q < 2^256."""
    rng = random.Random(12)
    vals = [0, 1, P_SECP - 1, P_SECP - 2, (1 << 32) + 977, 1 << 255, (1 << 128) - 1]
    pairs = [(a, b) for a in vals for b in vals] + [(rng.randrange(P_SECP), rng.randrange(P_SECP)) for _ in range(20000)]
    a, b = _words([p[0] for p in pairs]), _words([p[1] for p in pairs])
    rem, quo = np.zeros_like(a), np.zeros_like(a)
    assert engine.lib.p2e_debug_f8(12, len(pairs), a.ctypes.data, b.ctypes.data, rem.ctypes.data) == 0
    assert engine.lib.p2e_debug_f8(13, len(pairs), a.ctypes.data, b.ctypes.data, quo.ctypes.data) == 0
    for (x, y), r, q in zip(pairs, _from_words(rem), _from_words(quo)):
        assert (q, r) == divmod(x * y, P_SECP)
