"""TEST INFRASTRUCTURE ONLY (checker).  Goldilocks FFT / coset LDE as plonky2 defines them.

plonky2's field crate is not under /root/reference (git dependency, Cargo.toml:17), so its published
conventions are restated: `fft` maps coefficients to values[i] = P(omega_n^i) in natural order with
omega_n = F::primitive_root_of_unity(log n) = POWER_OF_TWO_GENERATOR^(2^(32 - log n));
`PolynomialValues::lde_onto_coset(rate_bits)` = ifft -> zero-pad -> coset_fft(F::coset_shift()).
The two constants are pinned by their defining identities (tests/test_ntt.py): the multiplicative
generator 7 has full order p - 1 and POWER_OF_TWO_GENERATOR = 7^((p - 1) / 2^32).
"""
GL_P = 0xFFFFFFFF00000001
MULTIPLICATIVE_GROUP_GENERATOR = 7
POWER_OF_TWO_GENERATOR = 1753635133440165772
TWO_ADICITY = 32


def primitive_root_of_unity(n_log):
    assert n_log <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - n_log), GL_P)


def dft_naive(x, root):
    """O(n^2) definition: X[k] = sum_i x[i] root^(i k)."""
    n = len(x)
    return [sum(x[i] * pow(root, i * k, GL_P) for i in range(n)) % GL_P for k in range(n)]


def _fft_rec(x, root):
    n = len(x)
    if n == 1:
        return list(x)
    even = _fft_rec(x[0::2], root * root % GL_P)
    odd = _fft_rec(x[1::2], root * root % GL_P)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * odd[k] % GL_P
        out[k] = (even[k] + t) % GL_P
        out[k + n // 2] = (even[k] - t) % GL_P
        w = w * root % GL_P
    return out


def fft(coeffs):
    n = len(coeffs)
    return _fft_rec(coeffs, primitive_root_of_unity(n.bit_length() - 1))


def ifft(values):
    n = len(values)
    inv_root = pow(primitive_root_of_unity(n.bit_length() - 1), GL_P - 2, GL_P)
    n_inv = pow(n, GL_P - 2, GL_P)
    return [v * n_inv % GL_P for v in _fft_rec(values, inv_root)]


def lde_onto_coset(values, rate_bits, shift=MULTIPLICATIVE_GROUP_GENERATOR):
    coeffs = ifft(values)
    padded = coeffs + [0] * (len(coeffs) * ((1 << rate_bits) - 1))
    s = 1
    for i in range(len(coeffs)):
        padded[i] = padded[i] * s % GL_P
        s = s * shift % GL_P
    return fft(padded)


def eval_poly(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % GL_P
    return acc
