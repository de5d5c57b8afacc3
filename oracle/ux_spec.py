"""TEST INFRASTRUCTURE ONLY (checker).  Restated spec of the `plonky2_ux` gates the reference calls.

PARITY UNPINNED.  `plonky2_ux` (git https://github.com/Weobe/plonky2-ux.git, no rev pinned,
reference Cargo.toml:19) is NOT under /root/reference and no reference test pins a wire index, an
aux-limb width or a constraint value of any of its gates.  What follows restates the
published upstream crate it generalises -- `plonky2_u32` (0xPolygonZero/plonky2-u32: gates
range_check_u32.rs, comparison.rs, add_many_u32.rs, subtraction_u32.rs, arithmetic_u32.rs and their
generators) -- with the limb width `bits` (32 upstream, 29 / 34 here) and the aux-limb width as
PARAMETERS.  The reference's call sites these gates serve:
  range_check_ux_circuit   gadgets/nonnative.rs:330-331, 373, 454-460 (x, y, r, q: 9 x 29 bits; b: 8 + 8 x 34 bits)
  list_le_ux_circuit       gadgets/biguint.rs:222-229 (cmp_biguint)
  add_many_ux              gadgets/biguint.rs:261 (add_biguint: carry + a_limb + b_limb)
  sub_ux                   gadgets/biguint.rs:284 (sub_biguint)
  mul_ux                   gadgets/biguint.rs:305, 335 (mul_biguint / square)
Structural hint used for the defaults: 9 x 29-bit limbs and 8 x 34-bit limbs each fit ONE
range-check row of a 135/136-wire config (nonnative.rs:454-460), which upstream's 2-bit aux limbs do
not (9 * (1 + 15) = 144 wires) => default aux width 3 bits (9 * (1 + 10) = 99, 8 * (1 + 12) = 104).
If upstream plonky2_ux differs, only the parameters / wire maps below change; the kernels take them
as arguments.

All values are canonical Goldilocks integers; a row is a list of `num_wires` ints.
"""
from __future__ import annotations

GL_P = 0xFFFFFFFF00000001


def _ceil_div(a, b):
    return -(-a // b)


def _reduce_with_powers(vals, base):
    """plonky2 `reduce_with_powers`: sum vals[j] * base^j (Horner from the top)."""
    acc = 0
    for v in reversed(vals):
        acc = (acc * base + v) % GL_P
    return acc


def _range_product(v, size):
    """prod_{x < size} (v - x): vanishes iff v in [0, size)."""
    acc = 1
    for x in range(size):
        acc = acc * ((v - x) % GL_P) % GL_P
    return acc


def _split(v, limb_bits, n):
    """generator-side unfold: n little-endian limbs of limb_bits bits (higher bits dropped)."""
    out = []
    for _ in range(n):
        out.append(v & ((1 << limb_bits) - 1))
        v >>= limb_bits
    return out


# ---------------------------------------------------------------------------------------------
class UXRangeCheckGate:
    """plonky2_u32 U32RangeCheckGate (range_check_u32.rs) with `bits`-wide input limbs.

    wires:  input limb i                        -> i
            aux limb j of input limb i          -> num_input_limbs + aux_limbs_per_input_limb * i + j
    constraints, for each input limb i in order:
            reduce_with_powers(aux_i, 2^aux_bits) - input_i ;  then prod_{x<2^aux_bits}(aux_ij - x) for each j
    """

    def __init__(self, bits=29, aux_bits=3, num_input_limbs=9):
        self.bits, self.aux_bits, self.num_input_limbs = bits, aux_bits, num_input_limbs
        self.aux_limbs_per_input_limb = _ceil_div(bits, aux_bits)

    @property
    def num_wires(self):
        return self.num_input_limbs * (1 + self.aux_limbs_per_input_limb)

    num_constraints = num_wires

    @property
    def degree(self):
        return 1 << self.aux_bits

    def wire_ith_input_limb(self, i):
        return i

    def wire_ith_input_limb_jth_aux_limb(self, i, j):
        return self.num_input_limbs + self.aux_limbs_per_input_limb * i + j

    def eval_unfiltered(self, w):
        base = 1 << self.aux_bits
        out = []
        for i in range(self.num_input_limbs):
            aux = [w[self.wire_ith_input_limb_jth_aux_limb(i, j)] for j in range(self.aux_limbs_per_input_limb)]
            out.append((_reduce_with_powers(aux, base) - w[i]) % GL_P)
            out.extend(_range_product(a, base) for a in aux)
        return out

    def generate(self, limbs):
        """U32RangeCheckGenerator::run_once: aux limbs of each input limb.  Returns (row, err):
        err = 1 if a limb does not fit aux_limbs * aux_bits bits (the row is then unsatisfiable)."""
        row = [0] * self.num_wires
        err = 0
        for i, v in enumerate(limbs):
            row[i] = v
            aux = _split(v, self.aux_bits, self.aux_limbs_per_input_limb)
            if v >> (self.aux_bits * self.aux_limbs_per_input_limb):
                err = 1
            for j, a in enumerate(aux):
                row[self.wire_ith_input_limb_jth_aux_limb(i, j)] = a
        return row, err


# ---------------------------------------------------------------------------------------------
class UXComparisonGate:
    """plonky2_u32 ComparisonGate (comparison.rs): result_bool = (first <= second) for num_bits-bit inputs.

    wires: first 0, second 1, result_bool 2, most_significant_diff 3, first chunks 4.., second chunks
    4+nc.., equality dummies 4+2nc.., chunks_equal 4+3nc.., intermediate values 4+4nc.., msd bits 4+5nc..
    (chunk_bits + 1 of them).  list_le_circuit uses chunk_bits = 2."""

    def __init__(self, num_bits=29, num_chunks=None, chunk_bits=2):
        self.num_bits = num_bits
        self.num_chunks = num_chunks if num_chunks is not None else _ceil_div(num_bits, chunk_bits)
        self.chunk_bits = _ceil_div(num_bits, self.num_chunks)

    @property
    def num_wires(self):
        return 4 + 5 * self.num_chunks + self.chunk_bits + 1

    @property
    def num_constraints(self):
        return 6 + 5 * self.num_chunks + self.chunk_bits

    @property
    def degree(self):
        return 1 << self.chunk_bits

    def eval_unfiltered(self, w):
        nc, cb = self.num_chunks, self.chunk_bits
        size = 1 << cb
        first = [w[4 + c] for c in range(nc)]
        second = [w[4 + nc + c] for c in range(nc)]
        out = [(_reduce_with_powers(first, size) - w[0]) % GL_P, (_reduce_with_powers(second, size) - w[1]) % GL_P]
        msd_so_far = 0
        for i in range(nc):
            out.append(_range_product(first[i], size))
            out.append(_range_product(second[i], size))
            diff = (second[i] - first[i]) % GL_P
            dummy, eq, inter = w[4 + 2 * nc + i], w[4 + 3 * nc + i], w[4 + 4 * nc + i]
            out.append((diff * dummy - (1 - eq)) % GL_P)
            out.append(eq * diff % GL_P)
            out.append((inter - eq * msd_so_far) % GL_P)
            msd_so_far = (inter + (1 - eq) * diff) % GL_P
        msd = w[3]
        out.append((msd - msd_so_far) % GL_P)
        bits = [w[4 + 5 * nc + b] for b in range(cb + 1)]
        out.extend(b * (1 - b) % GL_P for b in bits)
        out.append((size + msd - _reduce_with_powers(bits, 2)) % GL_P)
        out.append((w[2] - bits[cb]) % GL_P)
        return out

    def generate(self, first, second):
        """ComparisonGenerator::run_once."""
        nc, cb = self.num_chunks, self.chunk_bits
        size = 1 << cb
        row = [0] * self.num_wires
        row[0], row[1] = first, second
        fc, sc = _split(first, cb, nc), _split(second, cb, nc)
        msd = 0
        for i in range(nc):
            row[4 + i], row[4 + nc + i] = fc[i], sc[i]
            if fc[i] == sc[i]:
                row[4 + 2 * nc + i] = 1                               # equality dummy
                row[4 + 3 * nc + i] = 1                               # chunks_equal
                row[4 + 4 * nc + i] = msd                             # intermediate value
            else:
                row[4 + 2 * nc + i] = pow((sc[i] - fc[i]) % GL_P, GL_P - 2, GL_P)
                row[4 + 3 * nc + i] = 0
                row[4 + 4 * nc + i] = 0
                msd = (sc[i] - fc[i]) % GL_P
        row[3] = msd
        t = (size + msd) % GL_P
        for b in range(cb + 1):
            row[4 + 5 * nc + b] = t & 1
            t >>= 1
        row[2] = int(first <= second)
        err = int(bool(first >> (cb * nc)) or bool(second >> (cb * nc)))
        return row, err


# ---------------------------------------------------------------------------------------------
class UXAddManyGate:
    """plonky2_u32 U32AddManyGate (add_many_u32.rs): per op, sum of num_addends limbs + carry-in =
    output_result + 2^bits * output_carry, both range-checked through limb_bits-wide limbs.

    routed wires of op i at (num_addends + 3) * i: addends, carry-in, output_result, output_carry;
    limbs of op i at (num_addends + 3) * num_ops + num_limbs * i.  Constraints per op:
    combined - computed; then, for j = num_limbs-1 .. 0, the range product of limb j; then
    combined_result_limbs - output_result; combined_carry_limbs - output_carry."""

    LOG2_MAX_NUM_ADDENDS = 4

    def __init__(self, bits=29, limb_bits=2, num_addends=3, num_ops=5):
        self.bits, self.limb_bits, self.num_addends, self.num_ops = bits, limb_bits, num_addends, num_ops
        self.num_result_limbs = _ceil_div(bits, limb_bits)
        self.num_carry_limbs = _ceil_div(self.LOG2_MAX_NUM_ADDENDS, limb_bits)
        self.num_limbs = self.num_result_limbs + self.num_carry_limbs

    @staticmethod
    def num_ops_for(num_addends, bits=29, limb_bits=2, num_wires=135, num_routed_wires=80):
        nl = _ceil_div(bits, limb_bits) + _ceil_div(UXAddManyGate.LOG2_MAX_NUM_ADDENDS, limb_bits)
        return min(num_wires // (num_addends + 3 + nl), num_routed_wires // (num_addends + 3))

    @property
    def num_wires(self):
        return (self.num_addends + 3 + self.num_limbs) * self.num_ops

    @property
    def num_constraints(self):
        return (3 + self.num_limbs) * self.num_ops

    @property
    def degree(self):
        return 1 << self.limb_bits

    def wire_ith_op_jth_addend(self, i, j):
        return (self.num_addends + 3) * i + j

    def wire_ith_carry(self, i):
        return (self.num_addends + 3) * i + self.num_addends

    def wire_ith_output_result(self, i):
        return (self.num_addends + 3) * i + self.num_addends + 1

    def wire_ith_output_carry(self, i):
        return (self.num_addends + 3) * i + self.num_addends + 2

    def wire_ith_output_jth_limb(self, i, j):
        return (self.num_addends + 3) * self.num_ops + self.num_limbs * i + j

    def eval_unfiltered(self, w):
        out = []
        lb = 1 << self.limb_bits
        for i in range(self.num_ops):
            computed = (sum(w[self.wire_ith_op_jth_addend(i, j)] for j in range(self.num_addends)) + w[self.wire_ith_carry(i)]) % GL_P
            res, car = w[self.wire_ith_output_result(i)], w[self.wire_ith_output_carry(i)]
            out.append((car * (1 << self.bits) + res - computed) % GL_P)
            comb_res = comb_car = 0
            for j in reversed(range(self.num_limbs)):
                limb = w[self.wire_ith_output_jth_limb(i, j)]
                out.append(_range_product(limb, lb))
                if j < self.num_result_limbs:
                    comb_res = (lb * comb_res + limb) % GL_P
                else:
                    comb_car = (lb * comb_car + limb) % GL_P
            out.append((comb_res - res) % GL_P)
            out.append((comb_car - car) % GL_P)
        return out

    def generate(self, ops):
        """ops: num_ops tuples (addends[num_addends], carry_in).  U32AddManyGenerator::run_once."""
        row = [0] * self.num_wires
        err = 0
        for i, (addends, carry) in enumerate(ops):
            for j, a in enumerate(addends):
                row[self.wire_ith_op_jth_addend(i, j)] = a
            row[self.wire_ith_carry(i)] = carry
            total = (sum(addends) + carry) % GL_P
            res, car = total & ((1 << self.bits) - 1), total >> self.bits
            row[self.wire_ith_output_result(i)] = res
            row[self.wire_ith_output_carry(i)] = car
            limbs = _split(res, self.limb_bits, self.num_result_limbs) + _split(car, self.limb_bits, self.num_carry_limbs)
            if car >> (self.limb_bits * self.num_carry_limbs):
                err = 1
            for j, l in enumerate(limbs):
                row[self.wire_ith_output_jth_limb(i, j)] = l
        return row, err


# ---------------------------------------------------------------------------------------------
class UXSubtractionGate:
    """plonky2_u32 U32SubtractionGate (subtraction_u32.rs): x - y - borrow_in = result - 2^bits * borrow_out.

    routed wires of op i at 5 i: x, y, borrow_in, output_result, output_borrow; limbs of op i at
    5 num_ops + num_limbs i.  Constraints per op: output_result - (x - y - borrow + 2^bits borrow_out);
    range products for j = num_limbs-1 .. 0; combined_limbs - output_result; borrow_out (1 - borrow_out)."""

    def __init__(self, bits=29, limb_bits=2, num_ops=6):
        self.bits, self.limb_bits, self.num_ops = bits, limb_bits, num_ops
        self.num_limbs = _ceil_div(bits, limb_bits)

    @property
    def num_wires(self):
        return (5 + self.num_limbs) * self.num_ops

    @property
    def num_constraints(self):
        return (3 + self.num_limbs) * self.num_ops

    @property
    def degree(self):
        return 1 << self.limb_bits

    def wire_ith_output_jth_limb(self, i, j):
        return 5 * self.num_ops + self.num_limbs * i + j

    def eval_unfiltered(self, w):
        out = []
        lb = 1 << self.limb_bits
        for i in range(self.num_ops):
            x, y, bin_, res, bout = (w[5 * i + k] for k in range(5))
            init = (x - y - bin_) % GL_P
            out.append((res - (init + (1 << self.bits) * bout)) % GL_P)
            comb = 0
            for j in reversed(range(self.num_limbs)):
                limb = w[self.wire_ith_output_jth_limb(i, j)]
                out.append(_range_product(limb, lb))
                comb = (lb * comb + limb) % GL_P
            out.append((comb - res) % GL_P)
            out.append(bout * (1 - bout) % GL_P)
        return out

    def generate(self, ops):
        """ops: num_ops tuples (x, y, borrow_in).  U32SubtractionGenerator::run_once:
        borrow_out = [canonical(x - y - borrow) > 2^bits]."""
        row = [0] * self.num_wires
        err = 0
        for i, (x, y, b) in enumerate(ops):
            row[5 * i], row[5 * i + 1], row[5 * i + 2] = x, y, b
            init = (x - y - b) % GL_P
            bout = int(init > (1 << self.bits))
            res = (init + (bout << self.bits)) % GL_P
            row[5 * i + 3], row[5 * i + 4] = res, bout
            if res >> (self.limb_bits * self.num_limbs):
                err = 1
            for j, l in enumerate(_split(res, self.limb_bits, self.num_limbs)):
                row[self.wire_ith_output_jth_limb(i, j)] = l
        return row, err


# ---------------------------------------------------------------------------------------------
class UXArithmeticGate:
    """plonky2_u32 U32ArithmeticGate (arithmetic_u32.rs): m0 * m1 + addend = low + 2^bits * high.

    routed wires of op i at 6 i: m0, m1, addend, output_low, output_high, inverse; limbs of op i at
    6 num_ops + num_limbs i (num_limbs = 2 ceil(bits / limb_bits): low half then high half).
    Constraints per op: the canonicity guard (inverse * (2^bits - 1 - high) - 1) * low; combined -
    computed; range products for j = num_limbs-1 .. 0; combined_low - low; combined_high - high."""

    def __init__(self, bits=29, limb_bits=2, num_ops=3):
        self.bits, self.limb_bits, self.num_ops = bits, limb_bits, num_ops
        self.half_limbs = _ceil_div(bits, limb_bits)
        self.num_limbs = 2 * self.half_limbs

    @property
    def num_wires(self):
        return (6 + self.num_limbs) * self.num_ops

    @property
    def num_constraints(self):
        return (4 + self.num_limbs) * self.num_ops

    @property
    def degree(self):
        return 1 << self.limb_bits

    def wire_ith_output_jth_limb(self, i, j):
        return 6 * self.num_ops + self.num_limbs * i + j

    def eval_unfiltered(self, w):
        out = []
        lb = 1 << self.limb_bits
        ux_max = (1 << self.bits) - 1
        for i in range(self.num_ops):
            m0, m1, add, lo, hi, inv = (w[6 * i + k] for k in range(6))
            computed = (m0 * m1 + add) % GL_P
            hi_not_max = (inv * ((ux_max - hi) % GL_P) - 1) % GL_P
            out.append(hi_not_max * lo % GL_P)
            out.append((hi * (1 << self.bits) + lo - computed) % GL_P)
            comb_lo = comb_hi = 0
            for j in reversed(range(self.num_limbs)):
                limb = w[self.wire_ith_output_jth_limb(i, j)]
                out.append(_range_product(limb, lb))
                if j < self.half_limbs:
                    comb_lo = (lb * comb_lo + limb) % GL_P
                else:
                    comb_hi = (lb * comb_hi + limb) % GL_P
            out.append((comb_lo - lo) % GL_P)
            out.append((comb_hi - hi) % GL_P)
        return out

    def generate(self, ops):
        """ops: num_ops tuples (m0, m1, addend).  U32ArithmeticGenerator::run_once; the inverse wire is
        (2^bits - 1 - high)^-1, or 0 when high is the maximum (upstream: `try_inverse().unwrap_or(ZERO)`)."""
        row = [0] * self.num_wires
        err = 0
        ux_max = (1 << self.bits) - 1
        for i, (m0, m1, add) in enumerate(ops):
            row[6 * i], row[6 * i + 1], row[6 * i + 2] = m0, m1, add
            out = (m0 * m1 + add) % GL_P
            lo, hi = out & ux_max, out >> self.bits
            d = (ux_max - hi) % GL_P
            row[6 * i + 3], row[6 * i + 4] = lo, hi
            row[6 * i + 5] = pow(d, GL_P - 2, GL_P) if d else 0
            if hi >> (self.limb_bits * self.half_limbs):
                err = 1
            limbs = _split(lo, self.limb_bits, self.half_limbs) + _split(hi, self.limb_bits, self.half_limbs)
            for j, l in enumerate(limbs):
                row[self.wire_ith_output_jth_limb(i, j)] = l
        return row, err
