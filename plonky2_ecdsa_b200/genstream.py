"""Host side of the serialized generator streams (include/p2e.h, p2e_genstream_*): a writer that emits
what `SimpleGenerator::serialize` writes for the reference's in-tree generators, and the handle object of
a loaded stream.  A Rust host would produce the same bytes with plonky2's `Write` trait.

Targets are tuples: ("wire", row, column) or ("virtual", index) -- plonky2's `Target::Wire` /
`Target::VirtualTarget`.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

GEN_MUL_NONNATIVE_BASE, GEN_MUL_NONNATIVE_SCALAR, GEN_CHECK_SUM = 0, 1, 2
GEN_NONNATIVE_ADDITION_BASE, GEN_NONNATIVE_ADDITION_SCALAR = 3, 4
GEN_NONNATIVE_MULTIPLE_ADDS_BASE, GEN_NONNATIVE_MULTIPLE_ADDS_SCALAR = 5, 6
GEN_NONNATIVE_SUBTRACTION_BASE, GEN_NONNATIVE_SUBTRACTION_SCALAR = 7, 8
GEN_NONNATIVE_INVERSE_BASE, GEN_NONNATIVE_INVERSE_SCALAR = 9, 10
GEN_BIGUINT_DIV_REM, GEN_GLV_DECOMPOSITION, GEN_COPY, GEN_CONSTANT, GEN_ARITHMETIC = 11, 12, 13, 14, 15
GEN_BASE_SPLIT, GEN_RANDOM_ACCESS, GEN_EQUALITY, GEN_INPUT = 16, 17, 18, 19


def wire(row: int, column: int):
    return ("wire", row, column)


def virtual(index: int):
    return ("virtual", index)


class GeneratorStreamWriter:
    """Byte-for-byte what the reference's serialize() bodies write, each record prefixed with its u32 tag."""

    def __init__(self):
        self.buf = bytearray()

    # plonky2 util::serialization primitives (little-endian)
    def write_u32(self, v):
        self.buf += struct.pack("<I", v)

    def write_usize(self, v):
        self.buf += struct.pack("<Q", v)

    def write_field(self, v):
        self.buf += struct.pack("<Q", v)

    def write_bool(self, b):
        self.buf += bytes([1 if b else 0])

    def write_target(self, t):
        if t[0] == "wire":
            self.write_bool(True)
            self.write_usize(t[1])
            self.write_usize(t[2])
        else:
            self.write_bool(False)
            self.write_usize(t[1])

    def write_target_bool(self, t):
        """plonky2 Write::write_target_bool: the BoolTarget's target"""
        self.write_target(t)

    def write_target_vec(self, ts):
        self.write_usize(len(ts))
        for t in ts:
            self.write_target(t)

    # records -------------------------------------------------------------------------------------
    def mul_nonnative_generator(self, modulus, row, num_limbs=9):
        """gates/mul_nonnative.rs:326-330"""
        self.write_u32(GEN_MUL_NONNATIVE_SCALAR if modulus else GEN_MUL_NONNATIVE_BASE)
        self.write_usize(row)
        self.write_usize(num_limbs)

    def check_sum_generator(self, row, num_limbs=9):
        """gates/mul_nonnative.rs:533-536"""
        self.write_u32(GEN_CHECK_SUM)
        self.write_usize(row)
        self.write_usize(num_limbs)

    def nonnative_addition_generator(self, modulus, a, b, total, overflow):
        """gadgets/nonnative.rs:647-652"""
        self.write_u32(GEN_NONNATIVE_ADDITION_SCALAR if modulus else GEN_NONNATIVE_ADDITION_BASE)
        self.write_target_vec(a)
        self.write_target_vec(b)
        self.write_target_vec(total)
        self.write_target_bool(overflow)

    def nonnative_multiple_adds_generator(self, modulus, summands, total, overflow):
        """gadgets/nonnative.rs:730-740"""
        self.write_u32(GEN_NONNATIVE_MULTIPLE_ADDS_SCALAR if modulus else GEN_NONNATIVE_MULTIPLE_ADDS_BASE)
        self.write_usize(len(summands))
        for s in summands:
            self.write_target_vec(s)
        self.write_target_vec(total)
        self.write_target(overflow)

    def nonnative_subtraction_generator(self, modulus, a, b, diff, overflow):
        """gadgets/nonnative.rs:812-817"""
        self.write_u32(GEN_NONNATIVE_SUBTRACTION_SCALAR if modulus else GEN_NONNATIVE_SUBTRACTION_BASE)
        self.write_target_vec(a)
        self.write_target_vec(b)
        self.write_target_vec(diff)
        self.write_target_bool(overflow)

    def nonnative_inverse_generator(self, modulus, x, inv, div):
        """gadgets/nonnative.rs:874-878"""
        self.write_u32(GEN_NONNATIVE_INVERSE_SCALAR if modulus else GEN_NONNATIVE_INVERSE_BASE)
        self.write_target_vec(x)
        self.write_target_vec(inv)
        self.write_target_vec(div)

    def biguint_div_rem_generator(self, a, b, div, rem):
        """gadgets/biguint.rs:520-526"""
        self.write_u32(GEN_BIGUINT_DIV_REM)
        for v in (a, b, div, rem):
            self.write_target_vec(v)

    def glv_decomposition_generator(self, k, k1, k2, k1_neg, k2_neg):
        """gadgets/glv.rs:144-150"""
        self.write_u32(GEN_GLV_DECOMPOSITION)
        self.write_target_vec(k)
        self.write_target_vec(k1)
        self.write_target_vec(k2)
        self.write_target_bool(k1_neg)
        self.write_target_bool(k2_neg)

    # the two in-tree GATES' own serialize bodies (gates/mul_nonnative.rs:181-183, 466-468): no tag, one usize
    def mul_nonnative_gate(self, num_limbs=9):
        self.write_usize(num_limbs)

    def check_sum_gate(self, num_limbs=9):
        self.write_usize(num_limbs)

    def copy(self, a, b):
        """CircuitBuilder::connect(a, b): one copy-constraint partition"""
        self.write_u32(GEN_COPY)
        self.write_target(a)
        self.write_target(b)

    def constant_generator(self, row, constant_index, wire_index, constant):
        """plonky2 ConstantGenerator (restated)"""
        self.write_u32(GEN_CONSTANT)
        self.write_usize(row)
        self.write_usize(constant_index)
        self.write_usize(wire_index)
        self.write_field(constant)

    def arithmetic_generator(self, row, const_0, const_1, i):
        """plonky2 ArithmeticBaseGenerator (restated): wire 4i+3 = const_0 * w[4i] * w[4i+1] + const_1 * w[4i+2]"""
        self.write_u32(GEN_ARITHMETIC)
        self.write_usize(row)
        self.write_field(const_0)
        self.write_field(const_1)
        self.write_usize(i)

    def input(self, t):
        """a target of the host's PartialWitness (set before generation)"""
        self.write_u32(GEN_INPUT)
        self.write_target(t)

    def base_split_generator(self, row, num_limbs):
        """plonky2 BaseSplitGenerator<2> (restated): BaseSumGate<2> wire 0 = sum -> wires 1 .. num_limbs = bits"""
        self.write_u32(GEN_BASE_SPLIT)
        self.write_usize(row)
        self.write_usize(num_limbs)

    def random_access_generator(self, row, bits, num_copies, num_extra_constants, copy):
        """plonky2 RandomAccessGenerator of one copy of a RandomAccessGate (restated)"""
        self.write_u32(GEN_RANDOM_ACCESS)
        for v in (row, bits, num_copies, num_extra_constants, copy):
            self.write_usize(v)

    def equality_generator(self, x, y, equal, inv):
        """plonky2 EqualityGenerator (restated)"""
        self.write_u32(GEN_EQUALITY)
        for t in (x, y, equal, inv):
            self.write_target(t)

    def bytes(self) -> bytes:
        return bytes(self.buf)


def deserialize_gate(data: bytes) -> int:
    """Gate::deserialize of MulNonnativeGate / CheckSumGate (gates/mul_nonnative.rs:185-195, 470-477): reads num_limbs.
    The kernels implement num_limbs() == 9 (gates/mul_nonnative.rs:37-39, 353-355); anything else is refused."""
    if len(data) != 8:
        raise ValueError("gate record is one usize")
    (nl,) = struct.unpack("<Q", data)
    if nl != 9:
        raise ValueError("num_limbs must be 9")
    return nl


class GeneratorStream:
    """A loaded stream (p2e_genstream_load).  Loading, slot / level queries need no device; run() does."""

    def __init__(self, lib, data: bytes, error_cls=RuntimeError):
        self.lib = lib
        self._err = error_cls
        buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
        self.handle = lib.p2e_genstream_load(ctypes.addressof(buf), len(data))
        if self.handle < 0:
            raise error_cls(lib.p2e_last_error().decode())

    def close(self):
        if self.handle >= 0:
            self.lib.p2e_genstream_free(self.handle)
            self.handle = -1

    def _info(self, what):
        v = self.lib.p2e_genstream_info(self.handle, what)
        if v < 0:
            raise self._err(self.lib.p2e_last_error().decode())
        return int(v)

    num_slots = property(lambda self: self._info(0))
    num_generators = property(lambda self: self._info(1))
    num_levels = property(lambda self: self._info(2))
    num_groups = property(lambda self: self._info(3))
    num_targets = property(lambda self: self._info(5))

    def slot(self, target) -> int:
        if target[0] == "wire":
            v = self.lib.p2e_genstream_slot(self.handle, 1, target[1], target[2])
        else:
            v = self.lib.p2e_genstream_slot(self.handle, 0, target[1], 0)
        if v < 0:
            raise self._err(self.lib.p2e_last_error().decode())
        return int(v)

    @property
    def input_slots(self):
        n = self._info(4)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.p2e_genstream_input_slots(self.handle, out.ctypes.data, n)
        return [int(v) for v in out[:n]]

    @property
    def generator_levels(self):
        n = self._info(1)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.p2e_genstream_generator_levels(self.handle, out.ctypes.data, n)
        return [int(v) for v in out[:n]]

    def run(self, inputs: dict, outputs: list):
        """inputs: {target: values [n]} for every input partition; outputs: targets to read back.
        Returns (values [len(outputs)][n], err [n])."""
        items = list(inputs.items())
        n = len(np.asarray(items[0][1])) if items else 0
        in_slots = np.array([self.slot(t) for t, _ in items], dtype=np.uint32)
        in_vals = np.ascontiguousarray(np.array([np.asarray(v, dtype=np.uint64) for _, v in items], dtype=np.uint64).reshape(len(items), n))
        out_slots = np.array([self.slot(t) for t in outputs], dtype=np.uint32)
        out_vals = np.zeros((len(outputs), n), dtype=np.uint64)
        err = np.zeros(n, dtype=np.uint32)
        rc = self.lib.p2e_genstream_run(self.handle, n, in_slots.ctypes.data, len(items), in_vals.ctypes.data,
                                        out_slots.ctypes.data, len(outputs), out_vals.ctypes.data, err.ctypes.data)
        if rc != 0:
            raise self._err(self.lib.p2e_last_error().decode())
        return out_vals, err


class TargetCircuitBuilder:
    """The reference's gadgets at the TARGET level (what `CircuitBuilder` + the extension traits of
    gadgets/nonnative.rs / gadgets/curve.rs do when a circuit is built): allocates virtual targets and gate
    rows, records copy constraints, and emits every generator through a GeneratorStreamWriter -- i.e. the
    byte stream a Rust host would hand to p2e_genstream_load for the same circuit.  A nonnative value is a
    list of 9 limb targets."""

    def __init__(self):
        self.w = GeneratorStreamWriter()
        self.n_virtual = 0
        self.n_rows = 0

    def add_virtual_targets(self, k):
        v = [virtual(self.n_virtual + i) for i in range(k)]
        self.n_virtual += k
        return v

    def add_virtual_nonnative_target(self):
        return self.add_virtual_targets(9)

    def add_gate(self):
        self.n_rows += 1
        return self.n_rows - 1

    def connect(self, a, b):
        self.w.copy(a, b)

    def add_nonnative(self, modulus, a, b):
        """gadgets/nonnative.rs:245-276"""
        s, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_addition_generator(modulus, a, b, s, ov)
        return s

    def sub_nonnative(self, modulus, a, b):
        """gadgets/nonnative.rs:356-388"""
        d, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_subtraction_generator(modulus, a, b, d, ov)
        return d

    def add_many_nonnative(self, modulus, xs):
        """gadgets/nonnative.rs:310-353"""
        s, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_multiple_adds_generator(modulus, xs, s, ov)
        return s

    def mul_nonnative(self, modulus, x, y):
        """gadgets/nonnative.rs:390-464: a MulNonnativeGate row and a CheckSumGate row, both fed by copies"""
        row = self.add_gate()
        for i in range(9):
            self.connect(x[i], wire(row, i))
            self.connect(y[i], wire(row, 9 + i))
        self.w.mul_nonnative_generator(modulus, row)
        row2 = self.add_gate()
        for i in range(17):
            self.connect(wire(row, 36 + i), wire(row2, i))
        self.w.check_sum_generator(row2)
        return [wire(row, 18 + i) for i in range(9)]

    def inv_nonnative(self, modulus, x):
        """gadgets/nonnative.rs:502-536"""
        inv, div = self.add_virtual_targets(9), self.add_virtual_targets(9)
        self.w.nonnative_inverse_generator(modulus, x, inv, div)
        return inv

    def curve_add(self, p1, p2):
        """gadgets/curve.rs:202-223"""
        (x1, y1), (x2, y2) = p1, p2
        u = self.sub_nonnative(0, y2, y1)
        v = self.sub_nonnative(0, x2, x1)
        v_inv = self.inv_nonnative(0, v)
        s = self.mul_nonnative(0, u, v_inv)
        s2 = self.mul_nonnative(0, s, s)
        x_sum = self.add_nonnative(0, x2, x1)
        x3 = self.sub_nonnative(0, s2, x_sum)
        x_diff = self.sub_nonnative(0, x1, x3)
        prod = self.mul_nonnative(0, s, x_diff)
        y3 = self.sub_nonnative(0, prod, y1)
        return x3, y3

    def curve_double(self, p, a_const):
        """gadgets/curve.rs:160-185; a_const: the 9 limb targets of the curve constant A"""
        x, y = p
        double_y = self.add_nonnative(0, y, y)
        inv_double_y = self.inv_nonnative(0, double_y)
        x_squared = self.mul_nonnative(0, x, x)
        triple = self.add_many_nonnative(0, [x_squared, x_squared, x_squared, a_const])
        lam = self.mul_nonnative(0, triple, inv_double_y)
        lam2 = self.mul_nonnative(0, lam, lam)
        x_double = self.add_nonnative(0, x, x)
        x3 = self.sub_nonnative(0, lam2, x_double)
        x_diff = self.sub_nonnative(0, x, x3)
        lx = self.mul_nonnative(0, lam, x_diff)
        y3 = self.sub_nonnative(0, lx, y)
        return x3, y3

    def bytes(self) -> bytes:
        return self.w.bytes()


# --------------------------------------------------------------------------------------------------------------
# The complete verify_secp256k1_message_circuit at the target level: what a Rust host serialises for the reference's
# own circuit (gadgets/ecdsa.rs:30-53), including the plonky2 glue between the in-tree generators (constants,
# arithmetic gates, base-2 splits, random accesses, equality).  Gate-row packing follows plonky2's habits loosely
# (one ConstantGate output per row, 20 arithmetic operations per row, 4 random accesses per row): the physical
# placement is the host framework's, only the generator graph matters to the executor.
SECP_P = (1 << 256) - (1 << 32) - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GL_P = (1 << 64) - (1 << 32) + 1
_GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
_GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
_GLV_BETA = 0x7AE96A2B657C07106E64479EAC3434E99CF0497512F58995C1396C28719501EE
_GLV_S = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def _pt_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % SECP_P == 0:
            return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, SECP_P) % SECP_P
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, SECP_P) % SECP_P
    x = (lam * lam - P[0] - Q[0]) % SECP_P
    return (x, (lam * (P[0] - x) - P[1]) % SECP_P)


def _pt_mul(k, P):
    R = None
    while k:
        if k & 1:
            R = _pt_add(R, P)
        P = _pt_add(P, P)
        k >>= 1
    return R


def _pt_neg(P):
    return (P[0], (-P[1]) % SECP_P)


def constant_limbs(v):
    """constant_biguint: to_u32_digits then convert_base 32 -> 29 (gadgets/biguint.rs:27-51, 165-175): zero has no limbs,
    a trailing zero limb is kept when the 32-bit digits leave whole 29-bit groups behind."""
    digits = []
    while v:
        digits.append(v & 0xFFFFFFFF)
        v >>= 32
    res, rem, offset = [], 0, 0
    for d in digits:
        rem += d << offset
        offset += 32
        while offset >= 29:
            res.append(rem & 0x1FFFFFFF)
            rem >>= 29
            offset -= 29
    if rem:
        res.append(rem)
    return res


class EcdsaCircuitBuilder(TargetCircuitBuilder):
    """verify_secp256k1_message_circuit (gadgets/ecdsa.rs:30-53) and every gadget under it, at the target level.
    `tape_targets` lists the output targets of the in-tree generators in gadget call order -- the tape layout of
    p2e_ecdsa_tape_layout -- so a stream run can be compared slot for slot with the chain kernels' witness."""

    ARITH_OPS_PER_ROW = 20
    RA_BITS, RA_COPIES, RA_EXTRA = 4, 4, 2          # RandomAccessGate::new_from_config(standard_ecc_config, 4): 80 routed wires

    def __init__(self, blinding_scalar):
        super().__init__()
        self.blinding_scalar = blinding_scalar        # keccak(0u64) as a scalar (curve_msm.rs:33-39): a circuit constant
        self.tape_targets = []
        self._consts = {}
        self._arith = {}
        self._ra = None

    # ---- plonky2 glue -------------------------------------------------------------------------------------
    def constant(self, c):
        c %= GL_P
        if c not in self._consts:
            row = self.add_gate()
            self.w.constant_generator(row, 0, 0, c)
            self._consts[c] = wire(row, 0)
        return self._consts[c]

    def zero(self):
        return self.constant(0)

    def one(self):
        return self.constant(1)

    def arithmetic(self, c0, c1, m0, m1, addend):
        key = (c0 % GL_P, c1 % GL_P)
        row, i = self._arith.get(key, (None, self.ARITH_OPS_PER_ROW))
        if i == self.ARITH_OPS_PER_ROW:
            row, i = self.add_gate(), 0
        self._arith[key] = (row, i + 1)
        self.w.arithmetic_generator(row, key[0], key[1], i)
        for k, t in enumerate((m0, m1, addend)):
            self.connect(t, wire(row, 4 * i + k))
        return wire(row, 4 * i + 3)

    def mul_add(self, a, b, c):
        return self.arithmetic(1, 1, a, b, c)

    def mul(self, a, b):
        return self.arithmetic(1, 0, a, b, self.zero())

    def sub(self, a, b):
        return self.arithmetic(1, GL_P - 1, a, self.one(), b)

    def not_(self, b):
        return self.sub(self.one(), b)

    def is_equal(self, x, y):
        equal, inv = self.add_virtual_targets(2)
        self.w.equality_generator(x, y, equal, inv)
        return equal

    def split_le_base2(self, x, num_limbs):
        row = self.add_gate()
        self.connect(x, wire(row, 0))
        self.w.base_split_generator(row, num_limbs)
        return [wire(row, 1 + i) for i in range(num_limbs)]

    def random_access(self, index, items):
        assert len(items) == 1 << self.RA_BITS
        if self._ra is None or self._ra[1] == self.RA_COPIES:
            self._ra = [self.add_gate(), 0]
        row, copy = self._ra
        self._ra[1] += 1
        vs = 1 << self.RA_BITS
        base = (2 + vs) * copy
        self.connect(index, wire(row, base))
        for k, t in enumerate(items):
            self.connect(t, wire(row, base + 2 + k))
        self.w.random_access_generator(row, self.RA_BITS, self.RA_COPIES, self.RA_EXTRA, copy)
        return wire(row, base + 1)

    # ---- gadgets/nonnative.rs (values are lists of limb targets; constants may have fewer than 9 limbs) ------
    def constant_nonnative(self, v, modulus):
        return [self.constant(l) for l in constant_limbs(v % (SECP_N if modulus else SECP_P))]

    def add_nonnative(self, modulus, a, b):
        s = super().add_nonnative(modulus, a, b)
        self.tape_targets += s + [virtual(self.n_virtual - 1)]
        return s

    def sub_nonnative(self, modulus, a, b):
        d = super().sub_nonnative(modulus, a, b)
        self.tape_targets += d + [virtual(self.n_virtual - 1)]
        return d

    def add_many_nonnative(self, modulus, xs):
        if len(xs) == 1:
            return xs[0]
        s = super().add_many_nonnative(modulus, xs)
        self.tape_targets += s + [virtual(self.n_virtual - 1)]
        return s

    def mul_nonnative(self, modulus, x, y):
        """nonnative.rs:390-464; absent limbs are connected to zero (:406-424)"""
        row = self.add_gate()
        for i in range(9):
            self.connect(x[i] if i < len(x) else self.zero(), wire(row, i))
            self.connect(y[i] if i < len(y) else self.zero(), wire(row, 9 + i))
        self.w.mul_nonnative_generator(modulus, row)
        row2 = self.add_gate()
        for i in range(17):
            self.connect(wire(row, 36 + i), wire(row2, i))
        self.w.check_sum_generator(row2)
        self.tape_targets += [wire(row, 18 + i) for i in range(35)] + [wire(row2, 17 + i) for i in range(16)]
        return [wire(row, 18 + i) for i in range(9)]

    def inv_nonnative(self, modulus, x):
        inv, div = self.add_virtual_targets(9), self.add_virtual_targets(9)
        self.w.nonnative_inverse_generator(modulus, x, inv, div)
        self.tape_targets += div + inv
        return inv

    def neg_nonnative(self, modulus, x):
        return self.sub_nonnative(modulus, [], x)             # nonnative.rs:491-500 (zero has no limbs)

    def mul_nonnative_by_bool(self, a, b):
        return [self.mul(l, b) for l in a]                      # biguint.rs:360-374

    def nonnative_conditional_neg(self, modulus, x, b):
        not_b = self.not_(b)                                    # nonnative.rs:584-596
        neg = self.neg_nonnative(modulus, x)
        t = self.mul_nonnative_by_bool(neg, b)
        f = self.mul_nonnative_by_bool(x, not_b)
        return self.add_nonnative(modulus, t, f)

    def connect_nonnative(self, a, b):
        for i in range(max(len(a), len(b))):
            self.connect(a[i] if i < len(a) else self.zero(), b[i] if i < len(b) else self.zero())

    # ---- gadgets/curve.rs ---------------------------------------------------------------------------------
    def constant_affine_point(self, P):
        return (self.constant_nonnative(P[0], 0), self.constant_nonnative(P[1], 0))

    def curve_assert_valid(self, p):
        a, b = self.constant_nonnative(0, 0), self.constant_nonnative(7, 0)     # curve.rs:123-135
        y2 = self.mul_nonnative(0, p[1], p[1])
        x2 = self.mul_nonnative(0, p[0], p[0])
        x3 = self.mul_nonnative(0, x2, p[0])
        ax = self.mul_nonnative(0, a, p[0])
        axb = self.add_nonnative(0, ax, b)
        rhs = self.add_nonnative(0, x3, axb)
        self.connect_nonnative(y2, rhs)

    def curve_conditional_neg(self, p, b):
        return (p[0], self.nonnative_conditional_neg(0, p[1], b))

    def curve_double(self, p, a_const=None):
        return super().curve_double(p, self.constant_nonnative(0, 0))

    def curve_conditional_add(self, p1, p2, b):
        not_b = self.not_(b)                                    # curve.rs:225-243
        s = self.curve_add(p1, p2)
        xt = self.mul_nonnative_by_bool(s[0], b)
        yt = self.mul_nonnative_by_bool(s[1], b)
        xf = self.mul_nonnative_by_bool(p1[0], not_b)
        yf = self.mul_nonnative_by_bool(p1[1], not_b)
        return (self.add_nonnative(0, xt, xf), self.add_nonnative(0, yt, yf))

    def random_access_curve_points(self, index, pts):
        """curve_windowed_mul.rs:74-118: 9 + 9 random accesses over per-limb columns (absent limbs read zero)"""
        def col(c, i):
            return [p[c][i] if i < len(p[c]) else self.zero() for p in pts]
        return ([self.random_access(index, col(0, i)) for i in range(9)], [self.random_access(index, col(1, i)) for i in range(9)])

    # ---- gadgets/split_nonnative.rs -------------------------------------------------------------------------
    def _bits(self, x):
        bits = []
        for l in x:
            bits += self.split_le_base2(l, 29)
        return bits

    def split_nonnative_to_4_bit_limbs(self, x):
        bits = self._bits(x)
        while len(bits) % 4:
            bits.append(self.zero())
        two, four = self.constant(2), self.constant(4)
        out = []
        for k in range(0, len(bits), 4):
            a, b, c, d = bits[k:k + 4]
            lower = self.mul_add(b, two, a)
            upper = self.mul_add(d, two, c)
            out.append(self.mul_add(upper, four, lower))
        return out

    def split_nonnative_to_2_bit_limbs(self, x):
        bits = self._bits(x)
        while len(bits) % 2:
            bits.append(self.zero())
        two = self.constant(2)
        return [self.mul_add(bits[k + 1], two, bits[k]) for k in range(0, len(bits), 2)]

    # ---- gadgets/curve_fixed_base.rs:18-66 ------------------------------------------------------------------
    def fixed_base_curve_mul(self, base, scalar):
        limbs = self.split_nonnative_to_4_bit_limbs(scalar)
        rando = _pt_mul(self.blinding_scalar, (_GX, _GY))
        zero = self.zero()
        result = self.constant_affine_point(rando)
        point = base
        for limb in limbs[:len(scalar) * 8]:
            muls, acc = [], None
            for _ in range(16):
                muls.append(acc)
                acc = _pt_add(point, acc)
            muls = muls[1:]
            muls.insert(0, muls[0])
            pts = [self.constant_affine_point(m) for m in muls]
            is_zero = self.is_equal(limb, zero)
            should_add = self.not_(is_zero)
            r = self.random_access_curve_points(limb, pts)
            result = self.curve_conditional_add(result, r, should_add)
            for _ in range(4):
                point = _pt_add(point, point)
        return self.curve_add(result, self.constant_affine_point(_pt_neg(rando)))

    # ---- gadgets/curve_msm.rs:21-79 -------------------------------------------------------------------------
    def curve_msm(self, p, q, n, m):
        limbs_n = self.split_nonnative_to_2_bit_limbs(n)
        limbs_m = self.split_nonnative_to_2_bit_limbs(m)
        assert len(limbs_n) == len(limbs_m)
        rando = _pt_mul(self.blinding_scalar, (_GX, _GY))
        rando_t = self.constant_affine_point(rando)
        neg_rando = self.constant_affine_point(_pt_neg(rando))
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
        four, zero = self.constant(4), self.zero()
        result = rando_t
        for ln, lm in reversed(list(zip(limbs_n, limbs_m))):
            result = self.curve_double(self.curve_double(result))
            index = self.mul_add(four, lm, ln)
            r = self.random_access_curve_points(index, pre)
            is_zero = self.is_equal(index, zero)
            should_add = self.not_(is_zero)
            result = self.curve_conditional_add(result, r, should_add)
        spm = rando
        for _ in range(2 * len(limbs_n)):
            spm = _pt_add(spm, spm)
        return self.curve_add(result, self.constant_affine_point(_pt_neg(spm)))

    # ---- gadgets/glv.rs -----------------------------------------------------------------------------------------
    def decompose_secp256k1_scalar(self, k):
        k1, k2 = self.add_virtual_targets(5), self.add_virtual_targets(5)          # glv.rs:53-85
        n1, n2 = self.add_virtual_targets(2)
        self.w.glv_decomposition_generator(k, k1, k2, n1, n2)
        self.tape_targets += k1 + k2 + [n1, n2]
        k1_raw = self.nonnative_conditional_neg(1, k1, n1)
        k2_raw = self.nonnative_conditional_neg(1, k2, n2)
        s = self.constant_nonnative(_GLV_S, 1)
        sbk = self.mul_nonnative(1, s, k2_raw)
        sbk = self.add_nonnative(1, sbk, k1_raw)
        self.connect_nonnative(sbk, k)
        return k1, k2, n1, n2

    def glv_mul(self, p, k):
        k1, k2, n1, n2 = self.decompose_secp256k1_scalar(k)                          # glv.rs:87-104
        beta = self.constant_nonnative(_GLV_BETA, 0)
        bpx = self.mul_nonnative(0, beta, p[0])
        sp = (bpx, p[1])
        p_neg = self.curve_conditional_neg(p, n1)
        sp_neg = self.curve_conditional_neg(sp, n2)
        return self.curve_msm(p_neg, sp_neg, k1, k2)

    # ---- gadgets/ecdsa.rs:30-53 ---------------------------------------------------------------------------------
    def verify_secp256k1_message(self):
        """Allocates msg, r, s, pk.x, pk.y as virtual nonnative targets (the stream's inputs) and returns them."""
        msg, r, s = (self.add_virtual_nonnative_target() for _ in range(3))
        pk = (self.add_virtual_nonnative_target(), self.add_virtual_nonnative_target())
        for v in (msg, r, s, pk[0], pk[1]):
            for t in v:
                self.w.input(t)
        self.curve_assert_valid(pk)
        c = self.inv_nonnative(1, s)
        u1 = self.mul_nonnative(1, msg, c)
        u2 = self.mul_nonnative(1, r, c)
        p1 = self.fixed_base_curve_mul((_GX, _GY), u1)
        p2 = self.glv_mul(pk, u2)
        pt = self.curve_add(p1, p2)
        self.connect_nonnative(r, pt[0])
        return msg, r, s, pk
