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
        self.write_target(overflow)

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
        self.write_target(overflow)

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
        self.write_target(k1_neg)
        self.write_target(k2_neg)

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

    def bytes(self) -> bytes:
        return bytes(self.buf)


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
