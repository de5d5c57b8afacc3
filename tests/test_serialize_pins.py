"""What CAN be pinned of the stream format without the out-of-tree sources: the field order and the `Write` primitive used
for each field of every in-tree `serialize` body.  tests/golden/serialize_bodies.json is extracted from the reference
checkout by tests/golden/make_serialize_golden.py; the writer of plonky2_ecdsa_b200.genstream must issue exactly those
primitive calls in that order.  (The byte width of `write_usize` / the encoding of a `Target` are plonky2's -- out of tree,
restated, unpinned: they are isolated in the four write_* primitives of GeneratorStreamWriter / the Reader of
csrc/p2e_genstream.cu, and in-tree usage fixes nothing about them beyond "usize is the only integer primitive used".)"""
import importlib.util
import json
import os

import pytest

from plonky2_ecdsa_b200 import genstream as G

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "serialize_bodies.json")))


class Recorder(G.GeneratorStreamWriter):
    def __init__(self):
        super().__init__()
        self.calls = []
        self.depth = 0

    def _rec(self, name, fn, *a):
        if self.depth == 0:
            self.calls.append(name)
        self.depth += 1
        try:
            return fn(*a)
        finally:
            self.depth -= 1

    def write_usize(self, v):
        return self._rec("write_usize", super().write_usize, v)

    def write_target(self, t):
        return self._rec("write_target", super().write_target, t)

    def write_target_bool(self, t):
        return self._rec("write_target_bool", super().write_target_bool, t)

    def write_target_vec(self, ts):
        return self._rec("write_target_vec", super().write_target_vec, ts)

    def write_u32(self, v):            # the record tag (the host's generator registry index): not part of serialize()
        return super().write_u32(v)


def _v(k, n=9):
    return [G.virtual(100 * k + i) for i in range(n)]


CASES = {
    "MulNonnativeGenerator": lambda w: w.mul_nonnative_generator(0, 3),
    "CheckSumGenerator": lambda w: w.check_sum_generator(4),
    "NonNativeAdditionGenerator": lambda w: w.nonnative_addition_generator(0, _v(1), _v(2), _v(3), G.virtual(7)),
    "NonNativeSubtractionGenerator": lambda w: w.nonnative_subtraction_generator(1, _v(1), _v(2), _v(3), G.virtual(7)),
    "NonNativeInverseGenerator": lambda w: w.nonnative_inverse_generator(0, _v(1), _v(2), _v(3)),
    "BigUintDivRemGenerator": lambda w: w.biguint_div_rem_generator(_v(1), _v(2, 3), _v(3, 7), _v(4, 3)),
    "GLVDecompositionGenerator": lambda w: w.glv_decomposition_generator(_v(1), _v(2, 5), _v(3, 5), G.virtual(8), G.virtual(9)),
    "MulNonnativeGate": lambda w: w.mul_nonnative_gate(),
    "CheckSumGate": lambda w: w.check_sum_gate(),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_writer_issues_the_reference_bodies_primitive_calls(name):
    w = Recorder()
    CASES[name](w)
    assert w.calls == [prim for prim, _ in GOLDEN[name]], (name, w.calls, GOLDEN[name])


def test_multiple_adds_body_repeats_the_summand_vector():
    """gadgets/nonnative.rs:730-740: usize summands_len, then one target_vec PER summand, the sum, the overflow target."""
    w = Recorder()
    w.nonnative_multiple_adds_generator(0, [_v(1), _v(2), _v(3), _v(4, 0)], _v(5), G.virtual(9))
    want = [p for p, _ in GOLDEN["NonNativeMultipleAddsGenerator"]]
    assert want == ["write_usize", "write_target_vec", "write_target_vec", "write_target"]
    assert w.calls == ["write_usize"] + ["write_target_vec"] * 4 + ["write_target_vec", "write_target"]


def test_golden_is_what_the_reference_checkout_says():
    """Re-extracts from /root/reference when it is present (the build container); skipped on the GPU box."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference checkout not present")
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_serialize_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert mk.extract("/root/reference") == GOLDEN


def test_only_usize_integers_appear_in_tree():
    prims = {p for body in GOLDEN.values() for p, _ in body}
    assert prims == {"write_usize", "write_target", "write_target_bool", "write_target_vec"}


def test_gate_records_round_trip():
    w = G.GeneratorStreamWriter()
    w.mul_nonnative_gate()
    assert G.deserialize_gate(w.bytes()) == 9
    with pytest.raises(ValueError):
        G.deserialize_gate(b"\x08" + b"\0" * 7)
