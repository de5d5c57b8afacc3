"""Serialized generator streams + the level-scheduled executor (include/p2e.h
p2e_genstream_*).  The test builds circuits at the TARGET level the way the reference's gadgets do
(gadgets/nonnative.rs:245-276, 356-388, 390-464, 502-536; gadgets/curve.rs:160-223), serializes the
generators with the reference's own serialize() bodies, and compares every target the engine produces with
a worklist evaluation of the same generators by the oracle (plonky2's generate_partial_witness, restated:
run a generator once all of its dependencies are set; copies share one value)."""
import random
import struct

import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import GeneratorStreamWriter, virtual, wire
from oracle import spec as S

P, N = S.SECP_P, S.SECP_N
GL_P = S.GL_P


class TargetCircuit:
    """CircuitBuilder stand-in: allocates targets / gate rows, records generators + copy constraints, and
    can evaluate itself with the oracle's generator functions."""

    def __init__(self):
        self.w = GeneratorStreamWriter()
        self.n_virtual = 0
        self.n_rows = 0
        self.gens = []          # (fn, ins, outs): fn(list of input values) -> list of output values
        self.copies = []

    def add_virtual_targets(self, k):
        v = [virtual(self.n_virtual + i) for i in range(k)]
        self.n_virtual += k
        return v

    def add_gate(self):
        self.n_rows += 1
        return self.n_rows - 1

    def connect(self, a, b):
        self.copies.append((a, b))
        self.w.copy(a, b)

    # gadgets/nonnative.rs:245-276
    def add_nonnative(self, m, a, b):
        s, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_addition_generator(m, a, b, s, ov)
        mod = N if m else P
        self.gens.append((lambda v, la=len(a): (lambda r: r[0] + [r[1]])(S.nonnative_add_gen(v[:la], v[la:], mod)), a + b, s + [ov]))
        return s

    # gadgets/nonnative.rs:356-388
    def sub_nonnative(self, m, a, b):
        d, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_subtraction_generator(m, a, b, d, ov)
        mod = N if m else P
        self.gens.append((lambda v, la=len(a): (lambda r: r[0] + [r[1]])(S.nonnative_sub_gen(v[:la], v[la:], mod)), a + b, d + [ov]))
        return d

    # gadgets/nonnative.rs:310-353
    def add_many_nonnative(self, m, xs):
        s, ov = self.add_virtual_targets(9), self.add_virtual_targets(1)[0]
        self.w.nonnative_multiple_adds_generator(m, xs, s, ov)
        mod = N if m else P
        k = len(xs)
        self.gens.append((lambda v: (lambda r: r[0] + [r[1]])(S.nonnative_add_many_gen([v[9 * t:9 * t + 9] for t in range(k)], mod)),
                          [t for x in xs for t in x], s + [ov]))
        return s

    # gadgets/nonnative.rs:390-464: a MulNonnativeGate row fed by copies, a CheckSumGate row fed by copies
    def mul_nonnative(self, m, x, y):
        row = self.add_gate()
        for i in range(9):
            self.connect(x[i], wire(row, i))
            self.connect(y[i], wire(row, 9 + i))
        self.w.mul_nonnative_generator(m, row)
        mod = N if m else P

        def mul(v):
            q, r, cs = S.mul_nonnative_gen(v[:9], v[9:], mod)
            return r + q + cs
        self.gens.append((mul, [wire(row, c) for c in range(18)], [wire(row, c) for c in range(18, 53)]))
        row2 = self.add_gate()
        for i in range(17):
            self.connect(wire(row, 36 + i), wire(row2, i))
        self.w.check_sum_generator(row2)
        self.gens.append((S.check_sum_gen, [wire(row2, c) for c in range(17)], [wire(row2, c) for c in range(17, 33)]))
        return [wire(row, 18 + i) for i in range(9)]

    # gadgets/nonnative.rs:502-536
    def inv_nonnative(self, m, x):
        inv, div = self.add_virtual_targets(9), self.add_virtual_targets(9)
        self.w.nonnative_inverse_generator(m, x, inv, div)
        mod = N if m else P
        self.gens.append((lambda v: (lambda r: r[0] + r[1])(S.nonnative_inv_gen(v, mod)), list(x), div + inv))
        return inv

    # gadgets/curve.rs:202-223
    def curve_add(self, p1, p2):
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

    # gadgets/curve.rs:160-185 (A = 0 enters as a constant target: an input here)
    def curve_double(self, p, a_const):
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

    # ---- oracle evaluation: plonky2's generate_partial_witness restated (worklist over partitions)
    def evaluate(self, inputs):
        parent = {}

        def find(t):
            parent.setdefault(t, t)
            while parent[t] != t:
                parent[t] = parent[parent[t]]
                t = parent[t]
            return t
        for a, b in self.copies:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[ra] = rb
        val = {find(t): v for t, v in inputs.items()}
        pending = list(self.gens)
        while pending:
            rest = []
            for fn, ins, outs in pending:
                if all(find(t) in val for t in ins):
                    res = fn([val[find(t)] for t in ins])
                    for t, v in zip(outs, res):
                        r = find(t)
                        assert val.get(r, v) == v, "partition set twice with different values"
                        val[r] = v
                else:
                    rest.append((fn, ins, outs))
            assert len(rest) < len(pending), "generators left unfinished"
            pending = rest
        return lambda t: val[find(t)]

    def all_targets(self):
        ts = set()
        for _, ins, outs in self.gens:
            ts.update(ins)
            ts.update(outs)
        return sorted(ts)


def _curve_circuit():
    c = TargetCircuit()
    pts = [(c.add_virtual_targets(9), c.add_virtual_targets(9)) for _ in range(2)]
    a_const = c.add_virtual_targets(9)
    s = c.curve_add(pts[0], pts[1])
    d = c.curve_double(s, a_const)
    t = c.curve_add(d, pts[0])
    return c, pts, a_const, t


# ------------------------------------------------------------------------------------------ CPU tier
def test_stream_bytes_follow_the_reference_serialize_bodies():
    w = GeneratorStreamWriter()
    w.mul_nonnative_generator(0, 7)
    # tag, then write_usize(row), write_usize(num_limbs)  (gates/mul_nonnative.rs:326-330)
    assert w.bytes() == struct.pack("<IQQ", 0, 7, 9)
    w = GeneratorStreamWriter()
    w.nonnative_subtraction_generator(1, [virtual(3)], [wire(2, 5)], [virtual(i) for i in range(9)], virtual(11))
    b = w.bytes()
    assert b[:4] == struct.pack("<I", 8)
    assert b[4:12] == struct.pack("<Q", 1) and b[12] == 0 and b[13:21] == struct.pack("<Q", 3)        # a: one virtual target
    assert b[21:29] == struct.pack("<Q", 1) and b[29] == 1 and b[30:46] == struct.pack("<QQ", 2, 5)   # b: one wire
    assert len(b) == 4 + (8 + 9) + (8 + 17) + (8 + 9 * 9) + 9


def test_product_side_builder_emits_the_same_stream():
    """plonky2_ecdsa_b200.TargetCircuitBuilder (the host mirror bench.py uses) and this file's oracle-carrying
    TargetCircuit wire the gadgets identically."""
    c, _, _, _ = _curve_circuit()
    b = pkg.TargetCircuitBuilder()
    pts = [(b.add_virtual_nonnative_target(), b.add_virtual_nonnative_target()) for _ in range(2)]
    a_const = b.add_virtual_nonnative_target()
    b.curve_add(b.curve_double(b.curve_add(pts[0], pts[1]), a_const), pts[0])
    assert b.bytes() == c.w.bytes()


def test_load_partitions_levels_and_inputs():
    c, pts, a_const, t = _curve_circuit()
    lib = pkg.load_library()
    g = pkg.GeneratorStream(lib, c.w.bytes())
    assert g.num_generators == len(c.gens) == 41
    # every connect() merges two targets into one slot
    assert g.num_slots == g.num_targets - len(c.copies)
    assert g.slot(wire(0, 3)) == g.slot(c.gens[0][2][3])             # x limb 3 of the first product = u limb 3
    # inputs: 2 points + the constant A, 9 limbs each
    want = sorted(g.slot(tt) for p in pts for co in p for tt in co) + sorted(g.slot(tt) for tt in a_const)
    assert sorted(g.input_slots) == sorted(want)
    # levels follow the dataflow: u, v, x_sum at level 0; v_inv after v; s after v_inv; the CheckSum row after its product
    lv = g.generator_levels
    assert lv[0] == lv[1] == 0 and lv[2] == 1 and lv[3] == 2 and lv[4] == 3
    assert g.num_levels == max(lv) + 1 and g.num_levels >= 20
    # same level + same kind + same shape share a launch group
    assert g.num_groups < g.num_generators
    g.close()


def test_load_rejects_malformed_streams():
    lib = pkg.load_library()
    good = GeneratorStreamWriter()
    good.check_sum_generator(1)
    with pytest.raises(RuntimeError, match="truncated"):
        pkg.GeneratorStream(lib, good.bytes()[:-3])
    with pytest.raises(RuntimeError, match="unknown generator tag"):
        pkg.GeneratorStream(lib, struct.pack("<I", 99))
    bad = GeneratorStreamWriter()
    bad.mul_nonnative_generator(0, 1, num_limbs=8)
    with pytest.raises(RuntimeError, match="num_limbs"):
        pkg.GeneratorStream(lib, bad.bytes())
    # a dependency cycle can never be scheduled (plonky2 would leave both generators unfinished)
    cyc = GeneratorStreamWriter()
    a, b = [virtual(i) for i in range(9)], [virtual(9 + i) for i in range(9)]
    cyc.nonnative_inverse_generator(0, a, b, [virtual(18 + i) for i in range(9)])
    cyc.nonnative_inverse_generator(0, b, a, [virtual(27 + i) for i in range(9)])
    with pytest.raises(RuntimeError, match="cycle"):
        pkg.GeneratorStream(lib, cyc.bytes())
    # div / rem sizes must follow biguint.rs:391-399
    dr = GeneratorStreamWriter()
    dr.biguint_div_rem_generator([virtual(i) for i in range(9)], [virtual(9 + i) for i in range(5)],
                                 [virtual(20 + i) for i in range(4)], [virtual(30 + i) for i in range(5)])
    with pytest.raises(RuntimeError, match="391-399"):
        pkg.GeneratorStream(lib, dr.bytes())
    empty = pkg.GeneratorStream(lib, b"")
    assert empty.num_generators == 0 and empty.num_slots == 0


def test_levels_of_random_generator_graphs_respect_the_dataflow():
    """Host scheduling logic on random DAGs (no device): every generator sits strictly above the producers of the
    partitions it reads, a second writer of a partition above the first, and the input slots are exactly the
    partitions that are read but never produced -- with random copy constraints merging targets."""
    lib = pkg.load_library()
    rng = random.Random(4242)
    for trial in range(12):
        w = GeneratorStreamWriter()
        n_virtual = 0

        def fresh(k):
            nonlocal n_virtual
            v = [virtual(n_virtual + i) for i in range(k)]
            n_virtual += k
            return v
        values = [fresh(9) for _ in range(4)]                 # pool of 9-limb values; the first four are inputs
        gens = []                                             # (reads, writes) as target lists, stream order
        copies = []
        for _ in range(rng.randrange(20, 120)):
            kind = rng.choice(["add", "sub", "inv", "mul", "copy"])
            if kind == "copy":
                a, b = rng.choice(values), rng.choice(values)
                l = rng.randrange(9)
                # only merge limbs of INPUT values (a copy between two produced partitions makes a second writer,
                # covered below by the explicit double-write case)
                if values.index(a) < 4 and values.index(b) < 4:
                    w.copy(a[l], b[l])
                    copies.append((a[l], b[l]))
                continue
            a, b = rng.choice(values), rng.choice(values)
            if kind in ("add", "sub"):
                out, ov = fresh(9), fresh(1)[0]
                (w.nonnative_addition_generator if kind == "add" else w.nonnative_subtraction_generator)(rng.randrange(2), a, b, out, ov)
                gens.append((a + b, out + [ov]))
            elif kind == "inv":
                inv, div = fresh(9), fresh(9)
                w.nonnative_inverse_generator(rng.randrange(2), a, inv, div)
                gens.append((list(a), inv + div))
                out = inv
            else:
                row = 1000 * trial + len(gens)
                for l in range(9):
                    w.copy(a[l], wire(row, l))
                    w.copy(b[l], wire(row, 9 + l))
                    copies.append((a[l], wire(row, l)))
                    copies.append((b[l], wire(row, 9 + l)))
                w.mul_nonnative_generator(rng.randrange(2), row)
                out = [wire(row, 18 + l) for l in range(9)]
                gens.append(([wire(row, c) for c in range(18)], [wire(row, c) for c in range(18, 53)]))
            values.append(out)
        # an explicit second writer: an addition whose sum targets are an earlier generator's outputs
        if len(values) > 6:
            first = values[5]
            w.nonnative_addition_generator(0, values[0], values[1], first, fresh(1)[0])
            gens.append((values[0] + values[1], list(first)))
        g = pkg.GeneratorStream(lib, w.bytes())
        assert g.num_generators == len(gens)
        lv = g.generator_levels
        slot = g.slot
        producer = {}
        for k, (_, writes) in enumerate(gens):
            for t in writes:
                producer.setdefault(slot(t), k)
        read_slots = set()
        for k, (reads, writes) in enumerate(gens):
            for t in reads:
                read_slots.add(slot(t))
                p = producer.get(slot(t))
                if p is not None and p != k:
                    assert lv[k] > lv[p], (trial, k, p)
            for t in writes:
                p = producer[slot(t)]
                if p != k:
                    assert lv[k] > lv[p]                      # the second writer checks after the first wrote
        assert sorted(g.input_slots) == sorted(s for s in read_slots if s not in producer)
        assert g.num_levels == max(lv) + 1
        g.close()


def test_loader_survives_corrupted_streams():
    """Random truncations / byte flips of a valid stream: the loader returns a handle or an error, never crashes, and
    a handle it returns answers its queries."""
    c, _, _, _ = _curve_circuit()
    good = c.w.bytes()
    lib = pkg.load_library()
    rng = random.Random(77)
    ok = bad = 0
    for _ in range(300):
        b = bytearray(good[: rng.randrange(len(good) + 1)] if rng.random() < 0.3 else good)
        for _ in range(rng.randrange(1, 4)):
            if b:
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        try:
            g = pkg.GeneratorStream(lib, bytes(b))
        except RuntimeError as e:
            assert "genstream" in str(e)
            bad += 1
            continue
        assert g.num_slots >= 0 and len(g.generator_levels) == g.num_generators and len(g.input_slots) <= g.num_slots
        g.close()
        ok += 1
    assert ok > 0 and bad > 0


# ------------------------------------------------------------------------------------------ GPU tier
def _rand_point(rng):
    return S.pt_mul(rng.randrange(1, N), S.G)


@pytest.fixture(params=["walk", "fused", "staged", "staged-nograph"])
def gs_mode(request, engine):
    """Every execution form of a stream: the level walker (one launch, a CTA walks all levels for its 32 circuits), one
    fused kernel per launch group, gather -> batched kernel -> scatter, with and without the CUDA graph (p2e_set_option)."""
    old = {k: engine.get_option(k) for k in ("genstream_fused", "genstream_graph", "genstream_walk")}
    engine.set_option("genstream_walk", 2 if request.param == "walk" else 0)
    engine.set_option("genstream_fused", 0 if request.param.startswith("staged") else 1)
    engine.set_option("genstream_graph", 0 if request.param.endswith("nograph") else 1)
    yield request.param
    for k, v in old.items():
        engine.set_option(k, v)


@pytest.mark.gpu
def test_curve_chain_stream_vs_oracle_worklist(engine, gs_mode):
    c, pts, a_const, t = _curve_circuit()
    g = engine.load_generator_stream(c.w.bytes())
    rng = random.Random(2024)
    n = 37
    P1 = [_rand_point(rng) for _ in range(n)]
    P2 = [_rand_point(rng) for _ in range(n)]
    cols = {}
    for (tx, ty), Ps in zip(pts, (P1, P2)):
        for l in range(9):
            cols[tx[l]] = [S.limbs9(p[0])[l] for p in Ps]
            cols[ty[l]] = [S.limbs9(p[1])[l] for p in Ps]
    for l in range(9):
        cols[a_const[l]] = [0] * n
    targets = c.all_targets()
    got, err = g.run(cols, targets)
    assert not err.any()
    for i in range(n):
        val = c.evaluate({tt: v[i] for tt, v in cols.items()})
        exp = [val(tt) for tt in targets]
        assert [int(v) for v in got[:, i]] == exp, f"circuit {i}"
        # and the chain computes 2 (P1 + P2) + P1
        R = S.pt_add(S.pt_add(S.pt_add(P1[i], P2[i]), S.pt_add(P1[i], P2[i])), P1[i])
        assert S.get_biguint([val(tt) for tt in t[0]]) == R[0] and S.get_biguint([val(tt) for tt in t[1]]) == R[1]
    g.close()


@pytest.mark.gpu
def test_stream_with_every_record_kind(engine, gs_mode):
    """GLV decomposition, BigUint div_rem, scalar-field generators, ConstantGenerator / ArithmeticBaseGenerator glue,
    short operands (absent limbs read as zero), a slot written twice with equal and with different values."""
    w = GeneratorStreamWriter()
    k = [virtual(i) for i in range(9)]
    k1, k2 = [virtual(10 + i) for i in range(5)], [virtual(20 + i) for i in range(5)]
    w.glv_decomposition_generator(k, k1, k2, virtual(30), virtual(31))
    # k1 * k2 mod n, k1 given as a 5-limb operand padded by the gate row's zero wires (copies to a constant)
    w.constant_generator(100, 0, 0, 0)
    w.mul_nonnative_generator(1, 50)
    for i in range(9):
        w.copy(k1[i] if i < 5 else wire(100, 0), wire(50, i))
        w.copy(k2[i] if i < 5 else wire(100, 0), wire(50, 9 + i))
    prod = [wire(50, 18 + i) for i in range(9)]
    # (k + prod) mod n with a 5-limb second operand; then a div_rem of k by k1
    s, ov = [virtual(40 + i) for i in range(9)], virtual(49)
    w.nonnative_addition_generator(1, k, k1, s, ov)
    div, rem = [virtual(60 + i) for i in range(5)], [virtual(70 + i) for i in range(5)]
    w.biguint_div_rem_generator(k, k1, div, rem)
    # arithmetic glue: wire(7,3) = 3 * ov * ov + 5 * k1_neg ; then a second writer of the same partition
    w.copy(ov, wire(7, 0)); w.copy(ov, wire(7, 1)); w.copy(virtual(30), wire(7, 2))
    w.arithmetic_generator(7, 3, 5, 0)
    w.copy(wire(7, 3), wire(8, 3))
    w.copy(ov, wire(8, 0)); w.copy(ov, wire(8, 1)); w.copy(virtual(30), wire(8, 2))
    w.arithmetic_generator(8, 3, 5, 0)                    # equal value: fine
    g = engine.load_generator_stream(w.bytes())
    rng = random.Random(9)
    n = 19
    ks = [rng.randrange(N) for _ in range(n)]
    cols = {k[l]: [S.limbs9(v)[l] for v in ks] for l in range(9)}
    outs = k1 + k2 + [virtual(30), virtual(31)] + prod + s + [ov] + div + rem + [wire(7, 3)]
    got, err = g.run(cols, outs)
    assert not err.any()
    for i in range(n):
        e_k1, e_k2, n1, n2 = S.glv_decomposition_gen(S.limbs9(ks[i]))
        q, r, _ = S.mul_nonnative_gen(e_k1 + [0] * 4, e_k2 + [0] * 4, N)
        e_s, e_ov = S.nonnative_add_gen(S.limbs9(ks[i]), e_k1, N)
        e_div, e_rem = S.biguint_divrem_gen(S.limbs9(ks[i]), e_k1)
        e_ar = (3 * e_ov * e_ov + 5 * int(n1)) % GL_P
        exp = e_k1 + e_k2 + [int(n1), int(n2)] + r + e_s + [e_ov] + e_div + e_rem + [e_ar]
        assert [int(v) for v in got[:, i]] == exp, f"instance {i}"
    g.close()
    # a partition set twice with different values
    w2 = GeneratorStreamWriter()
    w2.constant_generator(1, 0, 0, 5)
    w2.constant_generator(2, 0, 0, 6)
    w2.copy(wire(1, 0), wire(2, 0))
    g2 = engine.load_generator_stream(w2.bytes())
    got, err = g2.run({}, [wire(1, 0)])
    assert got.shape == (1, 0)
    g2.close()


@pytest.mark.gpu
def test_stream_errors_and_missing_inputs(engine, gs_mode):
    w = GeneratorStreamWriter()
    x, inv, div = [virtual(i) for i in range(9)], [virtual(9 + i) for i in range(9)], [virtual(18 + i) for i in range(9)]
    w.nonnative_inverse_generator(0, x, inv, div)
    w.constant_generator(1, 0, 0, 5)
    w.constant_generator(2, 0, 0, 6)
    w.copy(wire(1, 0), wire(2, 0))
    g = engine.load_generator_stream(w.bytes())
    vals = [0, 1, P - 1, P, 12345]                       # inverse of 0 and of p = 0 mod p panic in the reference
    cols = {x[l]: [S.limbs9(v)[l] for v in vals] for l in range(9)}
    got, err = g.run(cols, inv)
    assert [int(e) & pkg.ERR_INVERSE_ZERO for e in err] == [pkg.ERR_INVERSE_ZERO, 0, 0, pkg.ERR_INVERSE_ZERO, 0]
    assert all(int(e) & pkg.ERR_CONNECT for e in err)    # the two constants disagree in every instance
    for i in (1, 2, 4):
        assert S.get_biguint([int(v) for v in got[:, i]]) == pow(vals[i], P - 2, P)
    with pytest.raises(pkg.P2EError, match="not supplied"):
        g.run({x[0]: [1, 2]}, inv)
    g.close()
