"""SURVEY 8f-2 closed over a real circuit: the complete verify_secp256k1_message_circuit (gadgets/ecdsa.rs:30-53) as a
serialized generator stream -- in-tree generators plus the plonky2 glue between them (constants, arithmetic, base-2
splits, random accesses, equality; restated, unpinned) -- run by the level-scheduled executor, against the chain
kernels' witness of the same signatures, slot for slot."""
import numpy as np
import pytest

import plonky2_ecdsa_b200 as pkg
from oracle import spec as S
from plonky2_ecdsa_b200 import genstream as G
from plonky2_ecdsa_b200 import workload as W


@pytest.fixture(scope="module")
def circuit():
    b = G.EcdsaCircuitBuilder(S.blinding_scalar())
    ins = b.verify_secp256k1_message()
    return b, ins


def test_circuit_stream_shape(circuit):
    """CPU tier: the stream parses, its in-tree generators are the tape's (SURVEY 3.1 counts), every tape slot has a target."""
    b, _ = circuit
    lib = pkg.load_library()
    gs = G.GeneratorStream(lib, b.bytes())
    try:
        assert len(b.tape_targets) == lib.p2e_ecdsa_tape_slots() == 82615
        assert len(gs.input_slots) == 45                       # msg, r, s, pk.x, pk.y: 5 x 9 limbs
        # 4642 in-tree generators + glue; the curve chain makes the stream deep
        assert gs.num_generators > 4642 + 139 * 18 and gs.num_levels > 311 * 4
        assert constant_limbs_ok()
    finally:
        gs.close()


def test_library_blinding_scalar_is_the_oracles():
    from plonky2_ecdsa_b200 import engine
    assert engine.blinding_scalar() == S.blinding_scalar()


def constant_limbs_ok():
    for v in (0, 1, 7, S.SECP_P - 1, S.GLV_S, S.GLV_BETA, 1 << 231, (1 << 232) - 1, 1 << 232):
        if G.constant_limbs(v) != S.convert_base(S.to_u32_digits(v), 32, 29):
            return False
    return True


@pytest.mark.gpu
@pytest.mark.parametrize("walk,n", [(1, 24), (1, 77), (0, 24)])
def test_circuit_stream_equals_the_tape_witness(engine, circuit, walk, n):
    """walk = 1: the level walker (what a stream of 1 862 levels runs by default: one launch); walk = 0: one fused kernel per
    launch group in a CUDA graph.  n = 77: a ragged last CTA of the walker (32 circuits per CTA)."""
    b, (msg, r, s, pk) = circuit
    old = engine.get_option("genstream_walk")
    engine.set_option("genstream_walk", walk)
    gs = engine.load_generator_stream(b.bytes())
    wl = W.ecdsa_workload(n, seed=2718)
    vals = {"msg": wl["msg"].copy(), "r": wl["r"].copy(), "s": wl["s"].copy(), "pkx": wl["pkx"].copy(), "pky": wl["pky"].copy()}
    vals["r"][0, 5] ^= np.uint64(1)             # a wrong signature: every generator still has a witness, the final connect fails
    inputs = {}
    for tv, key in ((msg, "msg"), (r, "r"), (s, "s"), (pk[0], "pkx"), (pk[1], "pky")):
        for l, t in enumerate(tv):
            inputs[t] = vals[key][l]
    got, err = gs.run(inputs, b.tape_targets)
    want, werr = engine.verify_secp256k1_message_witness(vals["msg"], vals["r"], vals["s"], vals["pkx"], vals["pky"])
    # connect_nonnative(r, point.x) (gadgets/ecdsa.rs:48-52) makes r and the last curve_add's x3 ONE partition: for the wrong
    # signature the stream keeps the input's value there and flags the mismatch (plonky2: "partition set twice with
    # different values", generation stops); everything generated before that point still equals the tape's witness
    x3 = 82615 - 231 + 150
    ok = np.ones(n, dtype=bool)
    ok[5] = False
    assert np.array_equal(got[:, ok], want[:, ok])
    assert np.array_equal(got[:x3, 5], want[:x3, 5])
    assert np.array_equal(got[x3:x3 + 9, 5], vals["r"][:, 5]) and not np.array_equal(want[x3:x3 + 9, 5], vals["r"][:, 5])
    assert np.array_equal(err, werr) and err[5] == pkg.ERR_CONNECT and not np.delete(err, 5).any()
    gs.close()
    engine.set_option("genstream_walk", old)
