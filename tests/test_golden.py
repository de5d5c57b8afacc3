"""Golden fixtures (tests/golden/golden.json, made by tests/golden/make_golden.py):
CPU tier re-derives them with the oracle and the C port; GPU tier checks the CUDA path."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cpu_baseline as CB
from oracle import spec as S
from tests import util as U

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def arr(v):
    return np.array(v, dtype=np.uint64)


def test_golden_regenerates_from_oracle():
    for modulus in (0, 1):
        g = G[f"mul_nonnative_mod{modulus}"]
        out, err = U.expect_mul(arr(g["x"]), arr(g["y"]), modulus)
        assert out.tolist() == g["out"] and err.tolist() == g["err"]
    g = G["ecdsa_witness"]
    i = g["inputs"]
    flat = S.ecdsa_witness(i["msg"][0], i["r"][0], i["s"][0], (i["pkx"][0], i["pky"][0])).flat()
    assert hashlib.sha256(np.asarray(flat, dtype=np.uint64).tobytes()).hexdigest() == g["digests"][0]["sha256"]


def test_golden_c_port():
    lib = CB.load()
    for modulus in (0, 1):
        g = G[f"mul_nonnative_mod{modulus}"]
        out, err = CB.mul_nonnative(lib, modulus, arr(g["x"]), arr(g["y"]))
        assert out.tolist() == g["out"] and err.tolist() == g["err"]
    g = G["eval_mul_nonnative_mod0"]
    assert CB.eval_mul_nonnative(lib, 0, arr(g["wires"])).tolist() == g["out"]
    g = G["eval_check_sum"]
    assert CB.eval_check_sum(lib, arr(g["wires"])).tolist() == g["out"]


@pytest.mark.gpu
def test_golden_gpu(engine):
    for modulus in (0, 1):
        g = G[f"mul_nonnative_mod{modulus}"]
        out, err = engine.mul_nonnative(modulus, arr(g["x"]), arr(g["y"]))
        assert out.tolist() == g["out"] and err.tolist() == g["err"]
    g = G["eval_mul_nonnative_mod0"]
    assert engine.eval_mul_nonnative_gate(0, arr(g["wires"])).tolist() == g["out"]
    g = G["eval_check_sum"]
    assert engine.eval_check_sum_gate(arr(g["wires"])).tolist() == g["out"]
    g = G["glv_decompose"]
    out, err = engine.decompose_secp256k1_scalar(arr(g["k"]))
    assert out.tolist() == g["out"] and not err.any()


@pytest.mark.gpu
def test_golden_ecdsa_witness_gpu(engine):
    from plonky2_ecdsa_b200 import workload as W
    g = G["ecdsa_witness"]
    i = g["inputs"]
    ins = [W.ints_to_limbs(i[k]) for k in ("msg", "r", "s", "pkx", "pky")]
    wit, err = engine.verify_secp256k1_message_witness(*ins)
    assert not err.any()
    for j, d in enumerate(g["digests"]):
        col = np.ascontiguousarray(wit[:, j])
        assert col.shape[0] == d["slots"]
        assert [int(v) for v in col[:8]] == d["first8"] and [int(v) for v in col[-8:]] == d["last8"]
        assert hashlib.sha256(col.tobytes()).hexdigest() == d["sha256"]


def test_golden_generator_stream_bytes_are_what_the_builder_emits():
    """The committed stream is byte-for-byte what the product-side target-level builder writes for the same circuit."""
    import plonky2_ecdsa_b200 as pkg
    b = pkg.TargetCircuitBuilder()
    pts = [(b.add_virtual_nonnative_target(), b.add_virtual_nonnative_target()) for _ in range(2)]
    a_const = b.add_virtual_nonnative_target()
    b.curve_add(b.curve_double(b.curve_add(pts[0], pts[1]), a_const), pts[0])
    assert b.bytes().hex() == G["generator_stream"]["bytes_hex"]
    gs = pkg.GeneratorStream(pkg.load_library(), bytes.fromhex(G["generator_stream"]["bytes_hex"]))
    assert gs.num_generators == 41 and len(gs.input_slots) == 45
    gs.close()


@pytest.mark.gpu
def test_golden_filtered_combined_lde_and_generator_stream_gpu(engine):
    import plonky2_ecdsa_b200 as pkg
    fc = G["filtered_combined"]
    kw = dict(selector=fc["selector"], gate_index=fc["gate_index"], group_range=tuple(fc["group"]),
              many_selectors=fc["many_selectors"], first_power=fc["first_power"])
    got = engine.eval_filtered_combined(pkg.GATE_MUL_NONNATIVE, (0,), arr(G["eval_mul_nonnative_mod0"]["wires"]), fc["alphas"], **kw)
    assert got.tolist() == fc["mul_nonnative_mod0"]
    got = engine.eval_filtered_combined(pkg.GATE_CHECK_SUM, (), arr(G["eval_check_sum"]["wires"]), fc["alphas"], **kw)
    assert got.tolist() == fc["check_sum"]
    g = G["coset_lde"]
    assert engine.lde_onto_coset(arr(g["values"]), g["rate_bits"]).tolist() == g["out"]
    g = G["generator_stream"]
    gs = engine.load_generator_stream(bytes.fromhex(g["bytes_hex"]))
    ins = {tuple(k): v for k, v in g["inputs"]}
    targets = [tuple(t) for t in g["targets"]]
    got, err = gs.run(ins, targets)
    assert not err.any()
    for j, d in enumerate(g["sha256"]):
        assert hashlib.sha256(np.ascontiguousarray(got[:, j]).tobytes()).hexdigest() == d
    res, _ = gs.run(ins, [tuple(t) for t in g["result_targets"]])
    assert res.T.tolist() == g["result"]
    gs.close()
