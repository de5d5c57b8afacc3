"""Generates the committed golden fixtures from the oracle (oracle/spec.py).

The Rust reference cannot be executed in this image (no cargo/rustc, un-vendored git deps), so the
fixtures are ORACLE outputs on seeded inputs; the oracle itself is pinned by tests/test_oracle_kat.py.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np                                    # noqa: E402
from oracle import spec as S                          # noqa: E402
from plonky2_ecdsa_b200 import workload as W          # noqa: E402
from tests import util as U                           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def digest_u64(vals):
    return hashlib.sha256(np.asarray(vals, dtype=np.uint64).tobytes()).hexdigest()


def main():
    g = {}
    # A8 + A9: first 64 rows of the config-2 workload for both moduli, full outputs
    for modulus in (0, 1):
        x, y = W.mulmod_workload(64, modulus)
        out, err = U.expect_mul(x, y, modulus)
        g[f"mul_nonnative_mod{modulus}"] = {"x": x.tolist(), "y": y.tolist(), "out": out.tolist(), "err": err.tolist()}
    # A11 / A12 on 16 rows of seeded Goldilocks wires
    w = W.random_goldilocks(W.SEED, 53, 16)
    g["eval_mul_nonnative_mod0"] = {"wires": w.tolist(), "out": U.expect_eval_mul(w, 0).tolist()}
    w2 = W.random_goldilocks(W.SEED + 1, 33, 16)
    g["eval_check_sum"] = {"wires": w2.tolist(), "out": U.expect_eval_check_sum(w2).tolist()}
    # A10 on 32 seeded scalars
    k = W.random_field_limbs(W.SEED + 2, 32, 1)
    g["glv_decompose"] = {"k": k.tolist(), "out": U.expect_glv(k)[0].tolist()}
    # A13: full ECDSA tapes of the first 4 signatures of the config-3 workload, as SHA-256 digests
    wl = W.ecdsa_workload(4)
    i = wl["ints"]
    dig = []
    for j in range(4):
        flat = S.ecdsa_witness(i["msg"][j], i["r"][j], i["s"][j], (i["pkx"][j], i["pky"][j])).flat()
        dig.append({"sha256": digest_u64(flat), "slots": len(flat), "first8": flat[:8], "last8": flat[-8:]})
    g["ecdsa_witness"] = {"inputs": {k2: v for k2, v in i.items()}, "digests": dig}
    # 8f-3: selector filter + challenge fold of the two in-tree gates on the same 16 rows (spec: oracle/vanishing_spec.py)
    from oracle import vanishing_spec as V
    from oracle import ntt_spec as NT
    sel = [1, 0, 2, V.UNUSED_SELECTOR, 1, 1, 7, 1, 1, 1, 1, 1, 1, 1, 1, 1]
    alphas = [0x0123456789ABCDEF, V.GL_P - 3]
    fc = {"selector": sel, "alphas": alphas, "gate_index": 1, "group": [0, 3], "many_selectors": True, "first_power": 7}
    for name, wires, ev in (("mul_nonnative_mod0", w, lambda c: S.mul_nonnative_gate_eval(c, S.SECP_P)), ("check_sum", w2, S.check_sum_gate_eval)):
        cols = []
        for j in range(16):
            cons = ev([int(v) for v in wires[:, j]])
            cols.append(V.gate_terms_combined(cons, 1, (0, 3), sel[j], True, alphas, 7))
        fc[name] = np.array(cols, dtype=np.uint64).T.tolist()
    g["filtered_combined"] = fc
    # 8f-3: coset LDE of two 16-point polynomials, rate 8
    vals = W.random_goldilocks(W.SEED + 4, 2, 16)
    g["coset_lde"] = {"values": vals.tolist(), "rate_bits": 3, "out": [NT.lde_onto_coset([int(v) for v in row], 3) for row in vals]}
    # 8f-2 / 8f-4: a serialized generator stream (curve_add -> curve_double -> curve_add at the target level) and the
    # values a worklist evaluation by the oracle assigns to every target, for two input pairs
    from tests.test_genstream import _curve_circuit
    c, pts, a_const, t = _curve_circuit()
    targets = c.all_targets()
    P1 = [S.pt_mul(0x1234567 + 11 * j, S.G) for j in range(2)]
    P2 = [S.pt_mul(0x7654321 + 13 * j, S.G) for j in range(2)]
    ins = {}
    for (tx, ty), Ps in zip(pts, (P1, P2)):
        for l in range(9):
            ins[tx[l]] = [S.limbs9(p[0])[l] for p in Ps]
            ins[ty[l]] = [S.limbs9(p[1])[l] for p in Ps]
    for l in range(9):
        ins[a_const[l]] = [0, 0]
    digs, finals = [], []
    for j in range(2):
        val = c.evaluate({tt: v[j] for tt, v in ins.items()})
        digs.append(digest_u64([val(tt) for tt in targets]))
        finals.append([val(tt) for tt in t[0] + t[1]])
    g["generator_stream"] = {"bytes_hex": c.w.bytes().hex(), "inputs": [[list(k), v] for k, v in ins.items()],
                             "targets": [list(tt) for tt in targets], "sha256": digs,
                             "result_targets": [list(tt) for tt in t[0] + t[1]], "result": finals}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f)
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
