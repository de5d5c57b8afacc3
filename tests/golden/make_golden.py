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
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f)
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
