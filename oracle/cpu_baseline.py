"""CPU ORACLE driver (test infrastructure / CPU baseline): runs oracle/p2e_oracle.c -- the C
restatement of the reference's CPU generators and gate evaluators -- on all host cores over a
bounded sample of a bench workload.  Used only by bench.py (cpu_baseline leg and
`--impl reference`) and tests/.  The Rust reference cannot be built here, hence kind == "port"."""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libp2e_oracle.so")
_u64p = ctypes.c_void_p
_sz = ctypes.c_size_t


def load():
    src = os.path.join(HERE, "p2e_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    lib = ctypes.CDLL(SO)
    lib.orc_mul_nonnative.argtypes = [ctypes.c_int, _sz, _sz, _sz, _u64p, _u64p, _u64p, _u64p]
    lib.orc_inv_nonnative.argtypes = [ctypes.c_int, _sz, _sz, _sz, _u64p, _u64p, _u64p]
    lib.orc_eval_mul_nonnative.argtypes = [ctypes.c_int, _sz, _sz, _sz, _u64p, _u64p]
    lib.orc_eval_check_sum.argtypes = [_sz, _sz, _sz, _u64p, _u64p]
    lib.orc_ecdsa_witness.argtypes = [_sz, _sz, _sz] + [_u64p] * 9
    lib.orc_set_inverse_mode.argtypes = [ctypes.c_int]
    lib.orc_set_inverse_mode.restype = None
    for f in ("orc_mul_nonnative", "orc_inv_nonnative", "orc_eval_mul_nonnative", "orc_eval_check_sum", "orc_ecdsa_witness"):
        getattr(lib, f).restype = None
    return lib


def _parallel(fn, n, threads):
    """fn(i0, i1) over [0, n) split across `threads` OS threads (ctypes releases the GIL)."""
    if threads <= 1:
        fn(0, n)
        return
    cuts = [n * t // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda t: fn(cuts[t], cuts[t + 1]), range(threads)))


def mul_nonnative(lib, modulus, x, y, threads=1):
    n = x.shape[1]
    out = np.empty((51, n), dtype=np.uint64)
    err = np.empty(n, dtype=np.uint32)
    _parallel(lambda a, b: lib.orc_mul_nonnative(modulus, n, a, b, x.ctypes.data, y.ctypes.data, out.ctypes.data, err.ctypes.data), n, threads)
    return out, err


def inv_nonnative(lib, modulus, x, threads=1):
    n = x.shape[1]
    out = np.empty((18, n), dtype=np.uint64)
    err = np.empty(n, dtype=np.uint32)
    _parallel(lambda a, b: lib.orc_inv_nonnative(modulus, n, a, b, x.ctypes.data, out.ctypes.data, err.ctypes.data), n, threads)
    return out, err


def eval_mul_nonnative(lib, modulus, w, threads=1):
    n = w.shape[1]
    out = np.empty((17, n), dtype=np.uint64)
    _parallel(lambda a, b: lib.orc_eval_mul_nonnative(modulus, n, a, b, w.ctypes.data, out.ctypes.data), n, threads)
    return out


def eval_check_sum(lib, w, threads=1):
    n = w.shape[1]
    out = np.empty((17, n), dtype=np.uint64)
    _parallel(lambda a, b: lib.orc_eval_check_sum(n, a, b, w.ctypes.data, out.ctypes.data), n, threads)
    return out


ECDSA_SLOTS = 82615
_ECDSA_CONSTS = None


def ecdsa_constants():
    """Circuit constants of verify_secp256k1_message_circuit, computed once by the Python spec (the
    reference computes them at circuit-build time): see the layout comment in p2e_oracle.c."""
    global _ECDSA_CONSTS
    if _ECDSA_CONSTS is None:
        from oracle import spec as S
        l9 = S.limbs9
        rando = S.pt_mul(S.blinding_scalar(), S.G)
        neg = S.pt_neg(rando)
        spm = rando
        for _ in range(146):
            spm = S.pt_add(spm, spm)
        nspm = S.pt_neg(spm)
        c = []
        for pt in (rando, neg, nspm):
            c += l9(pt[0]) + l9(pt[1])
        for v in (S.GLV_BETA, S.GLV_S, 2 * S.GLV_B2, 2 * S.GLV_MINUS_B1, S.GLV_A1, S.GLV_MINUS_B1, S.GLV_A2):
            c += l9(v)
        table = []
        for w in range(66):
            for pt in S.fixed_base_window(S.G, w):
                table += l9(pt[0]) + l9(pt[1])
        _ECDSA_CONSTS = (np.array(c, dtype=np.uint32), np.array(table, dtype=np.uint32))
    return _ECDSA_CONSTS


def ecdsa_witness(lib, msg, r, s, pkx, pky, threads=1):
    n = msg.shape[1]
    consts, table = ecdsa_constants()
    out = np.empty((ECDSA_SLOTS, n), dtype=np.uint64)
    err = np.empty(n, dtype=np.uint32)
    ins = [np.ascontiguousarray(a, dtype=np.uint64) for a in (msg, r, s, pkx, pky)]
    _parallel(lambda a, b: lib.orc_ecdsa_witness(n, a, b, *[x.ctypes.data for x in ins], consts.ctypes.data,
                                                 table.ctypes.data, out.ctypes.data, err.ctypes.data), n, threads)
    return out, err


def run(workload: str, steps: int = 1, warmup: int = 0, budget_s: float = 15.0):
    """Times the CPU port on a bounded sample of `workload`; returns a dict for bench.py."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from plonky2_ecdsa_b200 import workload as W
    lib = load()
    cores = os.cpu_count() or 1
    if workload == "mulmod":
        full = 1 << 20
        x, y = W.mulmod_workload(full, 0)
        fn = lambda xs, ys: mul_nonnative(lib, 0, xs, ys, cores)       # noqa: E731
        # calibrate on 2^14, then size the sample to the budget (capped at the full workload)
        t0 = time.perf_counter(); fn(np.ascontiguousarray(x[:, :1 << 14]), np.ascontiguousarray(y[:, :1 << 14])); dt = time.perf_counter() - t0
        n = int(min(full, max(1 << 14, (budget_s / max(steps + warmup, 1)) / dt * (1 << 14))))
        xs, ys = np.ascontiguousarray(x[:, :n]), np.ascontiguousarray(y[:, :n])
        for _ in range(warmup):
            fn(xs, ys)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(xs, ys)
        dt = (time.perf_counter() - t0) / steps
        return {"metric": "nonnative_mulmod_witnesses_per_s", "unit": "mulmod/s", "value": n / dt, "ms_per_step": dt * 1e3,
                "cores": cores, "kind": "port", "dtype": "u32 words / u64 (int)",
                "sample": f"first {n} of the 2^20 mulmod pairs per step; C port of BigUint mul + div_rem + convert_base + Goldilocks check_sum/carries (oracle/p2e_oracle.c), {cores} threads",
                "config": {"workload": "2^20 secp256k1 base-field nonnative mulmod witnesses (29-bit limbs): q, r, check_sum, carries",
                           "sample_units_per_step": n}}
    if workload == "eval":
        full = 1 << 20
        w = W.random_goldilocks(W.SEED, 53, full)
        fn = lambda ws: eval_mul_nonnative(lib, 0, ws, cores)          # noqa: E731
        t0 = time.perf_counter(); fn(np.ascontiguousarray(w[:, :1 << 14])); dt = time.perf_counter() - t0
        n = int(min(full, max(1 << 14, (budget_s / max(steps + warmup, 1)) / dt * (1 << 14))))
        ws = np.ascontiguousarray(w[:, :n])
        for _ in range(warmup):
            fn(ws)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(ws)
        dt = (time.perf_counter() - t0) / steps
        return {"metric": "mul_nonnative_gate_rows_per_s", "unit": "rows/s", "value": n / dt, "ms_per_step": dt * 1e3,
                "cores": cores, "kind": "port", "dtype": "u64 (Goldilocks)",
                "sample": f"first {n} of the 2^20 LDE rows per step; C port of eval_unfiltered per point, {cores} threads",
                "config": {"workload": "MulNonnativeGate eval_unfiltered_base_batch over 2^20 Goldilocks LDE rows",
                           "sample_units_per_step": n}}
    if workload == "ecdsa":
        full = 4096
        ecdsa_constants()
        cap = 1024
        d = W.ecdsa_workload(cap)
        ins = [d[k] for k in ("msg", "r", "s", "pkx", "pky")]
        fn = lambda m: ecdsa_witness(lib, *[np.ascontiguousarray(a[:, :m]) for a in ins], threads=cores)   # noqa: E731
        m0 = max(cores, 8)
        t0 = time.perf_counter(); fn(m0); dt = time.perf_counter() - t0
        n = int(min(cap, max(m0, (budget_s / max(steps + warmup, 1)) / dt * m0)))
        n = max(cores, n // cores * cores)
        for _ in range(warmup):
            fn(n)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(n)
        dt = (time.perf_counter() - t0) / steps
        # "best-CPU" variant: the same port with every Fermat ladder replaced by an extended-Euclid
        # inverse -- same witness bits (the inverse is unique), ~10x fewer word operations per inverse
        lib.orc_set_inverse_mode(1)
        try:
            n2 = min(cap, n * 4) // cores * cores
            fn(n2)
            t0 = time.perf_counter()
            fn(n2)
            dt2 = time.perf_counter() - t0
        finally:
            lib.orc_set_inverse_mode(0)
        fast = {"value": n2 / dt2, "unit": "witnesses/s", "cores": cores, "kind": "port-fast-inverse",
                "sample": f"first {n2} of the 4096 signatures; the C port with extended-Euclid inverses instead of the reference's Fermat ladders, {cores} threads"}
        return {"metric": "ecdsa_circuit_witnesses_per_s", "unit": "witnesses/s", "value": n / dt, "ms_per_step": dt * 1e3, "batched": fast,
                "cores": cores, "kind": "port", "dtype": "u32 words / u64 (int)",
                "sample": f"first {n} of the 4096 signatures per step; C port of the reference's generators, one Fermat ladder per curve op (oracle/p2e_oracle.c), {cores} threads",
                "config": {"workload": "4096 independent verify_secp256k1_message_circuit witnesses per GPU (full tape: 82615 slots each)",
                           "sample_units_per_step": n}}
    raise ValueError(f"no CPU baseline for workload {workload!r}")
