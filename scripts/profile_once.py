"""One invocation of each hot kernel, for `ncu` captures (never a bench number).
Usage: python scripts/profile_once.py ecdsa | ux | fc | lde | std"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import workload as W

what = sys.argv[1] if len(sys.argv) > 1 else "ecdsa"
eng = pkg.Engine(0); lib = eng.lib
dev = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
if what == "ecdsa":
    # one batch on the library stream with the per-kernel event marks on (keeps every kernel on ONE stream, in order)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096          # a multiple of 4096: the 4096-signature workload tiled
    _, nn, nw = eng.tape_compact_layout(eng.TAPE_ECDSA)
    d = W.ecdsa_workload(4096, seed=W.SEED)
    ind = [dev(np.ascontiguousarray(np.tile(d[k], (1, n // 4096)))) for k in ("msg", "r", "s", "pkx", "pky")]
    nar = torch.empty((nn, n), dtype=torch.int32, device="cuda")
    ed = torch.empty(n, dtype=torch.int32, device="cuda")
    eng.set_profiling(True)
    for _ in range(2):
        assert lib.p2e_ecdsa_witness_compact_dev(n, *[t.data_ptr() for t in ind], n, nar.data_ptr(), None, n, ed.data_ptr(), None) == 0
    eng.synchronize()
    assert not bool(ed.any())
elif what == "ux":
    n = 1 << 20
    for gate, bits, lb, count, na in [(0, 29, 3, 9, 0), (0, 34, 3, 8, 0), (1, 29, 2, 1, 0), (2, 29, 2, 5, 3), (3, 29, 2, 6, 0), (4, 29, 2, 3, 0)]:
        nw, nc, ni = eng.ux_gate_shape(gate, bits, lb, count, na)
        w = dev(W.random_goldilocks(W.SEED + gate, nw, n))
        o = torch.empty((nc, n), dtype=torch.int64, device="cuda")
        for _ in range(2):
            assert lib.p2e_eval_ux_dev(gate, bits, lb, count, na, n, w.data_ptr(), n, o.data_ptr(), n, None) == 0
        x = dev(W.random_goldilocks(W.SEED + 9, ni, n) & np.uint64((1 << bits) - 1))
        row = torch.empty((nw, n), dtype=torch.int64, device="cuda")
        e = torch.empty(n, dtype=torch.int32, device="cuda")
        assert lib.p2e_ux_witness_dev(gate, bits, lb, count, na, n, x.data_ptr(), n, row.data_ptr(), n, e.data_ptr(), None) == 0
        eng.synchronize()
        del w, o, x, row
elif what == "fc":
    import ctypes
    n = 1 << 20
    w = dev(W.random_goldilocks(W.SEED + 5, 100, n))          # 99 wires (widest gate here) + a selector column
    o = torch.empty((2, n), dtype=torch.int64, device="cuda")
    al = np.array([0x0123456789ABCDEF, 0x0FEDCBA987654321], dtype=np.uint64)
    sel = w.data_ptr() + 99 * n * 8
    for gate, params in [(0, (0,)), (1, ()), (2, (0, 29, 3, 9, 0))]:
        prm = (ctypes.c_int * 5)(*(list(params) + [0] * 5)[:5])
        scr = torch.empty(int(lib.p2e_eval_filtered_scratch_bytes(gate, ctypes.addressof(prm), n)) // 8 + 2, dtype=torch.int64, device="cuda")
        for _ in range(2):
            assert lib.p2e_eval_filtered_combined_dev(gate, ctypes.addressof(prm), 1, 0, 4, 1, n, w.data_ptr(), n, sel, al.ctypes.data, 2, 4, 0,
                                                      o.data_ptr(), n, scr.data_ptr(), None) == 0
        eng.synchronize()
elif what == "lde":
    log_n, rb, polys = 17, 3, 135
    n, m = 1 << log_n, 1 << (log_n + rb)
    vd = dev(W.random_goldilocks(W.SEED + 21, polys, n))
    od = torch.empty((polys, m), dtype=torch.int64, device="cuda")
    scr = torch.empty(int(lib.p2e_ntt_scratch_bytes(log_n, rb, polys)) // 8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        assert lib.p2e_coset_lde_dev(log_n, rb, polys, vd.data_ptr(), n, od.data_ptr(), m, scr.data_ptr(), None) == 0
    eng.synchronize()
else:
    n = 1 << 20
    x, y = W.mulmod_workload(n, 0, seed=W.SEED)
    xd, yd = dev(x), dev(y)
    od = torch.empty((pkg.MUL_OUT_ROWS, n), dtype=torch.int64, device="cuda")
    ed = torch.empty(n, dtype=torch.int32, device="cuda")
    w = dev(W.random_goldilocks(W.SEED, 53, n))
    o = torch.empty((17, n), dtype=torch.int64, device="cuda")
    for _ in range(2):
        assert lib.p2e_mul_nonnative_witness_dev(0, n, xd.data_ptr(), yd.data_ptr(), n, od.data_ptr(), n, ed.data_ptr(), None) == 0
        assert lib.p2e_eval_mul_nonnative_dev(0, n, w.data_ptr(), n, o.data_ptr(), n, None) == 0
        assert lib.p2e_eval_check_sum_dev(n, w.data_ptr(), n, o.data_ptr(), n, None) == 0
    eng.synchronize()
eng.close()
print("done", what)
