"""Diagnostic: pipelined (S streams) time per batch of subsets of the chain kernels (option chain_phases; rows of such
runs are flagged ERR_INTERNAL).  Usage: python scripts/chain_phases.py [streams ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import workload as W

eng = pkg.Engine(0)
lib = eng.lib
n = 4096
_, nn, nw = eng.tape_compact_layout(eng.TAPE_ECDSA)
d = W.ecdsa_workload(n, seed=W.SEED)
ind = [torch.from_numpy(d[k].view(np.int64)).cuda() for k in ("msg", "r", "s", "pkx", "pky")]
ptrs = [t.data_ptr() for t in ind]
SMAX = 32
streams = [torch.cuda.Stream() for _ in range(SMAX)]
nars = [torch.empty((nn, n), dtype=torch.int32, device="cuda") for _ in range(SMAX)]
eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(SMAX)]


def run(S, K):
    for k in range(K):
        j = k % S
        assert lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[j].data_ptr(), None, n, eds[j].data_ptr(), streams[j].cuda_stream) == 0


def timed(S, K):
    run(S, 2 * S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for w in streams[:S]:
        w.wait_event(e0)
    run(S, K)
    for w in streams[:S]:
        ev = torch.cuda.Event()
        ev.record(w)
        torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


run(SMAX, SMAX)          # full run first: every arena holds valid scratch for the partial runs
torch.cuda.synchronize()
masks = [("all", 63), ("prep", 1), ("replay+join", 2), ("invert", 4), ("prep+replay+invert", 7), ("affine", 8), ("expand", 16),
         ("affine+expand", 24), ("replay+expand", 18), ("all but prep", 62), ("all but expand", 47)]
for S in [int(a) for a in sys.argv[1:]] or [8]:
    for name, m in masks:
        eng.set_option("chain_phases", m)
        print("streams %2d  %-20s %.4f ms per batch" % (S, name, timed(S, 12 * S)), flush=True)
eng.set_option("chain_phases", 63)
eng.close()
