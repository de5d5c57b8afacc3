"""Diagnostic: pipelined rate of small ECDSA batches on one GPU (the per-rank situation of strong scaling at N = 8: 512
signatures per call) and the host's enqueue cost per call.  Usage: python scripts/small_batch.py n streams [calls]"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import workload as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = int(sys.argv[3]) if len(sys.argv) > 3 else 16 * S
eng = pkg.Engine(0)
lib = eng.lib
for a in sys.argv[4:]:
    k, v = a.split("=")
    eng.set_option(k, int(v))
_, nn, nw = eng.tape_compact_layout(eng.TAPE_ECDSA)
d = W.ecdsa_workload(n, seed=W.SEED)
ind = [torch.from_numpy(d[k].view(np.int64)).cuda() for k in ("msg", "r", "s", "pkx", "pky")]
ptrs = [t.data_ptr() for t in ind]
streams = [torch.cuda.Stream() for _ in range(S)]
nars = [torch.empty((nn, n), dtype=torch.int32, device="cuda") for _ in range(S)]
eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(S)]


def run(K):
    for k in range(K):
        j = k % S
        assert lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[j].data_ptr(), None, n, eds[j].data_ptr(), streams[j].cuda_stream) == 0


run(2 * S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for w in streams:
    w.wait_event(e0)
t0 = time.perf_counter()
run(K)
t_host = time.perf_counter() - t0
for w in streams:
    ev = torch.cuda.Event()
    ev.record(w)
    torch.cuda.current_stream().wait_event(ev)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("n %d streams %d: %.4f ms per batch (%.4f ms per 4096 signatures), host enqueue %.1f us per call, errors %d" %
      (n, S, ms, ms * 4096 / n, t_host / K * 1e6, int(sum(int(e.any()) for e in eds))), flush=True)
eng.close()
