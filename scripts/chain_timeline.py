"""Timeline of pipelined ECDSA batches (diagnostic, never a bench number): S streams x K steps with the chain trace on
(p2e_set_option("chain_trace", 1)); prints, per call, when each kernel finished and how long each phase took under
contention.  Usage: python scripts/chain_timeline.py [streams] [steps] [split] [n]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import workload as W

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
SPLIT = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = pkg.Engine(0)
lib = eng.lib
eng.set_option("chain_split", SPLIT)
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
_, nn, nw = eng.tape_compact_layout(eng.TAPE_ECDSA)
d = W.ecdsa_workload(n, seed=W.SEED)
ind = [torch.from_numpy(d[k].view(np.int64)).cuda() for k in ("msg", "r", "s", "pkx", "pky")]
ptrs = [t.data_ptr() for t in ind]
streams = [torch.cuda.Stream() for _ in range(S)]
nars = [torch.empty((nn, n), dtype=torch.int32, device="cuda") for _ in range(S)]
eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(S)]


def step(k):
    j = k % S
    assert lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[j].data_ptr(), None, n, eds[j].data_ptr(), streams[j].cuda_stream) == 0


for k in range(2 * S):
    step(k)
torch.cuda.synchronize()
eng.set_option("chain_trace", 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for w in streams:
    w.wait_event(e0)
for k in range(K):
    step(k)
for w in streams:
    ev = torch.cuda.Event()
    ev.record(w)
    torch.cuda.current_stream().wait_event(ev)
e1.record()
torch.cuda.synchronize()
print("streams %d steps %d split %d: %.4f ms per step" % (S, K, SPLIT, e0.elapsed_time(e1) / K))
cap = 16 * K + 64
calls = np.zeros(cap, dtype=np.int32)
kern = np.zeros(cap, dtype=np.int32)
ms = np.zeros(cap, dtype=np.float32)
cnt = lib.p2e_chain_trace_dump(calls.ctypes.data, kern.ctypes.data, ms.ctypes.data, cap)
eng.set_option("chain_trace", 0)
t = {}
for i in range(cnt):
    t.setdefault(int(calls[i]), {})[int(kern[i])] = float(ms[i])
names = ["start", "prep", "replay", "join", "invert", "gate", "affine", "expand"]
print("call  " + " ".join("%8s" % x for x in names) + "   | phase durations: prep replay join invert wait affine expand")
allc = [c for c in sorted(t) if len(t[c]) >= 8]
dus = []
for c in allc:
    row = t[c]
    ts = [row[k] for k in range(8)]
    du = [ts[k] - ts[k - 1] for k in range(1, 8)]
    dus.append(du)
    if len(allc) > 48 and c not in allc[-48:]:
        continue
    print("%4d  " % c + " ".join("%8.3f" % x for x in ts) + "   | " + " ".join("%6.3f" % x for x in du))
m = np.array(dus)[len(dus) // 4:]
print("mean phase durations (ms): " + " ".join("%s %.3f" % (a, b) for a, b in zip(names[1:], m.mean(axis=0))) + "  | first->last kernel %.3f" % m.sum(axis=1).mean())
eng.close()
