"""Diagnostic: do the latency-bound kernels (replay) and the full-GPU kernels (expand) co-run?
Stream 0 gets only replay-calls, stream 1 only expand-calls (P2E_CHAIN_PHASES flipped per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plonky2_ecdsa_b200 as pkg
from plonky2_ecdsa_b200 import workload as W

eng = pkg.Engine(0); lib = eng.lib
n = 4096
slots = int(lib.p2e_ecdsa_tape_slots())
d = W.ecdsa_workload(n, seed=W.SEED)
ind = [torch.from_numpy(d[k].view('int64')).cuda() for k in ("msg", "r", "s", "pkx", "pky")]
ptrs = [t.data_ptr() for t in ind]
S = 4
streams = [torch.cuda.Stream() for _ in range(S)]
ods = [torch.empty((slots, n), dtype=torch.int64, device="cuda") for _ in range(S)]
eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(S)]
main = torch.cuda.Stream(); torch.cuda.set_stream(main)

def call(q, mask):
    os.environ["P2E_CHAIN_PHASES"] = str(mask)
    assert lib.p2e_ecdsa_witness_dev(n, *ptrs, n, ods[q].data_ptr(), n, eds[q].data_ptr(), streams[q].cuda_stream) == 0

def run(plan, k):
    """plan: list of (stream, mask) issued round-robin k times"""
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for w in streams: w.wait_event(e0)
    for _ in range(k):
        for q, m in plan:
            call(q, m)
    for w in streams:
        ev = torch.cuda.Event(); ev.record(w); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

for q in range(S):
    call(q, 63)
torch.cuda.synchronize()
for name, plan in [
    ("replay on s0", [(0, 2)]),
    ("expand on s1", [(1, 16)]),
    ("replay s0 || expand s1", [(0, 2), (1, 16)]),
    ("replay s0,s2 || expand s1", [(0, 2), (2, 2), (1, 16)]),
    ("replay s0,s2,s3 || expand s1", [(0, 2), (2, 2), (3, 2), (1, 16)]),
    ("small(7) s0 || big(24) s1", [(0, 7), (1, 24)]),
    ("small s0,s2 || big s1", [(0, 7), (2, 7), (1, 24)]),
    ("small s0,s2,s3 || big s1", [(0, 7), (2, 7), (3, 7), (1, 24)]),
    ("invert s0 || expand s1", [(0, 4), (1, 16)]),
]:
    run(plan, 3)
    print(f"{name:32s} {run(plan, 20):.3f} ms per round", flush=True)
os.environ["P2E_CHAIN_PHASES"] = "63"
