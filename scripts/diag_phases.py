"""Diagnostic: pipelined (S streams) time per batch of subsets of the chain kernels (P2E_CHAIN_PHASES)."""
import os, sys, ctypes
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
S = int(os.environ.get("S", "4"))
streams = [torch.cuda.Stream() for _ in range(S)]
ods = [torch.empty((slots, n), dtype=torch.int64, device="cuda") for _ in range(S)]
eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(S)]
main = torch.cuda.Stream(); torch.cuda.set_stream(main)

def run(k):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for w in streams: w.wait_event(e0)
    for j in range(k):
        q = j % S
        assert lib.p2e_ecdsa_witness_dev(n, *ptrs, n, ods[q].data_ptr(), n, eds[q].data_ptr(), streams[q].cuda_stream) == 0
    for w in streams:
        ev = torch.cuda.Event(); ev.record(w); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

os.environ["P2E_CHAIN_PHASES"] = "63"
run(8)
for mask, name in [(63, "all"), (7, "prep+replay+invert"), (24, "affine+expand"), (2, "replay"), (1, "prep"), (4, "invert"), (16, "expand"), (8, "affine"), (3, "prep+replay"), (18, "replay+expand"), (22, "replay+invert+expand"), (63, "all")]:
    os.environ["P2E_CHAIN_PHASES"] = str(mask)
    run(4)
    print(f"S={S} mask={mask:2d} {name:24s} {run(40):.3f} ms/batch", flush=True)
os.environ["P2E_CHAIN_PHASES"] = "63"
