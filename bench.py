#!/usr/bin/env python
"""bench.py -- `python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]`

One "step" = one pass of the hot path over one batch of synthetic input:
  workload mulmod : 2^20 nonnative secp256k1 base-field mulmod witnesses (BASELINE.json configs[1])
  workload ecdsa  : 4096 verify_secp256k1_message_circuit witnesses per GPU (configs[2])
  workload eval   : MulNonnativeGate constraint evaluation over 2^20 LDE rows (configs[4])
`value` is whole-job throughput with inputs resident in HBM; `e2e` is the same metric through the
host-pointer C-ABI call (pinned host buffers, H2D + kernel + D2H inside the timed region).
N > 1: one process per GPU (torchrun), the batch is replicated per rank (weak scaling, no data-path
collective; a tiny NCCL all_reduce gathers the error counts at the end).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {"mulmod": 552, "eval": 560}        # algorithmic bytes per unit (u64 wires in)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                f = [t.strip() for t in o.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """Times the oracle's C restatement of the reference's CPU generators (the Rust crate cannot be
    built in this image: no cargo, un-vendored git deps) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    res = cpu_baseline.run(args.workload, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": res["metric"], "value": res["value"], "unit": res["unit"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic",
        "config": res["config"],
        "cpu_baseline": {"value": res["value"], "unit": res["unit"], "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": res["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm (GPU)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("P2E_BENCH_WORKLOAD", "mulmod"), choices=["mulmod", "ecdsa", "eval"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import plonky2_ecdsa_b200 as pkg
    from plonky2_ecdsa_b200 import workload as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = pkg.Engine(local_rank)
    lib = eng.lib
    st = torch.cuda.Stream()              # a real (non-NULL) stream: NULL would select the library's own
    torch.cuda.set_stream(st)
    stream_ptr = st.cuda_stream
    assert stream_ptr != 0
    peak_gbs, peak_kind = load_peaks()

    def dev(a):
        return torch.from_numpy(a.view(np.int64)).cuda()

    def pinned(a):
        t = torch.from_numpy(a.view(np.int64)).pin_memory()
        return t

    wl = args.workload
    launches_per_step = 1
    if wl == "mulmod":
        n = 1 << 20
        x, y = W.mulmod_workload(n, 0, seed=W.SEED + rank)
        xd, yd = dev(x), dev(y)
        od = torch.empty((pkg.MUL_OUT_ROWS, n), dtype=torch.int64, device="cuda")
        ed = torch.empty(n, dtype=torch.int32, device="cuda")

        def step():
            rc = lib.p2e_mul_nonnative_witness_dev(0, n, xd.data_ptr(), yd.data_ptr(), n, od.data_ptr(), n, ed.data_ptr(), stream_ptr)
            assert rc == 0, eng.last_error()
        hx, hy = pinned(x), pinned(y)
        ho = torch.empty((pkg.MUL_OUT_ROWS, n), dtype=torch.int64).pin_memory()
        he = torch.empty(n, dtype=torch.int32).pin_memory()

        def step_e2e():
            rc = lib.p2e_mul_nonnative_witness(0, n, hx.data_ptr(), hy.data_ptr(), ho.data_ptr(), he.data_ptr())
            assert rc == 0, eng.last_error()
        h2d, d2h = 2 * 9 * 8 * n, pkg.MUL_OUT_ROWS * 8 * n + 4 * n
        units = n
        metric, unit = "nonnative_mulmod_witnesses_per_s", "mulmod/s"
        workload_name = "2^20 secp256k1 base-field nonnative mulmod witnesses (29-bit limbs): q, r, check_sum, carries"
        alg_bytes = ALG_BYTES["mulmod"] * n
        kernel_name = "k_mul_witness<ModP>"
        dtype = "u32 limbs / u64 (int)"
    elif wl == "eval":
        n = 1 << 20
        w = W.random_goldilocks(W.SEED + rank, 53, n)
        wd = dev(w)
        od = torch.empty((17, n), dtype=torch.int64, device="cuda")

        def step():
            rc = lib.p2e_eval_mul_nonnative_dev(0, n, wd.data_ptr(), n, od.data_ptr(), n, stream_ptr)
            assert rc == 0, eng.last_error()
        hw = pinned(w)
        ho = torch.empty((17, n), dtype=torch.int64).pin_memory()

        def step_e2e():
            rc = lib.p2e_eval_mul_nonnative(0, n, hw.data_ptr(), ho.data_ptr())
            assert rc == 0, eng.last_error()
        h2d, d2h = 53 * 8 * n, 17 * 8 * n
        units = n
        metric, unit = "mul_nonnative_gate_rows_per_s", "rows/s"
        workload_name = "MulNonnativeGate eval_unfiltered_base_batch over 2^20 Goldilocks LDE rows"
        alg_bytes = ALG_BYTES["eval"] * n
        kernel_name = "k_eval_mul_nonnative<ModP>"
        dtype = "u64 (Goldilocks)"
    else:
        n = 4096
        slots = int(lib.p2e_ecdsa_tape_slots())
        if slots == 0:
            raise SystemExit("ecdsa workload: chain kernels not built")
        d = W.ecdsa_workload(n, seed=W.SEED + rank)
        ins = [d[k] for k in ("msg", "r", "s", "pkx", "pky")]
        ind = [dev(a) for a in ins]
        od = torch.empty((slots, n), dtype=torch.int64, device="cuda")
        ed = torch.empty(n, dtype=torch.int32, device="cuda")

        def step():
            rc = lib.p2e_ecdsa_witness_dev(n, *[t.data_ptr() for t in ind], n, od.data_ptr(), n, ed.data_ptr(), stream_ptr)
            assert rc == 0, eng.last_error()
        hin = [pinned(a) for a in ins]
        ho = torch.empty((slots, n), dtype=torch.int64).pin_memory()
        he = torch.empty(n, dtype=torch.int32).pin_memory()

        def step_e2e():
            rc = lib.p2e_ecdsa_witness(n, *[t.data_ptr() for t in hin], ho.data_ptr(), he.data_ptr())
            assert rc == 0, eng.last_error()
        h2d, d2h = 5 * 9 * 8 * n, slots * 8 * n + 4 * n
        units = n
        metric, unit = "ecdsa_circuit_witnesses_per_s", "witnesses/s"
        workload_name = "4096 independent verify_secp256k1_message_circuit witnesses per GPU (full tape)"
        alg_bytes = (slots + 45) * 8 * n
        kernel_name = "k_chain_expand"
        dtype = "u32 limbs / u64 (int)"
        launches_per_step = int(os.environ.get("P2E_ECDSA_LAUNCHES", "4"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)
    # keep the sampler alive for at least a few samples under load
    t_end = time.time() + 0.6
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join()
    # ---- end-to-end timing through the host-pointer C-ABI
    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    errs = int((ed != 0).sum().item()) if wl != "eval" else 0

    tms = torch.tensor([ms, e2e_ms, float(errs)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = tms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tms.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)         # the only collective: final gather of error counts
        ms, e2e_ms, errs = float(mx[0]), float(mx[1]), int(sm[2])
    if rank == 0:
        ms_per_step = ms / args.steps
        value = units * world / (ms_per_step * 1e-3)
        e2e_value = units * world / (e2e_ms / e2e_steps * 1e-3)
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload_name, "units_per_step_per_gpu": units, "seed": hex(W.SEED),
                       "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)" if wl != "ecdsa" else "outputs per step (2.7 GB) exceed L2",
                       "error_rows": errs},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "peak_kind": peak_kind, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes},
        }
        try:
            imad = eng.bench_imad(False)
            imadw = eng.bench_imad(True)
            line["int32_peak"] = {"imad_gops": imad, "imad_wide_gops": imadw}
            if wl == "mulmod":
                line["roofline_int"] = {"bound": "int32", "achieved": 162 * value / 1e9, "peak": imadw, "unit": "G IMAD.WIDE/s",
                                        "frac": 162 * value / 1e9 / imadw}
        except Exception as e:          # noqa: BLE001
            line["int32_peak"] = {"error": str(e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle import cpu_baseline
                cb = cpu_baseline.run(wl, steps=1, warmup=0, budget_s=12.0)
                line["cpu_baseline"] = {"value": cb["value"], "unit": cb["unit"], "cores": cb["cores"], "kind": cb["kind"],
                                        "sample": cb["sample"]}
            except Exception as e:      # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "error": str(e)}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
