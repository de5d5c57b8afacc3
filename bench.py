#!/usr/bin/env python
"""bench.py -- `python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]`

One "step" = one pass of the hot path over one batch of synthetic input.  Workloads (BASELINE.json):
  ecdsa  (default, configs[2]) 4096 verify_secp256k1_message_circuit witnesses per GPU, full tape
  mulmod (configs[1])          2^20 nonnative secp256k1 base-field mulmod witnesses
  eval   (configs[4])          MulNonnativeGate constraint evaluation over 2^20 Goldilocks LDE rows
  glv    (configs[3])          GLV decomposition + curve MSM witness chain (glv_mul) for 2^16 scalars
The headline line is the ecdsa workload ("ECDSA circuit witnesses/s"); the mulmod and eval workloads
are measured in the same run and reported under "secondary" with their own rooflines
("nonnative mulmod/s" is the second half of BASELINE.json's metric).
`value` = whole-job throughput with inputs resident in HBM; `e2e` = the same metric through the
host-pointer C-ABI call (pinned host buffers; H2D + kernels + D2H inside the timed region).
N > 1: one process per GPU (torchrun).  The ecdsa workload defaults to STRONG scaling of configs[2]: the 4096 signatures of a
step are sharded across the ranks (sharding.shard_range), no data-path collective; the final result gather (NCCL) and a
plain pinned D2H probe are measured beside it, and the weak line (4096 per rank) is reported under "weak_scaling".
`--scaling weak` makes the weak line the headline.  Timings are device times, max over ranks.
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
# Independent batches are pipelined on 32 CUDA streams: with the default 8 hardware work queues they alias and serialise
# (measured with 16 streams: N = 1: 0.79 -> 0.67 ms per step with 32 queues; N = 2, 2048-signature shards: 0.523 -> 0.386 ms).
# Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOAD_NAMES = {
    "ecdsa": "4096 independent verify_secp256k1_message_circuit witnesses per GPU (full tape: 82615 slots each)",
    "mulmod": "2^20 secp256k1 base-field nonnative mulmod witnesses (29-bit limbs): q, r, check_sum, carries",
    "mulmod_n": "2^20 secp256k1 SCALAR-field nonnative mulmod witnesses (29-bit limbs): q, r, check_sum, carries",
    "eval": "MulNonnativeGate eval_unfiltered_base_batch over 2^20 Goldilocks LDE rows",
    "glv": "glv_mul witness chain (GLV decomposition + 2-scalar MSM) for 2^16 secp256k1 scalars, P = 7G",
    "evalcs": "CheckSumGate eval_unfiltered_base_batch over 2^20 Goldilocks LDE rows",
    "evalrc": "u29 range-check gate (9 limbs, 3-bit aux limbs; plonky2_ux spec unpinned) eval over 2^20 Goldilocks LDE rows",
    "evalcmp": "u29 comparison gate (2-bit chunks; plonky2_ux spec unpinned) eval over 2^20 Goldilocks LDE rows",
    "evaladd": "u29 add-many gate (3 addends, 5 ops / row; plonky2_ux spec unpinned) eval over 2^20 Goldilocks LDE rows",
    "evalfc": "MulNonnativeGate eval_filtered_base_batch + alpha-combination (2 challenges, fused) over 2^20 Goldilocks LDE rows",
    "evalcsfc": "CheckSumGate eval_filtered_base_batch + alpha-combination (2 challenges, fused) over 2^20 Goldilocks LDE rows",
    "evalrcfc": "u29 range-check gate (9 limbs, 3-bit aux limbs; plonky2_ux spec unpinned) eval_filtered + alpha-combination (2 challenges, fused) over 2^20 Goldilocks LDE rows",
    "genstream": "serialized generator stream (8 x curve_add + curve_double chained: 224 in-tree generators, target-level wiring of "
                 "gadgets/curve.rs) run by the level-scheduled executor for 4096 circuits",
    "rcrows": "every plonky2_ux range-check gate row of the ECDSA circuit (9168 range_check_ux_circuit calls per signature: add_many, sub, "
              "mul_nonnative; 512 signatures, 3-bit aux limbs: row layout unpinned) from the device-resident compact witness",
    "genstream_ecdsa": "the complete verify_secp256k1_message_circuit as a serialized generator stream (30955 generators incl. the plonky2 glue, "
                       "1862 dependency levels) run by the level-scheduled executor for 4096 circuits",
    "lde": "coset LDE (ifft 2^17 -> zero-pad -> coset fft 2^20, rate 8) of 135 wire polynomials: the matrix the gate evaluators read",
}
# plonky2_ux eval workloads: (gate, bits, limb_bits, count, num_addends, kernel)
UX_WORKLOADS = {
    "evalrc": (0, 29, 3, 9, 0, "k_eval_ux_range_check<3>"),
    "evalcmp": (1, 29, 2, 1, 0, "k_eval_ux_comparison<2>"),
    "evaladd": (2, 29, 2, 5, 3, "k_eval_ux_add_many<2>"),
}
METRICS = {
    "ecdsa": ("ecdsa_circuit_witnesses_per_s", "witnesses/s"),
    "mulmod": ("nonnative_mulmod_witnesses_per_s", "mulmod/s"),
    "mulmod_n": ("nonnative_scalar_field_mulmod_witnesses_per_s", "mulmod/s"),
    "eval": ("mul_nonnative_gate_rows_per_s", "rows/s"),
    "glv": ("glv_mul_witness_chains_per_s", "chains/s"),
    "evalcs": ("check_sum_gate_rows_per_s", "rows/s"),
    "evalrc": ("ux_range_check_gate_rows_per_s", "rows/s"),
    "evalcmp": ("ux_comparison_gate_rows_per_s", "rows/s"),
    "evaladd": ("ux_add_many_gate_rows_per_s", "rows/s"),
    "evalfc": ("mul_nonnative_gate_filtered_combined_rows_per_s", "rows/s"),
    "evalcsfc": ("check_sum_gate_filtered_combined_rows_per_s", "rows/s"),
    "evalrcfc": ("ux_range_check_gate_filtered_combined_rows_per_s", "rows/s"),
    "genstream": ("generator_stream_circuits_per_s", "circuits/s"),
    "genstream_ecdsa": ("generator_stream_ecdsa_circuits_per_s", "circuits/s"),
    "lde": ("coset_lde_output_elements_per_s", "elements/s"),
    "rcrows": ("ecdsa_mul_range_check_gate_rows_per_s", "rows/s"),
}


# config 5 is quoted on a 2^20-row LDE; BASELINE.json's wording also admits trace 2^20 x rate 8 = 2^23 rows:
# P2E_BENCH_LOG_ROWS=23 re-runs the gate-evaluation workloads at that size
LOG_ROWS = int(os.environ.get("P2E_BENCH_LOG_ROWS", "20"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs.  Uses NVML in-process (pynvml):
    forking `nvidia-smi` from a process that holds tens of GB of pinned memory stalls the launching
    thread for milliseconds (page-table copy) and showed up as 2x noise in ms_per_step.  Falls back to
    ONE long-lived `nvidia-smi --loop-ms` child started before the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.nvml = None
        self.handle = None
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:       # noqa: BLE001
            self.nvml = None
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(index),
                                              "--loop-ms=50"], stdout=subprocess.PIPE, text=True)
            except Exception:   # noqa: BLE001
                self.proc = None

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml is not None:
            n = self.nvml
            bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                try:
                    self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for nm, b in bits.items():
                        if r & b:
                            self.reasons.add(nm)
                except Exception:       # noqa: BLE001
                    pass
                time.sleep(0.02)
            return
        if self.proc is None:
            return
        for line in self.proc.stdout:
            if self.stop_flag:
                break
            try:
                f = [t.strip() for t in line.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:           # noqa: BLE001
                pass
        try:
            self.proc.terminate()
        except Exception:               # noqa: BLE001
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:           # noqa: BLE001
                pass

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "how": "nvml" if self.nvml is not None else "nvidia-smi --loop-ms"}


def gs_walked(eng, gs):
    """Whether p2e_genstream_run* takes the level walker for this stream (option genstream_walk; the bench streams hold no
    BigUint div_rem group, the one kind the walker lacks)."""
    w = eng.get_option("genstream_walk")
    return w == 2 or (w == 1 and gs.num_levels >= 32)


def gs_label(eng, gs):
    how = ("k_gs_walk, ONE launch: a CTA owns 16-32 circuits and walks every level (warp = generator, __syncthreads per level)"
           if gs_walked(eng, gs) else "one fused kernel per launch group, one CUDA graph")
    return "level schedule: %d generators, %d levels, %d launch groups; %s" % (gs.num_generators, gs.num_levels, gs.num_groups, how)


def gs_launches(eng, gs):
    return 2 if gs_walked(eng, gs) else gs.num_groups + 1         # + k_gs_set_inputs


def workload_label(wl, strong, per_gpu):
    name = WORKLOAD_NAMES[wl].replace("2^20 Goldilocks", "2^%d Goldilocks" % LOG_ROWS)
    if strong:
        name = name.replace("per GPU", "per step, sharded %d per GPU (rank g owns [g N/G, (g+1) N/G))" % per_gpu)
    return name


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """Times the reference's CPU algorithms for the path on the box's host cores.  The Rust crate
    cannot be built in this image (no cargo/rustc, un-vendored git deps), so this is the oracle's C
    port of the reference's generators (oracle/p2e_oracle.c): kind == "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    res = cpu_baseline.run(args.workload, steps=args.steps, warmup=min(args.warmup, 1),
                           budget_s=float(os.environ.get("P2E_REF_BUDGET_S", "60")))
    # same workload name / scaling label as our arm at this N (the CPU arm always runs on rank 0's host cores only)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = args.scaling == "strong" and world > 1 and args.workload == "ecdsa"
    cfg = dict(res["config"])
    cfg["workload"] = workload_label(args.workload, strong, 4096 // max(world, 1))
    line = {
        "impl": "reference", "metric": res["metric"], "value": res["value"], "unit": res["unit"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": res["value"], "unit": res["unit"], "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": res["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm (GPU)
class Bench:
    def __init__(self, args):
        import numpy as np
        import torch
        import plonky2_ecdsa_b200 as pkg
        from plonky2_ecdsa_b200 import workload as W
        self.np, self.torch, self.pkg, self.W = np, torch, pkg, W
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.eng = pkg.Engine(self.local_rank)
        self.lib = self.eng.lib
        for opt in ("chain_split", "chain_replay_lanes", "genstream_walk", "genstream_walk_shape"):   # experiment hooks of the bench (the library reads no environment): P2E_OPT_CHAIN_SPLIT=0, ...
            if os.environ.get("P2E_OPT_" + opt.upper()) is not None:
                self.eng.set_option(opt, int(os.environ["P2E_OPT_" + opt.upper()]))
        self.st = torch.cuda.Stream()          # a real (non-NULL) stream: NULL selects the library's own
        torch.cuda.set_stream(self.st)
        self.sp = self.st.cuda_stream
        assert self.sp != 0
        self.peak_gbs, self.peak_kind = load_peaks()

    def dev(self, a):
        return self.torch.from_numpy(a.view(self.np.int64)).cuda()

    def pinned(self, a):
        return self.torch.from_numpy(a.view(self.np.int64)).pin_memory()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def d2h_probe(self, mb=512, reps=4):
        """Plain pinned D2H ceiling of this box with all ranks copying at once (names the limiter of the e2e numbers at N > 1):
        every rank copies `mb` MiB device -> pinned host `reps` times between barriers; returns per-rank and aggregate GB/s."""
        torch = self.torch
        n = mb << 20
        src = torch.empty(n, dtype=torch.uint8, device="cuda")
        dst = torch.empty(n, dtype=torch.uint8).pin_memory()
        dst.copy_(src)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        self.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        gbs = n * reps / float(t[0]) / 1e9
        return {"ranks": self.world, "per_rank_GBps": gbs, "aggregate_GBps": gbs * self.world, "bytes_per_rank": n * reps,
                "how": "all ranks copy device -> pinned host concurrently; time = max over ranks"}

    def setup(self, wl):
        """Returns dict(step, step_e2e, units, h2d, d2h, alg_bytes, kernel, launches, errs)."""
        np, torch, pkg, W, lib, eng, sp = self.np, self.torch, self.pkg, self.W, self.lib, self.eng, self.sp
        if wl in ("mulmod", "mulmod_n"):
            n = 1 << 20
            mod = 0 if wl == "mulmod" else 1
            x, y = W.mulmod_workload(n, mod, seed=W.SEED + self.rank)
            xd, yd = self.dev(x), self.dev(y)
            od = torch.empty((pkg.MUL_OUT_ROWS, n), dtype=torch.int64, device="cuda")
            ed = torch.empty(n, dtype=torch.int32, device="cuda")

            def step():
                rc = lib.p2e_mul_nonnative_witness_dev(mod, n, xd.data_ptr(), yd.data_ptr(), n, od.data_ptr(), n, ed.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hx, hy = self.pinned(x), self.pinned(y)
            ho = torch.empty((pkg.MUL_OUT_ROWS, n), dtype=torch.int64).pin_memory()
            he = torch.empty(n, dtype=torch.int32).pin_memory()

            def step_e2e():
                rc = lib.p2e_mul_nonnative_witness(mod, n, hx.data_ptr(), hy.data_ptr(), ho.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=2 * 9 * 8 * n, d2h=pkg.MUL_OUT_ROWS * 8 * n + 4 * n,
                        alg_bytes=552 * n, kernel="k_mul_witness<ModP>" if mod == 0 else "k_mul_witness<ModN>", launches=1, err=ed, dtype="u32 limbs / u64 (int)",
                        keep=(xd, yd, od, hx, hy, ho, he))
        if wl == "eval":
            n = 1 << LOG_ROWS
            w = W.random_goldilocks(W.SEED + self.rank, 53, n)
            wd = self.dev(w)
            od = torch.empty((17, n), dtype=torch.int64, device="cuda")

            def step():
                rc = lib.p2e_eval_mul_nonnative_dev(0, n, wd.data_ptr(), n, od.data_ptr(), n, sp)
                assert rc == 0, eng.last_error()
            hw = self.pinned(w)
            ho = torch.empty((17, n), dtype=torch.int64).pin_memory()

            def step_e2e():
                rc = lib.p2e_eval_mul_nonnative(0, n, hw.data_ptr(), ho.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=53 * 8 * n, d2h=17 * 8 * n, alg_bytes=560 * n,
                        kernel="k_eval_mul_nonnative<ModP>", launches=1, err=None, dtype="u64 (Goldilocks)",
                        keep=(wd, od, hw, ho))
        if wl == "evalcs":
            n = 1 << LOG_ROWS
            w = W.random_goldilocks(W.SEED + 3 + self.rank, 33, n)
            wd = self.dev(w)
            od = torch.empty((17, n), dtype=torch.int64, device="cuda")

            def step():
                rc = lib.p2e_eval_check_sum_dev(n, wd.data_ptr(), n, od.data_ptr(), n, sp)
                assert rc == 0, eng.last_error()
            hw = self.pinned(w)
            ho = torch.empty((17, n), dtype=torch.int64).pin_memory()

            def step_e2e():
                rc = lib.p2e_eval_check_sum(n, hw.data_ptr(), ho.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=33 * 8 * n, d2h=17 * 8 * n, alg_bytes=400 * n,
                        kernel="k_eval_check_sum", launches=1, err=None, dtype="u64 (Goldilocks)", keep=(wd, od, hw, ho))
        if wl in ("evalfc", "evalcsfc", "evalrcfc"):
            import ctypes
            gate, nw = {"evalfc": (0, 53), "evalcsfc": (1, 33), "evalrcfc": (2, 99)}[wl]
            n = 1 << LOG_ROWS
            w = W.random_goldilocks(W.SEED + 5 + gate + self.rank, nw + 1, n)      # the wires + the selector constant column
            w[nw, ::3] = 0                                                          # a third of the rows select this gate exactly
            wd = self.dev(w)
            od = torch.empty((2, n), dtype=torch.int64, device="cuda")
            prm = (ctypes.c_int * 5)(0, 29, 3, 9, 0) if gate == 2 else (ctypes.c_int * 5)(0, 0, 0, 0, 0)
            alphas = np.array([0x0123456789ABCDEF, 0xFEDCBA9876543210 % ((1 << 64) - (1 << 32) + 1)], dtype=np.uint64)
            sel_ptr = wd.data_ptr() + nw * n * 8
            scr = torch.empty(int(lib.p2e_eval_filtered_scratch_bytes(gate, ctypes.addressof(prm), n)) // 8 + 2, dtype=torch.int64, device="cuda")

            def step():
                rc = lib.p2e_eval_filtered_combined_dev(gate, ctypes.addressof(prm), 0, 0, 4, 1, n, wd.data_ptr(), n, sel_ptr,
                                                        alphas.ctypes.data, 2, 4, 0, od.data_ptr(), n, scr.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hw = self.pinned(w)
            ho = torch.empty((2, n), dtype=torch.int64).pin_memory()

            def step_e2e():
                rc = lib.p2e_eval_filtered_combined(gate, ctypes.addressof(prm), 0, 0, 4, 1, n, hw.data_ptr(), hw.data_ptr() + nw * n * 8,
                                                    alphas.ctypes.data, 2, 4, 0, ho.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=(nw + 1) * 8 * n, d2h=2 * 8 * n, alg_bytes=(nw + 1 + 2) * 8 * n,
                        kernel=["k_eval_mul_nonnative_fc<2>", "k_eval_check_sum_fc<2>", "k_eval_ux_range_check_fc<3, 2>"][gate],
                        launches=1 if gate < 2 else 2, err=None, dtype="u64 (Goldilocks)", keep=(wd, od, hw, ho, prm, alphas, scr))
        if wl == "genstream":
            b = pkg.TargetCircuitBuilder()
            p0 = (b.add_virtual_nonnative_target(), b.add_virtual_nonnative_target())
            q0 = (b.add_virtual_nonnative_target(), b.add_virtual_nonnative_target())
            a_const = b.add_virtual_nonnative_target()
            acc = p0
            for _ in range(8):
                acc = b.curve_double(b.curve_add(acc, q0), a_const)
            gs = eng.load_generator_stream(b.bytes())
            n = 4096
            d = W.ecdsa_workload(n, seed=W.SEED + self.rank)
            rolled = [np.ascontiguousarray(np.roll(d[k], 1, axis=1)) for k in ("pkx", "pky")]
            in_t = list(p0[0]) + list(p0[1]) + list(q0[0]) + list(q0[1]) + list(a_const)
            vals = np.concatenate([d["pkx"], d["pky"], rolled[0], rolled[1], np.zeros((9, n), dtype=np.uint64)], axis=0)
            in_slots = np.array([gs.slot(t) for t in in_t], dtype=np.uint32)
            out_t = list(acc[0]) + list(acc[1])
            out_slots = np.array([gs.slot(t) for t in out_t], dtype=np.uint32)
            n_slots = gs.num_slots
            vd = self.dev(np.ascontiguousarray(vals))
            table = torch.zeros((n_slots, n), dtype=torch.int64, device="cuda")
            ed = torch.zeros(n, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()

            def step():
                rc = lib.p2e_genstream_run_dev(gs.handle, n, in_slots.ctypes.data, len(in_slots), vd.data_ptr(), n, table.data_ptr(), n,
                                               ed.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hv = self.pinned(np.ascontiguousarray(vals))
            ho = torch.empty((len(out_slots), n), dtype=torch.int64).pin_memory()
            he = torch.empty(n, dtype=torch.int32).pin_memory()

            def step_e2e():
                rc = lib.p2e_genstream_run(gs.handle, n, in_slots.ctypes.data, len(in_slots), hv.data_ptr(), out_slots.ctypes.data,
                                           len(out_slots), ho.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()
            # algorithmic bytes: every slot of the value table is written once and read about once
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=len(in_slots) * 8 * n, d2h=len(out_slots) * 8 * n + 4 * n,
                        alg_bytes=2 * n_slots * 8 * n, kernel=gs_label(eng, gs), launches=gs_launches(eng, gs), err=ed, dtype="u32 limbs / u64 (int)",
                        keep=(gs, vd, table, hv, ho, he, in_slots, out_slots))
        if wl == "genstream_ecdsa":
            from plonky2_ecdsa_b200 import engine as E, genstream as G
            cb = G.EcdsaCircuitBuilder(E.blinding_scalar(lib))
            msg_t, r_t, s_t, pk_t = cb.verify_secp256k1_message()
            gs = eng.load_generator_stream(cb.bytes())
            n = 4096
            d = W.ecdsa_workload(n, seed=W.SEED + self.rank)
            in_t = list(msg_t) + list(r_t) + list(s_t) + list(pk_t[0]) + list(pk_t[1])
            vals = np.ascontiguousarray(np.concatenate([d["msg"], d["r"], d["s"], d["pkx"], d["pky"]], axis=0))
            in_slots = np.array([gs.slot(t) for t in in_t], dtype=np.uint32)
            out_t = cb.tape_targets[-10:-1]                 # y3 of the final curve_add stands for "the result" on the host path
            out_slots = np.array([gs.slot(t) for t in out_t], dtype=np.uint32)
            n_slots = gs.num_slots
            vd = self.dev(vals)
            table = torch.zeros((n_slots, n), dtype=torch.int64, device="cuda")
            ed = torch.zeros(n, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()

            def step():
                rc = lib.p2e_genstream_run_dev(gs.handle, n, in_slots.ctypes.data, len(in_slots), vd.data_ptr(), n, table.data_ptr(), n,
                                               ed.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hv = self.pinned(vals)
            ho = torch.empty((len(out_slots), n), dtype=torch.int64).pin_memory()
            he = torch.empty(n, dtype=torch.int32).pin_memory()

            def step_e2e():
                rc = lib.p2e_genstream_run(gs.handle, n, in_slots.ctypes.data, len(in_slots), hv.data_ptr(), out_slots.ctypes.data,
                                           len(out_slots), ho.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=len(in_slots) * 8 * n, d2h=len(out_slots) * 8 * n + 4 * n,
                        alg_bytes=2 * n_slots * 8 * n, kernel=gs_label(eng, gs), launches=gs_launches(eng, gs), err=ed, dtype="u32 limbs / u64 (int)",
                        keep=(gs, vd, table, hv, ho, he, in_slots, out_slots))
        if wl == "rcrows":
            import ctypes
            n = 512
            d = W.ecdsa_workload(n, seed=W.SEED + self.rank)
            ins = [d[k] for k in ("msg", "r", "s", "pkx", "pky")]
            ind = [self.dev(a) for a in ins]
            arr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in ind])
            inp = ctypes.cast(arr, ctypes.c_void_p)
            _, nn, _ = eng.tape_compact_layout(eng.TAPE_ECDSA)
            gen_c, cnt_c, bits_c, off_c = eng.tape_range_checks(eng.TAPE_ECDSA, 3)
            nc, wires = len(gen_c), int(off_c[-1])
            nar = torch.empty((nn, n), dtype=torch.int32, device="cuda")
            ed = torch.empty(n, dtype=torch.int32, device="cuda")
            ed2 = torch.empty(n, dtype=torch.int32, device="cuda")
            rows = torch.empty((wires, n), dtype=torch.int64, device="cuda")
            assert lib.p2e_tape_witness_compact_dev(eng.TAPE_ECDSA, n, inp, n, nar.data_ptr(), None, n, ed.data_ptr(), sp) == 0, eng.last_error()
            torch.cuda.synchronize()
            assert not bool(ed.any())

            def step():
                rc = lib.p2e_tape_range_check_rows_dev(eng.TAPE_ECDSA, n, inp, n, None, nar.data_ptr(), n, 0, nc, 3, rows.data_ptr(), n,
                                                       ed2.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hin = [self.pinned(a) for a in ins]
            he = torch.empty(n, dtype=torch.int32).pin_memory()

            def step_e2e():
                # host inputs -> witness (compact, device) -> rows (device) -> the error vector back: the rows stay in HBM for
                # their consumer (3.1 GB per 512 signatures -- shipping them would only measure the PCIe link)
                for t, h in zip(ind, hin):
                    t.copy_(h, non_blocking=True)
                assert lib.p2e_tape_witness_compact_dev(eng.TAPE_ECDSA, n, inp, n, nar.data_ptr(), None, n, ed.data_ptr(), sp) == 0
                step()
                he.copy_(ed2, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            # every wire written once (u64), every input read once (u32 in the compact witness)
            return dict(step=step, step_e2e=step_e2e, units=nc * n, h2d=5 * 9 * 8 * n, d2h=4 * n,
                        alg_bytes=(wires * 8 + int(cnt_c.sum()) * 4) * n, kernel="k_tape_range_check_rows", launches=1, err=ed2,
                        dtype="u32 / u64 (int)", keep=(ind, nar, ed, rows, hin, he, arr))
        if wl in UX_WORKLOADS:
            gate, bits, lb, count, na, kname = UX_WORKLOADS[wl]
            nw, nc, _ = eng.ux_gate_shape(gate, bits, lb, count, na)
            n = 1 << LOG_ROWS
            w = W.random_goldilocks(W.SEED + 11 + gate + self.rank, nw, n)
            wd = self.dev(w)
            od = torch.empty((nc, n), dtype=torch.int64, device="cuda")

            def step():
                rc = lib.p2e_eval_ux_dev(gate, bits, lb, count, na, n, wd.data_ptr(), n, od.data_ptr(), n, sp)
                assert rc == 0, eng.last_error()
            hw = self.pinned(w)
            ho = torch.empty((nc, n), dtype=torch.int64).pin_memory()

            def step_e2e():
                rc = lib.p2e_eval_ux(gate, bits, lb, count, na, n, hw.data_ptr(), ho.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, h2d=nw * 8 * n, d2h=nc * 8 * n, alg_bytes=(nw + nc) * 8 * n,
                        kernel=kname, launches=1, err=None, dtype="u64 (Goldilocks)", keep=(wd, od, hw, ho))
        if wl == "lde":
            log_n, rb, polys = 17, 3, 135
            n, m = 1 << log_n, 1 << (log_n + rb)
            v = W.random_goldilocks(W.SEED + 21 + self.rank, polys, n)
            vd = self.dev(v)
            od = torch.empty((polys, m), dtype=torch.int64, device="cuda")
            scr = torch.empty(int(lib.p2e_ntt_scratch_bytes(log_n, rb, polys)) // 8, dtype=torch.int64, device="cuda")

            def step():
                rc = lib.p2e_coset_lde_dev(log_n, rb, polys, vd.data_ptr(), n, od.data_ptr(), m, scr.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            hv = self.pinned(v)
            ho = torch.empty((polys, m), dtype=torch.int64).pin_memory()

            def step_e2e():
                rc = lib.p2e_coset_lde(log_n, rb, polys, hv.data_ptr(), ho.data_ptr())
                assert rc == 0, eng.last_error()
            # algorithmic bytes: read the n values, write the 8 n LDE values (the two intermediate passes are the
            # implementation's, not the algorithm's)
            return dict(step=step, step_e2e=step_e2e, units=polys * m, h2d=polys * n * 8, d2h=polys * m * 8,
                        alg_bytes=polys * (n + m) * 8, kernel="k_ntt_cols + k_ntt_rows (ifft 2^17, fft 2^20)", launches=4, err=None,
                        dtype="u64 (Goldilocks)", keep=(vd, od, scr, hv, ho))
        if wl == "glv":
            # device-resident only at full size (the 2^16 x 61.6k-slot witness is 32 GB); e2e on a 4096-scalar slice
            n = 1 << 16
            slots = int(lib.p2e_glv_mul_tape_slots())
            k = W.random_field_limbs(W.SEED + 7 + self.rank, n, 1)
            gx, gy = W._ec_mul_g(7)
            px = np.repeat(W.ints_to_limbs([gx]), n, axis=1)
            py = np.repeat(W.ints_to_limbs([gy]), n, axis=1)
            ind = [self.dev(np.ascontiguousarray(a)) for a in (k, px, py)]
            od = torch.empty((slots, n), dtype=torch.int64, device="cuda")
            ed = torch.empty(n, dtype=torch.int32, device="cuda")
            ptrs = [t.data_ptr() for t in ind]

            def step():
                rc = lib.p2e_glv_mul_witness_dev(n, *ptrs, n, od.data_ptr(), n, ed.data_ptr(), sp)
                assert rc == 0, eng.last_error()
            m = 4096
            hin = [self.pinned(np.ascontiguousarray(a[:, :m])) for a in (k, px, py)]
            ho = torch.empty((slots, m), dtype=torch.int64).pin_memory()
            he = torch.empty(m, dtype=torch.int32).pin_memory()

            def step_e2e():
                rc = lib.p2e_glv_mul_witness(m, *[t.data_ptr() for t in hin], ho.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()
            return dict(step=step, step_e2e=step_e2e, units=n, e2e_units=m, h2d=3 * 9 * 8 * m, d2h=slots * 8 * m + 4 * m,
                        alg_bytes=slots * 8 * n, kernel="k_chain_expand", launches=9, err=ed,
                        dtype="u32 limbs / u64 (int)", keep=(ind, od, hin, ho, he), slots=slots)
        # ---- ecdsa (configs[2]): 4096 signatures per step.  N > 1: strong scaling by default -- the 4096 signatures of a
        # step are sharded [g N/G, (g+1) N/G) across the ranks (plonky2_ecdsa_b200.sharding), every rank runs its shard.
        import ctypes
        from plonky2_ecdsa_b200 import sharding
        n_total = 4096
        strong = self.args.scaling == "strong" and self.world > 1
        if strong:
            lo, hi = sharding.shard_range(n_total, self.rank, self.world)
            d_all = W.ecdsa_workload(n_total, seed=W.SEED)               # every rank derives the same global batch
            d = {k: np.ascontiguousarray(d_all[k][:, lo:hi]) for k in ("msg", "r", "s", "pkx", "pky")}
            n = hi - lo
        else:
            d = W.ecdsa_workload(n_total, seed=W.SEED + self.rank)
            n = n_total
        slots = int(lib.p2e_ecdsa_tape_slots())
        _, nn, nw = eng.tape_compact_layout(eng.TAPE_ECDSA)
        ins = [d[k] for k in ("msg", "r", "s", "pkx", "pky")]
        ind = [self.dev(a) for a in ins]
        ptrs = [t.data_ptr() for t in ind]

        # Independent batches are enqueued round-robin on S streams (async *_dev API, one output buffer + one scratch
        # arena per stream) so the latency-bound per-signature kernels of one batch overlap the full-GPU kernels of
        # another.  32 streams = the 32 hardware queues (CUDA_DEVICE_MAX_CONNECTIONS): more alias, fewer leave the device
        # short of batches in flight (4096 per call: 16 / 24 / 32 streams 0.620 / 0.606 / 0.600 ms; 512 per call: 16 / 32 / 64 /
        # 128 streams 0.142 / 0.118 / 0.119 / 0.119 ms; scripts/small_batch.py)
        S = max(1, int(os.environ.get("P2E_BENCH_STREAMS", "32")))
        streams = [torch.cuda.Stream() for _ in range(S)]
        nars = [torch.empty((nn, n), dtype=torch.int32, device="cuda") for _ in range(S)]
        eds = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(S)]
        ed = eds[0]
        state = {"k": 0}

        # headline device-resident form: the derived compact witness (u32 narrow slots, CheckSum carries as low words +
        # packed tops, check_sum columns not stored: rebuilt by the consumer from the carries, include/p2e.h)
        def step():
            j = state["k"] % S
            state["k"] += 1
            rc = lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[j].data_ptr(), None, n, eds[j].data_ptr(), streams[j].cuda_stream)
            assert rc == 0, eng.last_error()

        def step_single():
            rc = lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[0].data_ptr(), None, n, ed.data_ptr(), sp)
            assert rc == 0, eng.last_error()

        # the two other device formats, measured the same way (secondary): compact with the check_sum columns, plain u64
        S2 = min(S, 12)
        wides = [torch.empty((nw, n), dtype=torch.int64, device="cuda") for _ in range(S2)]

        def step_compact_wide():
            j = state["k"] % S2
            state["k"] += 1
            rc = lib.p2e_ecdsa_witness_compact_dev(n, *ptrs, n, nars[j].data_ptr(), wides[j].data_ptr(), n, eds[j].data_ptr(), streams[j].cuda_stream)
            assert rc == 0, eng.last_error()
        ods = [torch.empty((slots, n), dtype=torch.int64, device="cuda") for _ in range(S2)]

        def step_plain():
            j = state["k"] % S2
            state["k"] += 1
            rc = lib.p2e_ecdsa_witness_dev(n, *ptrs, n, ods[j].data_ptr(), n, eds[j].data_ptr(), streams[j].cuda_stream)
            assert rc == 0, eng.last_error()

        hin = [self.pinned(a) for a in ins]
        he = torch.empty(n, dtype=torch.int32).pin_memory()
        hn = torch.empty((nn, n), dtype=torch.int32).pin_memory()
        in_arr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in hin])
        in_ptr = ctypes.cast(in_arr, ctypes.c_void_p)

        # e2e headline: host buffers in, derived compact witness out, through the host-pointer C-ABI call (every rank ships
        # its shard to host memory over its own PCIe link: in the single-process host these are regions of one buffer)
        def step_e2e():
            rc = lib.p2e_tape_witness_compact(eng.TAPE_ECDSA, n, in_ptr, hn.data_ptr(), None, he.data_ptr())
            assert rc == 0, eng.last_error()
        extra = {}
        if not strong:
            ho = torch.empty((slots, n), dtype=torch.int64).pin_memory()
            hw = torch.empty((nw, n), dtype=torch.int64).pin_memory()

            def step_e2e_plain():
                rc = lib.p2e_ecdsa_witness(n, *[t.data_ptr() for t in hin], ho.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()

            def step_e2e_compact():
                rc = lib.p2e_tape_witness_compact(eng.TAPE_ECDSA, n, in_ptr, hn.data_ptr(), hw.data_ptr(), he.data_ptr())
                assert rc == 0, eng.last_error()
            extra = dict(step_e2e_plain=step_e2e_plain, step_e2e_compact=step_e2e_compact, e2e_keep2=(ho, hw))
        gather = None
        if strong:
            # the final result gather of north_star: the compact shards to rank 0's HBM over NCCL / NVLink
            rg = sharding.RootGather(n_total, nn, torch.int32, torch.device("cuda"), self.dist, dst=0)

            def gather():
                rg(nars[0])
            extra["gather_keep"] = rg
            extra["gather_bytes"] = rg.bytes_into_root
        return dict(step=step, step_e2e=step_e2e, units=n, units_total=n_total if strong else n * self.world,
                    h2d=5 * 9 * 8 * n, d2h=nn * 4 * n + 4 * n, d2h_compact=nn * 4 * n + nw * 8 * n + 4 * n,
                    d2h_plain=slots * 8 * n + 4 * n, e2e_keep=(hn, in_arr),
                    alg_bytes=(nn * 4) * n, alg_bytes_plain=slots * 8 * n, kernel="k_chain_expand", launches=9, err=ed,
                    dtype="u32 words / 29-bit limbs (int)", keep=(ind, nars, eds, wides, ods, hin, he), slots=slots, streams=streams,
                    step_single=step_single, step_compact_wide=step_compact_wide, step_plain=step_plain, gather=gather,
                    narrow_rows=nn, wide_rows=nw, strong=strong, **extra)

    def measure(self, wl, steps, warmup, e2e_steps, sample_clocks=False):
        torch = self.torch
        s = self.setup(wl)
        step, step_e2e = s["step"], s["step_e2e"]
        sampler = None
        if sample_clocks and self.rank == 0:
            # rank 0 only; created before the warm-up so that nothing forks inside the timed region
            sampler = ClockSampler(self.local_rank)
        # every stream's scratch arena is allocated on its first batch and the library captures a repeated call as a CUDA graph
        # on its second occurrence: warm up at least two batches per stream (plus one replay)
        for _ in range(max(warmup, 3 * len(s.get("streams") or []))):
            step()
        self.barrier()
        if sampler is not None:
            sampler.start()
        streams = s.get("streams") or []

        stagger_cycles = 0

        def timed(fn, k):
            """k calls of fn between two events on the main stream; worker streams fork from / join into it."""
            nonlocal stagger_cycles
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.st)
            for j, w in enumerate(streams):
                w.wait_event(e0)
                if stagger_cycles and j:
                    # identical batches would otherwise run in lock-step (all latency-bound phases together,
                    # then all bandwidth-bound phases together): offset stream j's start, inside the timed region
                    with torch.cuda.stream(w):
                        torch.cuda._sleep(int(j * stagger_cycles))
            for _ in range(k):
                fn()
            for w in streams:
                ev = torch.cuda.Event()
                ev.record(w)
                self.st.wait_event(ev)
            e1.record(self.st)
            self.barrier()
            return e0.elapsed_time(e1)
        stagger_ms = float(os.environ.get("P2E_BENCH_STAGGER_MS", "0")) if len(streams) > 1 else 0.0
        stagger_cycles = stagger_ms * 1e-3 * 1.9e9
        ms = timed(step, steps)
        stagger_cycles = 0
        latency_ms = None
        latency2_ms = None
        if "step_single" in s:
            s["step_single"]()
            self.barrier()
            latency_ms = timed(s["step_single"], min(steps, 10)) / min(steps, 10)
            if wl == "ecdsa":
                # the same single batch with the two-lane (latency) form of the replay kernel, whatever the call size
                old_lanes = self.eng.get_option("chain_replay_lanes")
                self.eng.set_option("chain_replay_lanes", 2)
                s["step_single"]()
                self.barrier()
                latency2_ms = timed(s["step_single"], min(steps, 10)) / min(steps, 10)
                self.eng.set_option("chain_replay_lanes", old_lanes)
        if sampler is not None:
            t_end = time.time() + 0.7          # keep the load on until nvidia-smi has a few samples
            while time.time() < t_end:
                step()
            torch.cuda.synchronize()
            sampler.stop()
            sampler.join(timeout=2.0)
        kernel_ms = None
        formats = {}
        gather_res = None
        if wl == "ecdsa":
            self.eng.set_profiling(True)
            s["step_single"]()
            kernel_ms = self.eng.last_kernel_times()
            self.eng.set_profiling(False)
            # the other two device-resident formats, same pipelining
            for nm, key, by in (("compact_with_check_sum", "step_compact_wide", (s["narrow_rows"] * 4 + s["wide_rows"] * 8)),
                                ("plain_u64", "step_plain", s["slots"] * 8)):
                fn = s[key]
                for _ in range(36):                # 3 per stream (S2 <= 12): arena, graph capture, one replay
                    fn()
                self.barrier()
                k2 = max(8, min(steps, 24))
                t2 = timed(fn, k2) / k2
                formats[nm] = {"ms_per_step": t2, "bytes_per_witness": by}
            if s["gather"] is not None:
                # strong scaling: the compact shards gathered into rank 0's HBM over NCCL (one gather per step, on the
                # main stream after the shard's kernels)
                def step_gather():
                    s["step_single"]()
                    s["gather"]()
                step_gather()
                self.barrier()
                kg = max(4, min(steps, 10))
                saved, streams[:] = list(streams), []
                tg = timed(step_gather, kg) / kg
                tg0 = timed(s["gather"], kg) / kg
                streams[:] = saved
                gather_res = {"ms_per_step_compute_plus_gather": tg, "ms_gather_alone": tg0, "bytes_into_rank0": s["gather_bytes"]}
        step_e2e()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        self.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_plain_ms = e2e_compact_ms = None
        if "step_e2e_plain" in s:
            s["step_e2e_plain"]()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                s["step_e2e_plain"]()
            self.barrier()
            e2e_plain_ms = (time.perf_counter() - t0) * 1e3 / 2
        if "step_e2e_compact" in s:
            s["step_e2e_compact"]()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                s["step_e2e_compact"]()
            self.barrier()
            e2e_compact_ms = (time.perf_counter() - t0) * 1e3 / 2
        errs = int((s["err"] != 0).sum().item()) if s["err"] is not None else 0
        fm = [formats[k]["ms_per_step"] for k in sorted(formats)]
        gm = [gather_res["ms_per_step_compute_plus_gather"], gather_res["ms_gather_alone"]] if gather_res else []
        t = torch.tensor([ms, e2e_ms, float(errs)] + fm + gm, dtype=torch.float64, device="cuda")
        if self.dist is not None:
            mx = t.clone()
            self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)       # time = max over ranks
            sm = t.clone()
            self.dist.all_reduce(sm, op=self.dist.ReduceOp.SUM)       # final gather of the error counts
            ms, e2e_ms, errs = float(mx[0]), float(mx[1]), int(sm[2])
            for i, k in enumerate(sorted(formats)):
                formats[k]["ms_per_step"] = float(mx[3 + i])
            if gather_res:
                gather_res["ms_per_step_compute_plus_gather"] = float(mx[3 + len(fm)])
                gather_res["ms_gather_alone"] = float(mx[4 + len(fm)])
        ms_per_step = ms / steps
        units_total = s.get("units_total", s["units"] * self.world)          # units all ranks process per step
        value = units_total / (ms_per_step * 1e-3)
        e2e_units = s.get("e2e_units", s["units"]) * self.world if "units_total" not in s else units_total
        e2e_value = e2e_units / (e2e_ms / e2e_steps * 1e-3)
        res = {
            "metric": METRICS[wl][0], "value": value, "unit": METRICS[wl][1], "ms_per_step": ms_per_step,
            "dtype": s["dtype"], "units_per_step_per_gpu": s["units"], "units_per_step": units_total, "error_rows": errs,
            "e2e": {"value": e2e_value, "unit": METRICS[wl][1], "h2d_bytes_per_step": s["h2d"], "d2h_bytes_per_step": s["d2h"],
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps},
            "gpu_launches": s["launches"] * steps,
        }
        if wl == "ecdsa":
            res["e2e"]["format"] = ("derived compact (p2e_tape_witness_compact, wide == NULL: u32 narrow slots, 34-bit carries as u32 + packed "
                                    "2-bit tops; check_sum slots rebuilt on the host from the carries; written by the chain kernels directly)")
            res["e2e"]["bytes_are"] = "per rank (every rank ships its shard over its own PCIe link)"
            if e2e_compact_ms is not None:
                res["e2e_compact"] = {"value": units_total / (e2e_compact_ms * 1e-3), "unit": METRICS[wl][1],
                                      "d2h_bytes_per_step": s["d2h_compact"], "ms_per_step": e2e_compact_ms,
                                      "format": "compact (u32 narrow slots, u64 check_sum slots shipped, packed carries)"}
            if e2e_plain_ms is not None:
                res["e2e_u64"] = {"value": units_total / (e2e_plain_ms * 1e-3), "unit": METRICS[wl][1],
                                  "d2h_bytes_per_step": s["d2h_plain"], "ms_per_step": e2e_plain_ms,
                                  "format": "plain (p2e_ecdsa_witness: u64 per slot)"}
            res["device_format"] = "derived compact (p2e_ecdsa_witness_compact_dev, wide == NULL): %d B / witness" % (s["narrow_rows"] * 4)
            for k in formats:
                formats[k]["value"] = units_total / (formats[k]["ms_per_step"] * 1e-3)
            res["device_formats"] = formats
            if gather_res:
                gather_res["value_with_gather"] = units_total / (gather_res["ms_per_step_compute_plus_gather"] * 1e-3)
                gather_res["gather_GBps_into_rank0"] = gather_res["bytes_into_rank0"] / (gather_res["ms_gather_alone"] * 1e-3) / 1e9
                gather_res["note"] = ("one batch at a time (no pipelining across steps): shard kernels, then dist.gather of the compact shard "
                                      "into rank 0's HBM; rank 0's NVLink ingress bounds a gathered witness stream at 900 GB/s / bytes per witness")
                res["gather"] = gather_res
        if streams:
            res["streams"] = len(streams)
            res["single_stream_ms_per_step"] = latency_ms
            if latency2_ms is not None:
                res["single_stream_ms_per_step_two_lane_replay"] = latency2_ms
        if wl == "ecdsa" and kernel_ms:
            names = self.eng.CHAIN_KERNELS
            res["kernel_ms"] = {nm: round(v, 4) for nm, v in zip(names, kernel_ms)}
            k_expand = kernel_ms[names.index("k_chain_expand")]
            achieved = s["alg_bytes"] / (k_expand * 1e-3) / 1e9
            # traffic: dram read + write of this kernel for 4096 signatures from the ncu --set full capture of this build
            # (profiles/r02y_ncu_chain_kernels.txt: 219 MB + 1 003 MB), scaled to this launch's signatures
            res["roofline_hbm_expand"] = {"bound": "hbm", "kernel": "k_chain_expand<compact>", "achieved": achieved, "peak": self.peak_gbs,
                                          "traffic": int((219.446e6 + 1003.401e6) * s["units"] / 4096),
                                          "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02y_ncu_chain_kernels.txt",
                                          "unit": "GB/s", "frac": achieved / self.peak_gbs, "peak_kind": self.peak_kind,
                                          "algorithmic_bytes_per_launch": s["alg_bytes"], "kernel_ms": k_expand,
                                          "share_of_step": k_expand / sum(kernel_ms)}
            res["_kernel_ms"] = kernel_ms
        else:
            achieved = s["alg_bytes"] / (ms_per_step * 1e-3) / 1e9
            res["roofline"] = {"bound": "hbm", "kernel": s["kernel"], "achieved": achieved, "peak": self.peak_gbs, "unit": "GB/s",
                               "frac": achieved / self.peak_gbs, "peak_kind": self.peak_kind,
                               "traffic": None,
                               "algorithmic_bytes_per_launch": s["alg_bytes"]}
        if sampler is not None:
            res["clocks"] = sampler.summary()
        del s
        torch.cuda.empty_cache()
        return res


# Algorithmic integer work of one verify_secp256k1_message_circuit witness (batched-inversion
# algorithm): explicit mul_nonnative 1087 x 2 (x*y and q*m) + ~311 curve ops x 14 (Jacobian replay) + 3 x 624
# (Montgomery's trick over the Z's and denominators) + 503 (one Fermat ladder) = 8903 modmul-equivalents of 81 limb
# products (one IMAD.WIDE each): 0.721 M IMAD.WIDE per signature.  The roofline divides THIS figure, not executed
# instructions (the kernels execute fewer: 64 products per multiplication on 8 x 32-bit words).
ALG_MODMUL_EQ_PER_WITNESS = 2174 + 311 * 14 + 3 * 624 + 503
ALG_IMAD_WIDE_PER_WITNESS = ALG_MODMUL_EQ_PER_WITNESS * 81


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("P2E_BENCH_WORKLOAD", "ecdsa"), choices=list(WORKLOAD_NAMES))
    ap.add_argument("--scaling", default=os.environ.get("P2E_BENCH_SCALING", "auto"), choices=["auto", "strong", "weak"],
                    help="N > 1, ecdsa workload: strong = the 4096 signatures of a step are sharded across the ranks (BASELINE.json "
                         "configs[2]; default), weak = 4096 per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "auto":
        args.scaling = "strong" if (world > 1 and args.workload == "ecdsa") else "weak"
    if args.impl == "reference":
        return run_reference(args)

    # stdout must carry exactly ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1,
    # so everything until the final print goes to stderr at the file-descriptor level
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    b = Bench(args)
    wl = args.workload
    strong = args.scaling == "strong" and b.world > 1 and wl == "ecdsa"
    main_res = b.measure(wl, args.steps, args.warmup, e2e_steps=max(2, min(args.steps, 5)), sample_clocks=True)
    secondary = {}
    if not args.no_secondary:
        for w2 in ("mulmod", "mulmod_n", "eval", "evalcs", "evalfc", "evalcsfc", "evalrc", "evalrcfc", "evalcmp", "evaladd", "lde", "genstream", "genstream_ecdsa", "rcrows", "glv", "ecdsa"):
            if w2 != wl and w2 in WORKLOAD_NAMES:
                try:
                    r2 = b.measure(w2, steps=min(args.steps, 20) if w2 not in ("glv", "genstream_ecdsa") else 3, warmup=3, e2e_steps=2)
                except Exception as e:      # noqa: BLE001  (e.g. not enough free HBM for the 32 GB glv witness)
                    secondary[w2] = {"error": str(e)[:200]}
                    continue
                r2.pop("_kernel_ms", None)
                secondary[w2] = r2
    d2h_probe = None
    weak_line = None
    if wl == "ecdsa" and b.world > 1:
        d2h_probe = b.d2h_probe()
        if strong:
            # the weak line beside the strong one: every rank its own 4096 signatures per step
            args.scaling = "weak"
            try:
                rw = b.measure(wl, min(args.steps, 48), 3, e2e_steps=1)
                weak_line = {"value": rw["value"], "unit": rw["unit"], "ms_per_step": rw["ms_per_step"], "units_per_step": rw["units_per_step"],
                             "units_per_step_per_gpu": rw["units_per_step_per_gpu"], "streams": rw.get("streams")}
            except Exception as e:          # noqa: BLE001
                weak_line = {"error": str(e)[:200]}
            args.scaling = "strong"
    if b.rank == 0:
        upg = main_res["units_per_step_per_gpu"]
        line = {
            "metric": main_res["metric"], "value": main_res["value"], "unit": main_res["unit"], "n_gpus": b.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": main_res["dtype"], "data": "synthetic",
            "config": {"workload": workload_label(wl, strong, upg),
                       "units_per_step": main_res["units_per_step"], "units_per_step_per_gpu": upg,
                       "seed": hex(b.W.SEED), "error_rows": main_res["error_rows"],
                       "l2": "per-step inputs+outputs exceed the 126 MB L2 (no flush needed)",
                       "streams": main_res.get("streams", 1),
                       "pipelining": "steps are independent batches enqueued round-robin on `streams` CUDA streams through the async *_dev "
                                     "entry point; ms_per_step = timed region / steps; single_stream_ms_per_step = one batch at a time",
                       "single_stream_ms_per_step": main_res.get("single_stream_ms_per_step"),
                       "single_stream_ms_per_step_two_lane_replay": main_res.get("single_stream_ms_per_step_two_lane_replay"),
                       "device_format": main_res.get("device_format")},
            "e2e": main_res["e2e"], "e2e_compact": main_res.get("e2e_compact"), "e2e_u64": main_res.get("e2e_u64"), "gpu_launches": main_res["gpu_launches"],
            "clocks": main_res.get("clocks"),
        }
        for k in ("kernel_ms", "device_formats", "gather", "roofline_hbm_expand"):
            if k in main_res:
                line[k] = main_res[k]
        if d2h_probe is not None:
            line["d2h_probe"] = d2h_probe
        if weak_line is not None:
            line["weak_scaling"] = weak_line
        try:
            imad = b.eng.bench_imad(False)
            imadw = b.eng.bench_imad(True)
            line["int32_peak"] = {"imad_gops": imad, "imad_wide_gops": imadw, "how": "p2e_bench_imad: 8 independent chains/thread, best of 5"}
            if wl == "ecdsa":
                # THE roofline of the headline: whole job (all chain kernels, pipelined step) against the measured IMAD.WIDE
                # peak, on the algorithmic limb products of SURVEY.md section 8d (0.721 M IMAD.WIDE per signature)
                job = ALG_IMAD_WIDE_PER_WITNESS * upg
                ach = job / (main_res["ms_per_step"] * 1e-3) / 1e9
                km = main_res.get("_kernel_ms") or []
                names = b.eng.CHAIN_KERNELS
                line["roofline"] = {"bound": "int32", "kernel": "whole job: k_prep + k_chain_replay + k_chain_invert_* + k_chain_affine + k_chain_expand "
                                    "(pipelined step)", "achieved": ach, "peak": imadw, "unit": "G IMAD.WIDE/s", "frac": ach / imadw,
                                    "peak_kind": "measured in this run (p2e_bench_imad; not in MEASURED_PEAKS.json)", "traffic": None,
                                    "algorithmic_ops_per_step_per_gpu": job, "algorithmic_imad_wide_per_witness": ALG_IMAD_WIDE_PER_WITNESS,
                                    "single_batch_kernel_ms": {nm: round(v, 4) for nm, v in zip(names, km)},
                                    "single_batch_frac": (job / (sum(km) * 1e-3) / 1e9 / imadw) if km else None}
                if km:
                    k_replay = km[names.index("k_chain_replay")]
                    work = 3263 * 81 * upg
                    line["roofline_int32_replay"] = {"bound": "int32", "kernel": "k_chain_replay (main launch, one batch alone)", "achieved": work / (k_replay * 1e-3) / 1e9,
                                                     "peak": imadw, "unit": "G IMAD.WIDE/s", "frac": work / (k_replay * 1e-3) / 1e9 / imadw,
                                                     "algorithmic_ops_per_launch": work, "kernel_ms": k_replay}
            else:
                line["roofline"] = main_res["roofline"]
            for key, nm in (("mulmod", "mulmod_int32_frac"), ("mulmod_n", "mulmod_n_int32_frac")):
                if key in secondary and "value" in secondary[key]:
                    line.setdefault("secondary_notes", {})[nm] = 162 * secondary[key]["value"] / b.world / 1e9 / imadw
        except Exception as e:          # noqa: BLE001
            line["int32_peak"] = {"error": str(e)}
            line.setdefault("roofline", main_res.get("roofline"))
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and b.world == 1:
            try:
                from oracle import cpu_baseline
                cb = cpu_baseline.run(wl, steps=1, warmup=0, budget_s=15.0)
                line["cpu_baseline"] = {"value": cb["value"], "unit": cb["unit"], "cores": cb["cores"], "kind": cb["kind"],
                                        "sample": cb["sample"]}
                if "batched" in cb:
                    line["cpu_baseline_batched"] = cb["batched"]
            except Exception as e:      # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "error": str(e)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    b.eng.close()
    if b.dist is not None:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
