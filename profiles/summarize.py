"""Turns an ncu report (gpurun_out/*.ncu-rep, `ncu --set full`) into a compact per-kernel summary.
Usage: python profiles/summarize.py gpurun_out/r01_prof_top.ncu-rep > profiles/r01_ncu_top_kernels.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA (IMAD) pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "FMA-heavy pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path}  (cold-cache, serialised replays: compare shares, not absolutes)")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n== {name[:110]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:32s} {r[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        top = ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:6])
        print(f"   {'stall reasons (per issue)':32s} {top}")


if __name__ == "__main__":
    main(sys.argv[1])
