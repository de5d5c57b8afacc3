#!/usr/bin/env python
"""ncu_segments.py <report.ncu-rep> <kernel regex> -- executed warp instructions of one kernel split at the RET
boundaries of its out-of-line device functions (ncu --page source, -lineinfo builds), with the opcode mix of each
segment.  Tells which core of a kernel the issue slots go to."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, isrc, ie, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = []
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ia].startswith("0x"):
            if ins:
                break               # next kernel instance
            continue
        ins.append((r[isrc].strip(), int(r[ie] or 0), int(r[iss] or 0)))
    tot = sum(x[1] for x in ins)
    tots = max(1, sum(x[2] for x in ins))
    print("kernel", rx, "total warp instructions", tot)
    start, bounds = 0, []
    for i, (s, e, sm) in enumerate(ins):
        if s.startswith("RET"):
            bounds.append((start, i))
            start = i + 1
    bounds.append((start, len(ins) - 1))
    for a, b in bounds:
        seg = ins[a:b + 1]
        e = sum(x[1] for x in seg)
        if e < 0.005 * tot:
            continue
        ops = collections.Counter()
        for s, ee, _ in seg:
            m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
            op = m.group(1) if m else s
            ops["IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0]] += ee
        print(f"seg {a:5d}-{b:5d} static {b - a + 1:5d} exec {e:10d} ({e / tot:.3f}) samples {sum(x[2] for x in seg) / tots:.3f} "
              f"max exec/inst {max(x[1] for x in seg)}  ", ", ".join(f"{k}:{v / e:.2f}" for k, v in ops.most_common(9)))


if __name__ == "__main__":
    main()
