#!/usr/bin/env python
"""sass_mix.py <object-or-so> [name-filter] -- static SASS instruction mix per function (cuobjdump -sass).

Used on CPU before spending GPU time: instruction count, IMAD.WIDE count (FMA-heavy pipe, 4 issue cycles per
warp instruction on B200), other IMAD, ALU (IADD3/LOP3/SHF/...), memory, and calls.
"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn = None
    stats = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            stats[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            c = stats[fn]
            c["total"] += 1
            base = op.split(".")[0]
            if op.startswith("IMAD.WIDE"):
                c["imad_wide"] += 1
            elif base == "IMAD":
                c["imad_mov" if (".MOV" in op or ".IADD" in op or ".SHL" in op) else "imad"] += 1
            elif base in ("IADD3", "LOP3", "SHF", "ISETP", "SEL", "LEA", "IADD", "PRMT", "IABS", "PLOP3", "FLO", "POPC", "VIADD", "VIMNMX", "MOV", "CS2R", "UIADD3", "ULOP3", "USHF", "UMOV", "IMNMX", "SGXT", "BMSK", "R2P", "P2R"):
                c["alu"] += 1
            elif base in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU", "ATOMG", "RED", "LD", "ST", "ATOM"):
                c["mem"] += 1
                if base in ("LDL", "STL"):
                    c["local"] += 1
            elif base in ("CALL", "RET", "BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "SHFL", "NOP"):
                c[base.lower()] += 1
            else:
                c["other"] += 1
    cf = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
    keys = ["total", "imad_wide", "imad", "imad_mov", "alu", "mem", "local", "shfl", "call", "bra", "other"]
    print(" ".join("%9s" % k for k in keys), " function")
    for (fn, c), nm in zip(stats.items(), cf):
        if flt and flt not in nm:
            continue
        print(" ".join("%9d" % c[k] for k in keys), " ", nm[:110])


if __name__ == "__main__":
    main()
