#!/usr/bin/env python
"""Instruction mix of the loops of one kernel in a cuobjdump -sass listing (CPU-side check before spending GPU time).

    cuobjdump -sass genomics_rs_b200/csrc/_obj/gx_fill_k8_r4_c0.o > /tmp/x.sass
    python tools/sass_loops.py /tmp/x.sass 'ILi8ELi4ELb0ELb1ELi0ELb1ELb0' [--cells 64]

For every backward branch the body [target, branch] is summarised: instructions per pipe class and, with --cells N
(DP cells per iteration of the loop you are looking at), instructions per cell.  The ALU-pipe / FMA-pipe split follows
the issue-rate probe (tools/k0_probe.cu): VIADDMNMX, VIMNMX*, ISETP, SEL, LOP3, IADD3, SHF, PRMT, MOV ... issue on the
ALU pipe; IMAD* on the FMA pipe; LDS/STS/LDG/STG/SHFL on the LSU; the rest is control."""
import re
import sys
from collections import Counter

ALU = ("VIADDMNMX", "VIMNMX", "ISETP", "SEL", "LOP3", "IADD3", "IADD", "SHF", "PRMT", "MOV", "LEA", "VIADD", "IABS", "POPC", "FLO",
       "PLOP3", "P2R", "R2P", "CS2R", "S2R", "BMSK", "SGXT", "I2I", "VABSDIFF", "IMNMX", "FSEL", "FMNMX", "VOTE", "VOTEU")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")
LSU = ("LDS", "STS", "LDG", "STG", "SHFL", "LD", "ST", "ATOM", "RED", "LDGSTS", "LDSM", "MATCH", "LDC", "LDCU", "UBLKCP", "SYNCS", "CCTL", "MEMBAR", "ERRBAR")


def classify(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UBLKCP",):
        return "uniform"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in LSU:
        return "lsu"
    return "ctl"


def main():
    path, pat = sys.argv[1], sys.argv[2]
    cells = None
    if "--cells" in sys.argv:
        cells = int(sys.argv[sys.argv.index("--cells") + 1])
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "Function :" in l and pat in l)
    end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
    ins = []   # (addr, op, text)
    rx = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);")
    for l in lines[start:end]:
        m = rx.search(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), l.strip()))
    addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
    print(f"{lines[start].strip()}: {len(ins)} instructions")
    loops = []
    for i, (a, op, text) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text.split("BRA", 1)[1])
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr_idx:
                    loops.append((addr_idx[t], i))
    for lo, hi in sorted(loops, key=lambda x: x[1] - x[0], reverse=True):
        body = ins[lo:hi + 1]
        if len(body) < 40:
            continue
        cls = Counter(classify(op) for _, op, _ in body)
        ops = Counter(op.split(".")[0] for _, op, _ in body)
        line = f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(cls.items()))
        if cells:
            line += f"   per cell ({cells}): total {len(body) / cells:.2f} alu {cls['alu'] / cells:.2f} fma {cls['fma'] / cells:.2f} lsu {cls['lsu'] / cells:.2f}"
        print(line)
        print("    " + " ".join(f"{k}:{v}" for k, v in ops.most_common(24)))


if __name__ == "__main__":
    main()
