#!/usr/bin/env python3
"""Opcode histogram of the hot loop of a kernel in libb2048.so (largest backward branch),
grouped by the pipe the instruction issues to (ALU / FMA / LSU / other), per loop iteration.

    python profiles/sass_stats.py step_stream_kernelILb0
"""
import collections
import re
import subprocess
import sys

LIB = "reinforcement-learning-2048_b200/b2048/libb2048.so"
ALU = ("LOP3", "SHF", "PRMT", "ISETP", "SEL", "IADD3", "VIADD", "LEA", "VIMNMX", "PLOP3", "IABS", "MOV", "SGXT", "BMSK", "P2R", "R2P", "IADD")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD")
LSU = ("LDS", "LDG", "STG", "STS", "LD", "ST")


def main():
    pat = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    body = next(f for f in funcs if pat in f.split("\n")[0])
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    best = None
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr:
                span = addr - int(t.group(1), 16)
                if best is None or span > best[0]:
                    best = (span, int(t.group(1), 16), addr)
    _, lo, hi = best
    loop = [i for i in ins if lo <= i[0] <= hi]
    hist = collections.Counter(op.split(".")[0] for _, op, _ in loop)
    wide = sum(1 for _, op, _ in loop if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"))
    groups = collections.Counter()
    for k, v in hist.items():
        g = "ALU" if k in ALU else "FMA" if k in FMA else "LSU" if k in LSU else "other"
        groups[g] += v
    print(f"{pat}: loop 0x{lo:x}..0x{hi:x}  {len(loop)} instructions per iteration")
    print("  by pipe:", dict(groups), f"(IMAD.WIDE/HI: {wide})")
    print("  " + "  ".join(f"{k}:{v}" for k, v in hist.most_common()))


if __name__ == "__main__":
    main()
