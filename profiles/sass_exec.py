#!/usr/bin/env python3
"""Executed-path instruction count of the streaming kernel's loop: the loop body minus the cold
blocks (global-table fallback, misaligned-index Philox path), per iteration (= 4 boards)."""
import collections, re, subprocess, sys
LIB = sys.argv[1] if len(sys.argv) > 1 else "reinforcement-learning-2048_b200/b2048/libb2048.so"
ALU = ("LOP3", "SHF", "PRMT", "ISETP", "SEL", "IADD3", "VIADD", "LEA", "VIMNMX", "PLOP3", "MOV", "IADD")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
body = next(f for f in re.split(r"\n\s*Function : ", sass) if "step_stream_kernelILb0" in f.split("\n")[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+((?:@!?U?P\d\s+)?)([A-Z0-9_.]+)(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), m.group(3), m.group(4)))
# loop = largest backward branch
lo, hi = max(((int(re.search(r"0x([0-9a-f]+)", r).group(1), 16), a) for a, p, op, r in ins
              if op.startswith("BRA") and re.search(r"0x([0-9a-f]+)", r) and int(re.search(r"0x([0-9a-f]+)", r).group(1), 16) < a),
             key=lambda t: t[1] - t[0])
# cold blocks: forward predicated branches whose fall-through contains a CALL, or uniform BRA.U skipping >40 instrs
cold = []
for i, (a, p, op, r) in enumerate(ins):
    if lo <= a <= hi and op.startswith("BRA") and (p or re.match(r"\s*!?UP\d", r)):
        t = re.search(r"0x([0-9a-f]+)", r)
        if t and int(t.group(1), 16) > a:
            tgt = int(t.group(1), 16)
            inner = [x for x in ins if a < x[0] < tgt]
            if any(x[2].startswith("CALL") for x in inner) or len(inner) > 40:
                cold.append((a + 16, tgt))
c = collections.Counter()
wide = 0
for a, p, op, r in ins:
    if lo <= a <= hi and not any(x <= a < y for x, y in cold):
        c[op.split(".")[0]] += 1
        wide += bool(re.match(r"IMAD\.(WIDE|HI)", op))
alu = sum(v for k, v in c.items() if k in ALU)
print(f"loop 0x{lo:x}..0x{hi:x}, cold blocks {[(hex(x), hex(y)) for x, y in cold]}")
print(f"executed per iteration (4 boards): {sum(c.values())} instructions, ALU pipe {alu} ({alu / 4:.1f}/board), "
      f"IMAD {c['IMAD']} (wide/hi {wide}), LDS {c['LDS']}")
print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common()))
