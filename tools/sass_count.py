"""Counts the FP64-pipe instructions per pair in the inner loops of the pair kernels (cuobjdump -sass of
libnpsl.so) and classifies them by the number of distinct register sources.
    python tools/sass_count.py > profiles/r01_pair_sass_count.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "np-shape-lab_b200", "libnpsl.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
FP64 = re.compile(r"^(DFMA|DMUL|DADD)\b")
for body in funcs:
    name = body.split("\n", 1)[0].strip()
    if "k_pair" not in name or "reduce" in name:
        continue
    ins = []
    for l in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
    loops = []
    for addr, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            seg = [x for x in ins if tgt <= x[0] <= addr]
            nf = sum(1 for a, x in seg if FP64.match(x))
            nm = sum(1 for a, x in seg if x.startswith("MUFU.RSQ64H"))
            if nm:
                loops.append((addr - tgt, tgt, addr, nf, nm, seg))
    if not loops:
        continue
    demangled = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    print(demangled)
    for span, tgt, addr, nf, nm, seg in sorted(loops)[:2]:  # the innermost loops (plain and diagonal path)
        cls = {}
        for a, x in seg:
            if FP64.match(x):
                srcs = x[len(x.split()[0]):].split(",")[1:]
                regs = set(re.sub(r"[-|]|\.reuse", "", q.strip()) for q in srcs if re.match(r"\s*-?\|?R\d+", q.strip()))
                cls[len(regs)] = cls.get(len(regs), 0) + 1
        other = len(seg) - nf
        print("  loop 0x%04x-0x%04x: %d pair evaluations (MUFU.RSQ64H), %d FP64-pipe instructions = %.1f per pair; "
              "%d other instructions; FP64 by distinct register sources %s" % (tgt, addr, nm, nf, nf / nm, other, dict(sorted(cls.items()))))
