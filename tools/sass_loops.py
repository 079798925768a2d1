"""Inner loops of a kernel in libnpsl.so: FP64-pipe / LDS / STS / MUFU / other instruction counts per loop
(cuobjdump -sass), FP64 instructions classified by the number of distinct register sources.
    python tools/sass_loops.py <substring of the mangled kernel name> [min FP64 per loop]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("NPSL_LIB") or os.path.join(ROOT, "np-shape-lab_b200", "libnpsl.so")
want = sys.argv[1]
min_fp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
FP64 = re.compile(r"^(DFMA|DMUL|DADD)\b")
for body in re.split(r"\n\s*Function : ", txt)[1:]:
    name = body.split("\n", 1)[0].strip()
    if want not in name:
        continue
    ins = []
    for l in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
    print(name, "--", len(ins), "instructions")
    for addr, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            seg = [x for x in ins if tgt <= x[0] <= addr]
            nf = sum(1 for a, x in seg if FP64.match(x))
            if nf < min_fp or len(seg) > 600:
                continue
            cls = {}
            for a, x in seg:
                if FP64.match(x):
                    srcs = x[len(x.split()[0]):].split(",")[1:]
                    regs = set(re.sub(r"[-|]|\.reuse", "", q.strip()) for q in srcs if re.match(r"\s*-?\|?R\d+", q.strip()))
                    cls[len(regs)] = cls.get(len(regs), 0) + 1
            cnt = lambda p: sum(1 for a, x in seg if x.startswith(p))
            print("  loop 0x%04x-0x%04x: %3d instructions, FP64 %3d %s, LDS %d, STS %d, MUFU %d, LDG %d, other %d"
                  % (tgt, addr, len(seg), nf, dict(sorted(cls.items())), cnt("LDS"), cnt("STS"), cnt("MUFU"), cnt("LDG"),
                     len(seg) - nf - cnt("LDS") - cnt("STS") - cnt("MUFU") - cnt("LDG")))
