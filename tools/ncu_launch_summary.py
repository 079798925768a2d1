"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/ncu_launch_summary.py gpurun_out/r01_launches_v3.csv > profiles/r01_launches_v3_summary.txt"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
tot, cnt = {}, {}
for r in rows:
    if r is hdr or r[mi] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = r[ki].split("(")[0].replace("npsl::", "")
    tot[name] = tot.get(name, 0.0) + us
    cnt[name] = cnt.get(name, 0) + 1
all_us = sum(tot.values())
print("per-kernel totals of %s (cold cache, serialised under ncu; shares matter, not absolutes)\n" % sys.argv[1])
for name in sorted(tot, key=lambda k: -tot[k]):
    print("%-34s launches %4d  total %11.3f us  mean %10.3f us  share %5.2f%%" % (name, cnt[name], tot[name], tot[name] / cnt[name], 100 * tot[name] / all_us))
