"""GPU-box measurement on ONE GPU: timeline of the last k_pair_sym launch of rank r of a w-way split (NPSL_EMULATE_SHARD,
NPSL_SYM_TRACE): when the CTAs start, how long the units of each width run, how many CTAs are resident over time and how
much CTA-time the ragged start and end of the launch leave unused.
    python tools/trace_timeline.py [workload] [world] [rank] [NPSL_*=value ...]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for kv in sys.argv[4:]:
    os.environ[kv.split("=")[0]] = kv.split("=")[1]
path = tempfile.mktemp(prefix="npsl_trace_")
os.environ["NPSL_SYM_TRACE"] = path
os.environ["NPSL_EMULATE_SHARD"] = "%d/%d" % (rank, world)
from np_shape_lab_b200 import capi, workloads as W

w = W.build(name)
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
units, geo = capi.plan_pair_units(w.n_vertices, rank, world)
with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
    ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
    ctx.force_init()
    ctx.step(5)
    ctx.sync()
tr = np.loadtxt(path)
os.unlink(path)
if os.environ.get("NPSL_TRACE_KEEP"):  # raw rows + the units' widths for offline analysis
    np.save(os.environ["NPSL_TRACE_KEEP"], np.concatenate([tr, (units[: len(tr), 3] - units[: len(tr), 2])[:, None]], axis=1))
sm, t0, t1 = tr[:, 3].astype(int), tr[:, 4], tr[:, 5]
begin, end = t0.min(), t1.max()
t0, t1 = (t0 - begin) * 1e-3, (t1 - begin) * 1e-3  # us
span = (end - begin) * 1e-3
cols = (units[:, 3] - units[:, 2])[: len(t0)]
tiles = np.ceil(cols / geo["column_granularity"]).astype(int)
print("%s, rank %d of %d: %d units, %d tiles, launch span %.1f us (first CTA start to last CTA end)" % (
    name, rank, world, len(t0), tiles.sum(), span))
pro, epi = (tr[:, 6] - tr[:, 4]) * 1e-3, (tr[:, 5] - tr[:, 7]) * 1e-3  # start -> rows + table there; pairs done -> end
for k in sorted(set(tiles)):
    m = tiles == k
    d = (t1 - t0)[m]
    print("  units of %d tile(s): %5d, duration median %.1f us (min %.1f, max %.1f), %.2f us per tile; prologue median %.2f us "
          "(max %.2f), epilogue median %.2f us (max %.2f)" % (k, len(d), np.median(d), d.min(), d.max(), np.median(d) / k,
                                                          np.median(pro[m]), pro[m].max(), np.median(epi[m]), epi[m].max()))
# resident CTAs over time
ev = np.concatenate([np.stack([t0, np.ones_like(t0)], 1), np.stack([t1, -np.ones_like(t1)], 1)])
ev = ev[np.argsort(ev[:, 0], kind="stable")]
act = np.cumsum(ev[:, 1])
peak = act.max()
busy = float(np.sum((t1 - t0)))
print("  peak resident CTAs %d; CTA-time used %.0f us = %.1f %% of peak x span" % (peak, busy, 100 * busy / (peak * span)))
for frac in (1.0, 0.75, 0.5, 0.25):
    k = np.where(act >= frac * peak)[0]
    print("  >= %3.0f %% of the peak resident from %.1f to %.1f us" % (100 * frac, ev[k[0], 0], ev[k[-1] + 1, 0] if k[-1] + 1 < len(ev) else span))
# gap between a CTA's end and the start of the CTA that takes its place on the same SM (block scheduler + resource hand-over)
gaps = []
for sidx in np.unique(sm):
    m = np.where(sm == sidx)[0]
    starts = np.sort(t0[m])
    ends = np.sort(t1[m])
    later = starts[starts > 1.0]  # replacements (the first wave starts at 0)
    for k, st in enumerate(later):
        if k < len(ends) and ends[k] <= st:
            gaps.append(st - ends[k])
gaps = np.array(gaps)
if len(gaps):
    print("  hand-over on an SM (k-th CTA to end -> k-th replacement to start): median %.2f us, 90th percentile %.2f us, %d hand-overs"
          % (np.median(gaps), np.percentile(gaps, 90), len(gaps)))
per_sm_end = np.array([t1[sm == s].max() for s in np.unique(sm)])
per_sm_busy = np.array([np.sum((t1 - t0)[sm == s]) for s in np.unique(sm)])
print("  per SM: last CTA ends at %.1f .. %.1f us (median %.1f); CTA-time per SM %.0f .. %.0f us" % (
    per_sm_end.min(), per_sm_end.max(), np.median(per_sm_end), per_sm_busy.min(), per_sm_busy.max()))
