"""Host-only model of how the units of k_pair_sym fill the SMs of one B200 (no GPU needed): list scheduling in launch
order onto 148 SMs x 4 resident CTAs, where the two warps of a CTA share their SM sub-partitions with the warps of ONE
other CTA and the older CTA of such a pair gets the dispatch slots it can use -- the two speed classes the per-CTA trace
shows (tools/trace_timeline.py):

    older CTA of a pair   9.8 us per 32-column group   (R_FAST)
    younger CTA          21.0 us                        (R_SLOW; it turns "older" when its partner leaves)
    CTA alone in a pair   9.5 us                        (R_ALONE)

Calibrated on two measured launch spans at S64 (column chunks of 288 / 192 / 32 columns): 216.0 us modelled against
217.1 us measured for one rank of 8, 1599.1 against 1599.0 us for the full launch. It ranks unit mixes well enough to
say what to avoid (many mid-sized units: one that is dispatched late runs slow for most of its life and ends the
launch), not well enough to replace the measurement: the final mix was chosen on the GPU (profiles/r02_unit_geometry.txt).

    python tools/unit_schedule_model.py [n_vertices] [world]            # the library's current mix
    python tools/unit_schedule_model.py 40962 8 search                   # rank a grid of mixes (NPSL_SYM_* knobs)
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi

R_FAST, R_SLOW, R_ALONE, OVERHEAD_US = 9.8, 21.0, 9.5, 1.0


def units_for(n, env, world, rank=0):
    """Widths (in 32-column groups) of the units of one rank, in launch order, for the NPSL_SYM_* knobs in env."""
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        u, g = capi.plan_pair_units(n, rank, world)
    finally:
        for k in env:
            os.environ.pop(k, None)
    return np.ceil((u[:, 3] - u[:, 2]) / g["column_granularity"]).astype(int), g


def launch_span(tiles, n_sm=148):
    """Time from the first CTA's start to the last CTA's end, us."""
    work = [t + OVERHEAD_US / R_FAST for t in tiles]  # in 32-column groups; the per-unit overhead folded in
    n_pairs = n_sm * 2                                # two CTA pairs per SM (4 resident CTAs)
    pairs = [[] for _ in range(n_pairs)]              # per pair: [remaining work, launch index]
    nxt = 0
    for _ in range(2):
        for p in range(n_pairs):
            if nxt < len(work):
                pairs[p].append([work[nxt], nxt])
                nxt += 1

    def rates(p):
        c = pairs[p]
        if len(c) == 1:
            return [1.0 / R_ALONE]
        return [1.0 / R_FAST, 1.0 / R_SLOW] if c[0][1] < c[1][1] else [1.0 / R_SLOW, 1.0 / R_FAST]

    now, span = 0.0, 0.0
    version, last, heap = [0] * n_pairs, [0.0] * n_pairs, []

    def push(p):
        version[p] += 1
        if pairs[p]:
            r = rates(p)
            heapq.heappush(heap, (now + min(c[0] / r[i] for i, c in enumerate(pairs[p])), p, version[p]))

    for p in range(n_pairs):
        push(p)
    while heap:
        t_next, p, v = heapq.heappop(heap)
        if v != version[p]:
            continue
        r = rates(p)
        for i, c in enumerate(pairs[p]):
            c[0] -= (t_next - last[p]) * r[i]
        last[p] = now = t_next
        done = [c for c in pairs[p] if c[0] <= 1e-9]
        pairs[p] = [c for c in pairs[p] if c[0] > 1e-9]
        for _ in done:
            span = max(span, now)
            if nxt < len(work):
                pairs[p].append([work[nxt], nxt])
                nxt += 1
        push(p)
    return span


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40962
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    if len(sys.argv) > 3 and sys.argv[3] == "search":
        res = []
        for ct in (6, 7, 8, 9, 10, 12):
            for ft in (2, 3, 4, 5, 6):
                for fp in (10, 15, 25, 35):
                    for sp in (3, 5, 10):
                        if ft >= ct or sp > fp:
                            continue
                        env = dict(NPSL_SYM_CHUNK_TILES=ct, NPSL_SYM_FINE_TILES=ft, NPSL_SYM_FINEST_TILES=1,
                                   NPSL_SYM_FINE_PCT=fp, NPSL_SYM_FINEST_PCT=sp)
                        t, g = units_for(n, env, world)
                        res.append((launch_span(list(t)), g["chunks"], ct, ft, fp, sp))
        for x in sorted(res)[:20]:
            print("span %.1f us | %3d chunks | widths %d / %d / 1 groups of 32 columns, tails %d %% / %d %%" % x)
        return
    t, g = units_for(n, {}, world)
    print("%d vertices, rank 0 of %d: %d units, %d groups of 32 columns, widths %s" % (
        n, world, len(t), t.sum(), {int(a): int(b) for a, b in zip(*np.unique(t, return_counts=True))}))
    both = 1.0 / (1.0 / R_FAST + 1.0 / R_SLOW)
    print("modelled launch span %.1f us (all CTA pairs busy all the time: %.1f us)" % (
        launch_span(list(t)), t.sum() * both / (148 * 2)))


if __name__ == "__main__":
    main()
