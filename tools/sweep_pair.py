"""Development probe (GPU box): pair-kernel time at S64 for each launch plan."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
plans = json.loads(sys.argv[2]) if len(sys.argv) > 2 else [[r, s] for r in (1, 2, 4) for s in (0, 4, 8, 15, 30)]
w = W.build(name)
n = w.n_vertices
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
peak, mhz = capi.fp64_peak()
print("dfma peak %.3e /s" % peak)
with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n) as ctx:
    ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
    for rows, splits in plans:
        if rows < 0:  # symmetric kernel: [-rows_per_thread, threads (+1: tighter register cap)]
            ctx.set_pair_mode(2)
            ctx.set_pair_plan(-rows, splits)
        else:
            ctx.set_pair_mode(1)
            ctx.set_pair_plan(rows, splits)
        ctx.force_init()
        ctx.step(3)
        ctx.pair_timing(True)
        ctx.step(10)
        ms, cnt = ctx.pair_timing_read()
        ctx.pair_timing(False)
        g = ctx.pair_geometry()
        pairs = n * (n - 1)
        per_pair = 16 if rows < 0 else 28
        print("plan=%s mode=%d geom=%s  %.4f ms  %.3e ordered pairs/s  fp64 pipe frac=%.3f" % (
            [rows, splits], ctx.pair_mode(), g, ms, pairs / ms * 1e3, pairs * per_pair / (ms * 1e-3) / peak), flush=True)
