"""Profiling target (GPU box): force_calculation_init + a few MD steps of a named workload, nothing else.
    python tools/profile_target.py [workload] [steps] [rows_per_thread] [col_splits]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = W.build(name)
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
    ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
    if len(sys.argv) > 4:
        ctx.set_pair_plan(int(sys.argv[3]), int(sys.argv[4]))
    n_ions = int(os.environ.get("NPSL_PROFILE_IONS", "0"))
    if n_ions:  # explicit counterions on a shell around the membrane (timing only)
        rng = np.random.default_rng(1)
        d = rng.standard_normal((n_ions, 3))
        ctx.set_ions(d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(1.1, 1.6, size=(n_ions, 1)), None, -np.ones(n_ions),
                     0.06, 0.5 * (w.params["avg_edge_length"] + 0.06))
    ctx.force_init()
    ctx.step(steps)
    if os.environ.get("NPSL_PROFILE_SEGMENTS"):
        print(ctx.step_profile(20))
        ctx.sync()
        print("graph ms/step", ctx.step_timing(200, flush_l2=True) / 200)
    e = ctx.energy()
    print("ok", name, steps, "steps, E =", e["energy"], ctx.pair_geometry())
