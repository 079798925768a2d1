"""Development probe (GPU box): energy conservation of the extended system vs time step on the
synthetic icospheres, and GPU-vs-oracle trajectories on a mid-size one."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, mesh as M

def cons(e):
    return e["energy"] + e["mesh_bath_ke"] + e["mesh_bath_pe"] + e["ions_bath_ke"] + e["ions_bath_pe"]

def run(D, dt, steps, c=0.005):
    m = M.icosphere(D)
    inp = M.PhysicalInputs(q=600.0, c=c, dt=dt)
    p = M.derive_params(inp, m)
    q = M.assign_charges(m, inp.q)
    mb, ib = M.make_thermostats(inp, m.n_vertices)
    P = capi.make_params(p, mb, ib)
    with capi.Context(m.edges, m.faces, p["l0"], P, n_vertices=m.n_vertices) as ctx:
        ctx.upload_state(m.pos, np.zeros_like(m.pos), q)
        ctx.force_init()
        e0 = ctx.energy()
        out = []
        done = 0
        for k in steps:
            ctx.step(k - done); done = k
            e = ctx.energy()
            out.append((k, cons(e) / cons(e0) - 1.0, e["kinetic"], e["total_area"]))
    return e0, out

import json
cases = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [[64, 3e-4, 0.005, [20, 100, 300, 1000]]]
for D, dt, c, steps in cases:
    t = time.time()
    try:
        e0, out = run(D, dt, steps, c)
        print("D=%d dt=%g c=%g E0=%.6g" % (D, dt, c, cons(e0)), " ".join("k=%d dE/E=%.2e KE=%.3g A=%.5g" % o for o in out), "%.1fs" % (time.time() - t), flush=True)
    except Exception as ex:
        print("D=%d dt=%g c=%g FAILED %s" % (D, dt, c, ex), flush=True)
