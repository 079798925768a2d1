"""compute-sanitizer target (GPU box): a small charged mesh through both pair kernels, a few steps each.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W
w = W.build("S8_disc")
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
rng = np.random.default_rng(0)
pos = w.mesh.pos + 0.01 * rng.standard_normal(w.mesh.pos.shape)
for mode in (1, 2):
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
        ctx.upload_state(pos, np.zeros_like(pos), w.q)
        ctx.set_pair_mode(mode)
        ctx.force_init()
        ctx.step(3)
        e = ctx.energy()
        print("mode", mode, "E", e["energy"])
