import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from np_shape_lab_b200 import capi, workloads as W
w = W.build("S64_disc")
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
    ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
    ctx.force_init()
    ctx.step(5)
    ctx.pair_timing(True); ctx.step(20); ms, _ = ctx.pair_timing_read(); ctx.pair_timing(False)
    print(os.environ.get("NPSL_LIB", "current"), "pair kernel %.4f ms" % ms, "step %.4f" % (ctx.step_timing(50, flush_l2=True) / 50))
