"""GPU-box measurement: device time of the symmetric pair kernel and of the whole step for the math / unroll variants of
k_pair_sym (NPSL_SYM_MATH, NPSL_SYM_UNROLL), plus the deviation of each variant's forces from the exact variant's.
    python tools/pair_variants.py [workload] [variants e.g. 0:1,2:1,0:2,2:2,1:1]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
# math:unroll[:rows:threads(+1 = tighter register cap)]
variants = [tuple(int(x) for x in v.split(":")) for v in (sys.argv[2] if len(sys.argv) > 2 else "0:1,2:1,0:2,2:2,1:1").split(",")]
for kv in sys.argv[3:]:  # further NPSL_* knobs as NAME=value
    os.environ[kv.split("=")[0]] = kv.split("=")[1]
w = W.build(name)
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
rng = np.random.default_rng(3)
pos = w.mesh.pos + 0.001 * rng.standard_normal(w.mesh.pos.shape)
base = None
for var in variants:
    math, unroll = var[:2]
    os.environ["NPSL_SYM_MATH"], os.environ["NPSL_SYM_UNROLL"] = str(math), str(unroll)
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
        ctx.upload_state(pos, np.zeros_like(pos), w.q)
        if len(var) > 2:
            ctx.set_pair_plan(var[2], var[3])
        ctx.force_init()
        pf = ctx.download_force_components()["pair"]
        ctx.step(5)
        ctx.pair_timing(True)
        ctx.step(20)
        ms, cnt = ctx.pair_timing_read()
        ctx.pair_timing(False)
        per = ctx.step_timing_list(20, flush_l2=True, lead_in=1)
        if base is None:
            base = pf
        nb = np.linalg.norm(base, axis=1)
        dev = np.linalg.norm(pf - base, axis=1)
        print("%s math %d unroll %d (in effect: math %d): pair kernel %.4f ms, step %.4f ms; vs first variant: max|d|/max|F| %.2e, "
              "per vertex %.2e" % (":".join(str(x) for x in var), math, unroll, ctx.pair_math(), ms, float(per.mean()), dev.max() / nb.max(),
                                   float(np.max(dev / np.maximum(nb, 1e-6 * nb.max())))), flush=True)
