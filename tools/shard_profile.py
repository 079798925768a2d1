"""GPU-box measurement on ONE GPU: the kernels of rank r of a w-way split (NPSL_EMULATE_SHARD=r/w: the context does only
that rank's share of the pair term; forces are incomplete, timing only) -- per-segment times of directly launched steps and
the graph-launched step, against the ideal share of the single-GPU step.
    python tools/shard_profile.py [workload] [worlds e.g. 1,2,4,8] [rank]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
worlds = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for kv in sys.argv[4:]:  # further NPSL_* knobs as NAME=value
    os.environ[kv.split("=")[0]] = kv.split("=")[1]
w = W.build(name)
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
base = None
for world in worlds:
    os.environ["NPSL_EMULATE_SHARD"] = "%d/%d" % (min(rank, world - 1), world)
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
        ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
        ctx.force_init()
        ctx.step(5)
        prof = ctx.step_profile(20)
        per = ctx.step_timing_list(50, flush_l2=True, lead_in=1)
        ctx.pair_timing(True)
        ctx.step(20)
        pair_ms, _ = ctx.pair_timing_read()
        ctx.pair_timing(False)
        ms = float(np.median(per))
        if base is None:
            base = (ms * world, pair_ms * world)
        print("== share 1/%d: graph step %.1f us (median of 50; ideal %.1f, efficiency %.3f), pair kernel %.1f us (ideal %.1f)"
              % (world, 1e3 * ms, 1e3 * base[0] / world, base[0] / world / ms, 1e3 * pair_ms, 1e3 * base[1] / world))
        print(prof, flush=True)
