"""GPU-box check (torchrun, one rank per GPU): the row-sharded trajectory is bit-identical to the
single-GPU one. Rank 0 also runs an unsharded context on its own GPU and compares positions,
velocities, forces, kinetic energy and thermostat state after a number of steps.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py [workload] [steps] [GV]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "S32_disc"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
constrained = len(sys.argv) > 3 and sys.argv[3] == "GV"  # rigid volume constraint (replicated plan only)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = W.build(name)
n = w.n_vertices
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
box = [capi.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
rng = np.random.default_rng(5)
pos = w.mesh.pos + 0.002 * rng.standard_normal(w.mesh.pos.shape)
vel = 0.01 * rng.standard_normal(pos.shape)

ctx = capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n, rank=rank, world=world, nccl_id=box[0])
ctx.upload_state(pos, vel, w.q)
if constrained:
    ctx.set_volume_constraint(True)
ctx.force_init()
if constrained:
    ctx.constraint_init()
e0 = ctx.energy()
ctx.step(steps)
e1 = ctx.energy()
st = ctx.download_state()
th = ctx.get_thermostat()
comp = ctx.download_force_components()
lo, hi = ctx.owned_rows()
print("rank %d owns rows [%d,%d) pair mode %d exchange mode %d" % (rank, lo, hi, ctx.pair_mode(), ctx.exchange_mode()), flush=True)
ctx.close()
ok = True
if rank == 0:
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n) as one:
        one.upload_state(pos, vel, w.q)
        if constrained:
            one.set_volume_constraint(True)
        one.force_init()
        if constrained:
            one.constraint_init()
        f0 = one.energy()
        one.step(steps)
        f1 = one.energy()
        s1 = one.download_state()
        t1 = one.get_thermostat()
        c1 = one.download_force_components()
    for k in ("pos", "vel", "force"):
        same = np.array_equal(st[k], s1[k])
        print("%-6s bit-identical: %s (max |diff| %.3e)" % (k, same, float(np.max(np.abs(st[k] - s1[k])))))
        ok &= same
    for k in comp:
        same = np.array_equal(comp[k], c1[k])
        print("%-15s bit-identical: %s" % (k, same))
        ok &= same
    same = np.array_equal(th[0], t1[0]) and np.array_equal(th[1], t1[1])
    print("thermostat bit-identical:", same)
    ok &= same
    for k in ("kinetic", "bending", "stretching", "tension", "volume", "total_area", "total_volume"):
        same = e1[k] == f1[k]
        print("energy %-12s identical: %s  %.17g %.17g" % (k, same, e1[k], f1[k]))
        ok &= same
    for k in ("electrostatic", "lj"):  # summed over ranks on the host: equal to rounding
        rel = abs(e1[k] - f1[k]) / max(abs(f1[k]), 1e-300)
        print("energy %-12s rel diff %.2e" % (k, rel))
        ok &= rel < 1e-14 or f1[k] == 0
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", "world", world, name, "steps", steps)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
