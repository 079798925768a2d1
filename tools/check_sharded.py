"""GPU-box check (torchrun, one rank per GPU): a sharded trajectory is bit-identical to the single-GPU one.
Rank 0 also runs an unsharded context on its own GPU and compares positions, velocities, forces, force components,
kinetic energy and thermostat state after a number of steps. tests/test_gpu_multi.py runs it under pytest -m gpu.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py \
        [--workload S32_disc] [--steps 50] [--plan auto|rows|replicated] [--exchange p2p|nccl] [--gv] [--ions M] \
        [--collide] [--packed]

--gv       rigid volume constraint (SHAKE / RATTLE)
--ions M   M explicit counterions (every rank is given the same ions)
--collide  two vertices are moved inside the WCA cutoff 2^(1/6) lj_length: the LJ branch must be live (lj > 0) and the
           collision, which only one rank's units see, must reach every rank
--packed   additionally advance the same state through npsl_md_step_packed / npsl_md_step_host and compare bit for bit
"""
import argparse
import os
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="S32_disc")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--plan", default="auto", choices=["auto", "rows", "replicated"])
ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
ap.add_argument("--gv", action="store_true")
ap.add_argument("--ions", type=int, default=0)
ap.add_argument("--collide", action="store_true")
ap.add_argument("--packed", action="store_true")
args = ap.parse_args()
if args.plan != "auto":
    os.environ["NPSL_SHARD"] = args.plan
if args.exchange == "nccl":
    os.environ["NPSL_EXCHANGE"] = "nccl"

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = W.build(args.workload)
n = w.n_vertices
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
box = [capi.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
rng = np.random.default_rng(5)
pos = w.mesh.pos + 0.002 * rng.standard_normal(w.mesh.pos.shape)
vel = 0.01 * rng.standard_normal(pos.shape)
if args.collide:
    a, b = [int(x) for x in w.mesh.edges[len(w.mesh.edges) // 3]]
    d = w.params["lj_length_mesh_mesh"]
    u = (pos[b] - pos[a]) / np.linalg.norm(pos[b] - pos[a])
    pos[b] = pos[a] + 1.05 * d * u
ions = None
if args.ions:
    d = rng.standard_normal((args.ions, 3))
    ions = (d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(1.02, 1.6, size=(args.ions, 1)), -np.ones(args.ions))


def run(ctx):
    ctx.upload_state(pos, vel, w.q)
    if ions is not None:
        ctx.set_ions(ions[0], None, ions[1], 0.06, 0.05)
    if args.gv:
        ctx.set_volume_constraint(True)
    ctx.force_init()
    if args.gv:
        ctx.constraint_init()
    out = {"e0": ctx.energy()}
    ctx.step(args.steps)
    out["e1"] = ctx.energy()
    out["st"] = ctx.download_state()
    out["th"] = ctx.get_thermostat()
    out["comp"] = ctx.download_force_components()
    if ions is not None:
        out["ions"] = ctx.download_ions()
    if args.packed:  # the host-state entry points continue the trajectory exactly as npsl_step does
        k = 3
        st, th = out["st"], out["th"]

        def rewind():  # back to the state after args.steps steps, thermostat included
            ctx.set_thermostat(*th)
            ctx.upload_state(st["pos"], st["vel"], None)
            ctx.force_init()

        rewind()
        ctx.step(k)
        out["resident"] = tuple(ctx.download_state()[x] for x in ("pos", "vel", "force")) + (ctx.energy()["kinetic"],)
        hp, hv, hf, hk = ctx.host_state()
        rewind()
        hp[:], hv[:], hf[:] = st["pos"], st["vel"], st["force"]
        for _ in range(k):
            ctx.md_step_packed(1)
        out["packed1"] = (hp.copy(), hv.copy(), hf.copy(), float(hk[0]))
        rewind()
        hp[:], hv[:], hf[:] = st["pos"], st["vel"], st["force"]
        ctx.md_step_packed(k)
        out["packedk"] = (hp.copy(), hv.copy(), hf.copy(), float(hk[0]))
        rewind()
        a3 = [np.ascontiguousarray(st[x]) for x in ("pos", "vel", "force")]
        ke = ctx.md_step_host(a3[0], a3[1], a3[2], k)
        out["host"] = (a3[0], a3[1], a3[2], ke)
    return out


ctx = capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n, rank=rank, world=world, nccl_id=box[0])
mine = run(ctx)
lo, hi = ctx.owned_rows()
print("rank %d owns rows [%d,%d) pair mode %d exchange mode %d shard plan %d" %
      (rank, lo, hi, ctx.pair_mode(), ctx.exchange_mode(), ctx.shard_plan()), flush=True)
plan_ok = {"auto": True, "rows": ctx.shard_plan() == 1, "replicated": ctx.shard_plan() == 2}[args.plan]
plan_ok &= ctx.exchange_mode() == (1 if args.exchange == "nccl" else 2)
ctx.close()
ok = plan_ok
if not plan_ok:
    print("rank %d: requested plan / exchange not in effect" % rank)
if rank == 0:
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n) as one:
        ref = run(one)

    def same(label, a, b):
        global ok
        eq = np.array_equal(a, b)
        print("%-22s bit-identical: %s%s" % (label, eq, "" if eq else "  (max |diff| %.3e)" % float(np.max(np.abs(np.asarray(a) - np.asarray(b))))))
        ok &= bool(eq)

    for k in ("pos", "vel", "force"):
        same(k, mine["st"][k], ref["st"][k])
    for k in mine["comp"]:
        same("component " + k, mine["comp"][k], ref["comp"][k])
    same("thermostat mesh", mine["th"][0], ref["th"][0])
    same("thermostat ions", mine["th"][1], ref["th"][1])
    if ions is not None:
        for k in ("pos", "vel", "force"):
            same("ion " + k, mine["ions"][k], ref["ions"][k])
    for k in ("kinetic", "ions_kinetic", "bending", "stretching", "tension", "volume", "total_area", "total_volume"):
        eq = mine["e1"][k] == ref["e1"][k]
        print("energy %-14s identical: %s  %.17g %.17g" % (k, eq, mine["e1"][k], ref["e1"][k]))
        ok &= eq
    for k in ("electrostatic", "lj"):  # summed over ranks on the host: equal to rounding
        rel = abs(mine["e1"][k] - ref["e1"][k]) / max(abs(ref["e1"][k]), 1e-300)
        print("energy %-14s rel diff %.2e  (%.17g)" % (k, rel, ref["e1"][k]))
        ok &= rel < 1e-13 or ref["e1"][k] == 0
    if args.collide:
        live = ref["e0"]["lj"] > 0 and mine["e0"]["lj"] > 0
        print("LJ branch live: %s (lj energy at step 0: %.6g)" % (live, ref["e0"]["lj"]))
        ok &= live
    if args.packed:
        for name in ("resident", "packed1", "packedk", "host"):
            for lab, a, b in zip(("pos", "vel", "force"), mine[name][:3], ref[name][:3]):
                same("%s %s vs 1 GPU" % (name, lab), a, b)
            eq = mine[name][3] == ref[name][3]
            print("%s kinetic energy identical to 1 GPU: %s" % (name, eq))
            ok &= eq
        for name in ("packed1", "packedk", "host"):  # the three host-state forms against npsl_step on the resident state
            for lab, a, b in zip(("pos", "vel", "force"), ref[name][:3], ref["resident"][:3]):
                same("%s vs npsl_step %s" % (name, lab), a, b)
            eq = ref[name][3] == ref["resident"][3]
            print("%s kinetic energy identical to npsl_step: %s" % (name, eq))
            ok &= eq
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", "world", world, vars(args))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
