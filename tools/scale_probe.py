"""GPU-box probe (torchrun, one rank per GPU): sharding plans and unit granularities side by side.

For every configuration "plan[:chunk_tiles]" (plan = rows | replicated | auto) it creates a sharded context, checks
that the trajectory is bit-identical to a single-GPU run of the same library settings (rank 0 runs that one on its own
GPU), and reports steps/s (CUDA events, L2 flushed between steps, max over ranks) plus the per-segment breakdown.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/scale_probe.py S64_disc 300 rows replicated replicated:1
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from np_shape_lab_b200 import capi, workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "S64_disc"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
configs = sys.argv[3:] or ["rows", "replicated"]
check_steps = 40
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = W.build(name)
n = w.n_vertices
P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
rng = np.random.default_rng(5)
amp = min(1.0, 40962.0 / n) ** 0.5  # perturbation relative to the edge length
pos = w.mesh.pos + 0.002 * amp * rng.standard_normal(w.mesh.pos.shape)
vel = 0.01 * amp * rng.standard_normal(pos.shape)


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(ctx, k):
    ctx.upload_state(pos, vel, w.q)
    ctx.force_init()
    ctx.step(k)
    st = ctx.download_state()
    th = ctx.get_thermostat()
    return st, th, ctx.energy()


single = {}
results = []
for cfg in configs:
    plan, tiles = (cfg.split(":") + [""])[:2]
    if plan == "auto":
        os.environ.pop("NPSL_SHARD", None)
    else:
        os.environ["NPSL_SHARD"] = plan
    if tiles:
        os.environ["NPSL_SYM_CHUNK_TILES"] = tiles
    else:
        os.environ.pop("NPSL_SYM_CHUNK_TILES", None)
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n, rank=rank, world=world, nccl_id=box[0])
    st, th, en = run(ctx, check_steps)
    ok = True
    if rank == 0:
        if tiles not in single:
            with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n) as one:
                single[tiles] = run(one, check_steps)
                one.step(10)
                single[tiles] += (one.step_timing(steps, flush_l2=True) / steps,)
        s1, t1, e1, _ = single[tiles]
        for k in ("pos", "vel", "force"):
            ok &= bool(np.array_equal(st[k], s1[k]))
        ok &= bool(np.array_equal(th[0], t1[0]) and np.array_equal(th[1], t1[1]))
        ok &= en["kinetic"] == e1["kinetic"] and abs(en["electrostatic"] - e1["electrostatic"]) <= 1e-14 * abs(e1["electrostatic"])
    dist.barrier()
    ctx.step(20)
    ctx.sync()
    dist.barrier()
    ms = max_over_ranks(ctx.step_timing(steps, flush_l2=True)) / steps
    ctx.pair_timing(True)
    ctx.step(20)
    pair_ms, _ = ctx.pair_timing_read()
    ctx.pair_timing(False)
    pair_ms = max_over_ranks(pair_ms)
    prof = ctx.step_profile(20)
    shard_plan, xmode = ctx.shard_plan(), ctx.exchange_mode()
    ctx.close()
    if rank == 0:
        one_ms = single[tiles][3]
        r = {"config": cfg, "workload": name, "world": world, "shard_plan": shard_plan, "exchange_mode": xmode,
             "bit_identical_to_single_gpu": ok, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "pair_kernel_ms": pair_ms,
             "single_gpu_ms_per_step": one_ms, "strong_scaling_efficiency": one_ms / ms / world}
        results.append(r)
        print("PROBE " + json.dumps(r), flush=True)
        print(prof, flush=True)
    dist.barrier()
allok = all(r["bit_identical_to_single_gpu"] for r in results) if rank == 0 else True
flag = torch.tensor([1 if allok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
