#!/usr/bin/env python
"""Benchmark of the B200-native np-shape-lab force/energy/integrator path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload S64_disc] [--impl b200|reference]

A "step" is one iteration of the md_interface loop body (src/newmd.cpp:96-170): reverse
Nose-Hoover half-chain, half-kick, drift, dressup, force_calculation (all-pairs screened Coulomb +
bending / stretching / surface-tension / volume-tension), half-kick, kinetic energy, forward
half-chain. Metric: MD steps/s (and ordered pair interactions/s = N(N-1) steps/s, counted as the
reference evaluates them) on BASELINE.json's synthetic D=64 charged-disc mesh (40,962 vertices).

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for how each field is obtained.
For N > 1 launch with torchrun (one rank per GPU); rows of the pair interaction and the owned
vertex range are sharded, total work is fixed ("strong" scaling).

--impl reference times the reference's own CPU implementation (oracle/_ref/npsl_ref = its unmodified
sources compiled against shim headers) on the host cores; see run_reference().
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# stdout carries exactly one JSON line. NCCL writes its "NCCL version ..." banner (and any NCCL_DEBUG output) to the C
# stdout of rank 0, so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate
# of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "md_steps_per_s"
UNIT = "steps/s"
# FP64-pipe instructions the pair kernel executes per pair come from the library (npsl_pair_kernel_info: SASS counts of
# the inner loops, tools/sass_loops.py -> profiles/r02_pair_sass_loops.txt): k_pair_sym in its default form (reduced-
# accuracy exact scheme, 6 rows per thread) needs 171 per 6 unordered pairs = 28.5, i.e. 14.25 per ORDERED pair as the
# reference counts pairs; the ordered k_pair needs 27.
ALGORITHMIC_FLOPS_PER_PAIR = 24  # force-only, exp and rsqrt counted as one flop each


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="npsl_clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self, wait_s: float = 8.0):
        """Starts nvidia-smi in loop mode and returns once it has written its first sample (or after wait_s): its start-up
        (NVML initialisation touches every GPU of the box) must not fall into the timed region of a sharded job, where a
        hiccup on any GPU lands in every rank's brackets."""
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < wait_s:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.05)

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation (unmodified sources, oracle/_ref) on the host cores
def reference_sample(workload_name: str, budget_s: float, threads: int, reps: int = 1, warm: int = 0):
    """Times force_calculation of the reference (src/newforces.cpp:137-289) on a bounded row sample of
    the workload -- its own MPI row slice [lowerBoundMesh, upperBoundMesh] (src/main.cpp:448-455) --
    and extrapolates to a full step:  t_step = t_fixed + (t_sample - t_fixed) * N / rows, where t_fixed
    (mesh terms + gather, evaluated in full on every rank) is measured with a 1-row slice.
    Returns dict(value=steps/s, pairs_per_s, sample=..., per-call times)."""
    from np_shape_lab_b200 import mesh as M
    from np_shape_lab_b200 import workloads as W
    from oracle import refdump as R
    if not R.have_ref():
        raise RuntimeError("oracle/_ref/npsl_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    w = W.build(workload_name)
    n = w.n_vertices
    wd = tempfile.mkdtemp(prefix="npsl_refbench_")
    os.makedirs(os.path.join(wd, "infiles"))
    os.makedirs(os.path.join(wd, "outfiles"))
    M.write_dat(os.path.join(wd, "infiles", "3_%d.dat" % w.D), w.mesh)
    flags = W.reference_flags(w)

    def call(rows, reps_, warm_):
        out = R.run_ref(wd, ["--mode", "force"] + flags + ["--row-lo", 0, "--row-hi", rows - 1, "--reps", reps_,
                                                            "--warm", warm_], threads=threads, timeout=3600)
        return R.last_json_line(out)

    fixed = call(1, 2, 1)
    t_fixed = fixed["best_s"]
    probe_rows = min(n, 64 * max(1, threads))
    probe = call(probe_rows, 1, 0)
    per_row = max((probe["best_s"] - t_fixed) / probe_rows, 1e-9)
    # every call of the reference's force_calculation evaluates the mesh terms in full (t_fixed, ~0.24 s at S64) whatever
    # the row slice is, so a large --steps cannot fit the time budget: the number of timed calls is capped and reported
    asked = reps
    max_calls = max(1, int(budget_s / (t_fixed + per_row * probe_rows)))
    reps = max(1, min(reps, max_calls))
    warm = max(0, min(warm, max_calls - reps))
    per_call_budget = budget_s / max(1, reps + warm)
    rows = int(max(probe_rows, min(n, (per_call_budget - t_fixed) / per_row)))
    res = call(rows, reps, warm)
    t_sample = res["mean_s"]
    t_step = t_fixed + (t_sample - t_fixed) * (n / rows)
    import shutil
    shutil.rmtree(wd, ignore_errors=True)
    return {
        "value": 1.0 / t_step, "pairs_per_s": n * (n - 1) / t_step, "t_step_s": t_step, "t_sample_s": t_sample,
        "t_fixed_s": t_fixed, "rows": rows, "reps": reps, "asked_reps": asked, "threads": res["threads"], "n": n,
        "config": common_config(w),
        "sample": "reference force_calculation (unmodified src/newforces.cpp via oracle/_ref) on rows [0,%d) of %d "
                  "x all %d columns, %d timed call(s), mesh terms in full; extrapolated to a full step as "
                  "t_fixed + (t_sample - t_fixed) * N/rows with t_fixed = %.3f s from a 1-row slice; 1 process x %d "
                  "OpenMP threads (no MPI launcher in the image)%s" % (
                      rows, n, n, reps, t_fixed, res["threads"],
                      "" if reps == asked else "; %d steps were asked for, the %.0f s budget fits %d calls" % (asked, budget_s, reps)),
    }


def common_config(w) -> dict:
    """The `config` object: identical on the b200 arm and on the reference arm (the driver compares them)."""
    n = w.n_vertices
    return {"workload": w.name, "vertices": n, "edges": w.mesh.n_edges, "faces": w.mesh.n_faces,
            "ordered_pairs_per_step": n * (n - 1), "dt": w.inp.dt, "reference_cli": w.cli}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args) -> None:
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    threads = host_threads()
    total = args.steps + args.warmup
    # bounded: the whole run ends within a few minutes whatever K and W are
    budget = min(180.0, max(20.0, 2.0 * total))
    r = reference_sample(args.workload, budget, threads, reps=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["t_step_s"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (x87 80-bit vectors in the reference)",
        "data": "synthetic", "pair_interactions_per_s": r["pairs_per_s"],
        "config": r["config"],
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                         "sample": r["sample"]},
        "timed_calls": r["reps"],
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (one process per GPU)" % (args.gpus, args.gpus))
        args.gpus = world
    import numpy as np
    import torch
    import torch.distributed as dist
    from np_shape_lab_b200 import capi
    from np_shape_lab_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = W.build(args.workload)
    n = w.n_vertices
    P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
    nccl_id = None
    if world > 1:
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    ctx = capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n, rank=rank, world=world,
                       nccl_id=nccl_id)
    pos0 = np.ascontiguousarray(w.mesh.pos)
    vel0 = np.zeros_like(pos0)
    ctx.upload_state(pos0, vel0, w.q)
    ctx.force_init()
    e0 = ctx.energy()
    launches0 = ctx.launch_count()

    # ---- device-resident steps: W warm-up, then K timed (one CUDA-graph launch per step)
    warm = max(args.warmup, 3)
    ctx.step(warm)
    # a short untimed pass through the timing entry point: allocates and touches the L2-flush buffer and the events,
    # so that nothing is allocated inside the timed call (in a sharded job a rank's allocation would otherwise land in
    # its peers' brackets: every step waits for every rank's exchange signal)
    ctx.step_timing_list(min(args.steps, 3), flush_l2=True, lead_in=0)
    ctx.sync()
    extra_untimed = min(args.steps, 3) + 1
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # returns after the first sample: nvidia-smi's start-up stays outside the timed region
    # The timed region: one untimed lead-in step absorbs the skew the host-side barrier leaves between the ranks, then
    # EXACTLY K timed steps, each in its own CUDA-event bracket on the context's stream, the L2 flush between the
    # brackets. A step of a sharded job waits for every rank, so one stall on any GPU (another process's NVML call, a
    # clock transition) lands in every rank's bracket: a region with such an outlier (a step > 3x the median) is
    # measured again, at most twice; every attempt is reported in `attempts`, none is hidden.
    attempts = []
    for attempt in range(3):
        barrier()
        l_before = ctx.launch_count()
        t0 = time.perf_counter()
        per_step = ctx.step_timing_list(args.steps, flush_l2=True, lead_in=1)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        gpu_launches = ctx.launch_count() - l_before
        dev_ms = max_over_ranks(float(per_step.sum()))
        wall_ms = max_over_ranks(wall_ms)
        if world > 1:  # per step: the slowest rank
            tt = torch.tensor(per_step, dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            per_step = tt.cpu().numpy()
        med, worst = float(np.median(per_step)), float(np.max(per_step))
        attempts.append({"ms_per_step": dev_ms / args.steps, "median_ms": med, "max_ms": worst})
        if worst <= 3.0 * med:  # same decision on every rank: per_step is the all-reduced list
            break
    clocks = sampler.stop() if rank == 0 else {}
    ms_per_step = dev_ms / args.steps
    steps_per_s = 1e3 / ms_per_step
    e1 = ctx.energy()

    # ---- the pair kernel alone (CUDA events on the context's stream around each launch)
    n_pair = max(5, min(args.steps, 50))
    ctx.pair_timing(True)
    ctx.step(n_pair)
    pair_ms, pair_count = ctx.pair_timing_read()
    ctx.pair_timing(False)
    pair_ms = max_over_ranks(pair_ms)
    geo = ctx.pair_geometry()
    lo, hi = ctx.owned_rows()

    profile_txt = ctx.step_profile(20)  # per-segment device time of directly launched steps (CUDA events between launches)

    # ---- end to end through the host-state entry point: the state lives in host memory between steps (the reference
    # keeps VERTEX::posvec / velvec / forvec there); every step copies the page-locked block in and back out
    st = ctx.download_state()
    hp, hv, hf, hk = ctx.host_state()
    hp[:], hv[:], hf[:] = st["pos"], st["vel"], st["force"]
    for _ in range(3):
        ctx.md_step_packed(1)
    n_e2e = max(5, min(args.steps, 200))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        ctx.md_step_packed(1)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    ke = float(hk[0])
    e2e_steps_per_s = n_e2e / e2e_s
    bytes_dir = 3 * 3 * 8 * n  # pos, vel, force [n][3] doubles each way (+ 8 B of KE down)

    # ---- roofline denominators
    dfma_peak, nominal_mhz = capi.fp64_peak()
    dfma_peak = max_over_ranks(dfma_peak) if world > 1 else dfma_peak
    pairs_per_launch = (hi - lo) * (n - 1)
    mode = ctx.pair_mode()
    kinfo = ctx.pair_kernel_info()
    pair_math = ctx.pair_math()
    dfma_per_ordered_pair = kinfo["fp64_instr_per_pair"] / (2.0 if mode == 2 else 1.0)
    ctx_pair_mode, ctx_xchg_mode, ctx_shard_plan = mode, ctx.exchange_mode(), ctx.shard_plan()
    # the symmetric kernel's launch covers this rank's share (1/world) of all unordered pairs
    pairs_per_launch = n * (n - 1) // world if mode == 2 else pairs_per_launch
    achieved_dfma = pairs_per_launch * dfma_per_ordered_pair / (pair_ms * 1e-3) if pair_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    traffic = None
    try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this workload, if any
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("%s:%d" % (args.workload, world), {}).get("bytes")
    except Exception:
        pass

    # the O(N) kernels on the critical path: algorithmic bytes / segment time against the measured HBM peak
    seg = {}
    for ln in profile_txt.splitlines():
        t = ln.split()
        if len(t) >= 3 and t[-1] == "us":
            seg[" ".join(t[:-2])] = float(t[-2])
    n_own = hi - lo
    hbm_kernels = []
    for name, nbytes, what in (
            ("k_kick_drift", 160 * n_own, "x, v, f read, x, v written: 160 B per integrated vertex"),
            ("k_vertex", (8 * 24 + 32 + 32 + 8 + 64) * n_own, "8 pair sums, fmesh, v, q read, f, v written: 328 B per integrated vertex")):
        if name in seg and seg[name] > 0:
            hbm_kernels.append({"kernel": name, "us": seg[name], "algorithmic_bytes": nbytes, "gbs": nbytes / seg[name] * 1e-3,
                                "bytes": what})

    ctx.close()

    # ---- secondary block: BASELINE.json configs[4] (S128 rod, 163 842 vertices, "strong scaling at 1/2/4/8 GPUs"), a few
    # steps with the same timing discipline, so that every --gpus N line carries a driver-measured point of that curve
    secondary = None
    if args.secondary and args.secondary != args.workload:
        w2 = W.build(args.secondary)
        n2 = w2.n_vertices
        P2 = capi.make_params(w2.params, w2.mesh_bath, w2.ions_bath)
        id2 = None
        if world > 1:
            box = [capi.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            id2 = box[0]
        with capi.Context(w2.mesh.edges, w2.mesh.faces, w2.params["l0"], P2, n_vertices=n2, rank=rank, world=world,
                          nccl_id=id2) as c2:
            c2.upload_state(np.ascontiguousarray(w2.mesh.pos), np.zeros_like(w2.mesh.pos), w2.q)
            c2.force_init()
            c2.step(3)
            c2.step_timing_list(2, flush_l2=True, lead_in=0)
            c2.sync()
            barrier()
            k2 = max(3, min(args.steps, 10))
            per2 = c2.step_timing_list(k2, flush_l2=True, lead_in=1)
            barrier()
            ms2 = max_over_ranks(float(per2.sum())) / k2
            c2.pair_timing(True)
            c2.step(3)
            pair2, _ = c2.pair_timing_read()
            c2.pair_timing(False)
            pair2 = max_over_ranks(pair2)
            en2 = c2.energy()
            plan2 = c2.shard_plan()
        secondary = {"metric": METRIC, "value": 1e3 / ms2, "unit": UNIT, "n_gpus": world, "steps": k2, "ms_per_step": ms2,
                     "pair_interactions_per_s": 1e3 / ms2 * n2 * (n2 - 1), "pair_kernel_ms": pair2,
                     "config": common_config(w2), "shard_plan": {0: "single GPU", 1: "rows", 2: "replicated integrator"}[plan2],
                     "finite": bool(np.isfinite(en2["energy"])), "scaling": "strong"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cons = lambda e: e["energy"] + e["mesh_bath_ke"] + e["mesh_bath_pe"] + e["ions_bath_ke"] + e["ions_bath_pe"]
    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "pair_interactions_per_s": steps_per_s * n * (n - 1),
        "config": common_config(w),
        "setup": {
            "sharding": {0: "single GPU",
                         1: "rows: units of the pair term (8 virtual ranks) and owned vertices split over %d GPUs; positions, "
                            "pair-force sums and kinetic-energy partials exchanged each step" % world,
                         2: "replicated integrator: units of the pair term (8 virtual ranks) split over %d GPUs, every GPU "
                            "integrates all vertices; one exchange (pair-force sums) per step" % world}[ctx_shard_plan],
            "l2": "flushed between timed steps (256 MiB memset outside each step's CUDA-event bracket); "
                  "ms_per_step = sum of the K per-step brackets / K, max over ranks; one untimed lead-in step after the "
                  "barrier; %d further untimed steps (timing path warm-up) beside the W warm-up steps" % extra_untimed,
            "pair_plan": geo, "pair_kernel": {1: "ordered (k_pair)", 2: "symmetric (k_pair_sym)"}[ctx_pair_mode],
            "exchange": {0: "none (single GPU)", 1: "NCCL all-gathers", 2: "direct stores into peer memory over NVLink "
                         "(CUDA IPC) with epoch flags"}[ctx_xchg_mode],
        },
        "wall_ms_per_step_incl_flush": wall_ms / args.steps,
        "per_step_ms": {"median": float(np.median(per_step)), "max": float(np.max(per_step)), "min": float(np.min(per_step)),
                        "list": [round(float(x), 5) for x in per_step[:200]],
                        "note": "each step's bracket, max over ranks; value uses the plain sum of all K"},
        "attempts": attempts,
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "e2e": {"value": e2e_steps_per_s, "unit": UNIT, "h2d_bytes_per_step": bytes_dir,
                "d2h_bytes_per_step": bytes_dir + 8, "steps": n_e2e,
                "ms_per_step": 1e3 * e2e_s / n_e2e,
                "call": "npsl_md_step_packed(ctx, 1) per step on the page-locked block of npsl_host_state: posvec | velvec | "
                        "forvec host->device in one copy, one step, positions back beside the force evaluation, velvec | "
                        "forvec | KE back in one copy at the end; wall clock over the calls, max over ranks"},
        "roofline": {
            "bound": "fp64", "achieved": 2e-12 * achieved_dfma, "peak": 2e-12 * dfma_peak, "unit": "TFLOP/s",
            "frac": achieved_dfma / dfma_peak if dfma_peak else None, "traffic": traffic,
            "kernel": "k_pair_sym<%d rows x %d threads, %s> (each unordered pair once, applied to both vertices)" % (
                          kinfo["rows_per_thread"], kinfo["threads"],
                          {0: "rsqrt / exp to 3e-16", 1: "piecewise-polynomial radial factor", 2: "reduced-accuracy rsqrt / exp, < 1e-12"}[pair_math])
                      if mode == 2 else "k_pair<%d,false>" % geo["rows_per_thread"],
            "kernel_ms": pair_ms, "kernel_launches_timed": pair_count,
            "kernel_share_of_step": pair_ms / ms_per_step,
            "fp64_instr_per_ordered_pair": dfma_per_ordered_pair, "ordered_pairs_per_launch": pairs_per_launch,
            "algorithmic_tflops_%dflop_per_pair" % ALGORITHMIC_FLOPS_PER_PAIR:
                1e-12 * pairs_per_launch * ALGORITHMIC_FLOPS_PER_PAIR / (pair_ms * 1e-3) if pair_ms > 0 else None,
            # the same against the peak: the fraction on SURVEY 8d's algorithmic count (exp and rsqrt as one flop each)
            "algorithmic_frac": (pairs_per_launch * ALGORITHMIC_FLOPS_PER_PAIR / (pair_ms * 1e-3)) / (2.0 * dfma_peak)
                                if pair_ms > 0 and dfma_peak else None,
            "peak_source": "pure-DFMA probe kernel run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                           "nominal 148 SM x 64 DFMA/clk x %.0f MHz = %.2f TFLOP/s" %
                           (nominal_mhz, 148 * 64 * 2 * nominal_mhz * 1e-6),
            "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
        },
        "hbm_kernels": {"peak_gbs": peaks.get("hbm_gbs"), "kernels": hbm_kernels, "segments_us": seg,
                        "note": "segments of directly launched steps, event to event, so each includes its launch gap; these "
                                "kernels are latency-bound at this size (DESIGN.md section 3 has the ncu durations and the "
                                "bandwidth-bound k_pair_reduce)"},
        "checks": {"conserved_energy_drift": cons(e1) / cons(e0) - 1.0, "finite": bool(np.isfinite(cons(e1))),
                   "kinetic_after_e2e": ke,
                   # the device-timed step can only be faster than the same step with host copies around it
                   "timing_inconsistent": bool(ms_per_step > 1.02e3 * e2e_s / n_e2e)},
    }
    if secondary is not None:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_sample(args.workload, args.cpu_budget, host_threads(), reps=1, warm=0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                    "sample": r["sample"], "pair_interactions_per_s": r["pairs_per_s"]}
        except Exception as ex:  # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference",
                                    "sample": "unavailable: %s" % ex}
    emit(line)
    if profile_txt and args.profile:
        sys.stderr.write("step breakdown on rank 0 (%d GPU(s)):\n%s" % (world, profile_txt))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="S64_disc")
    ap.add_argument("--secondary", default="S128_rod", help="second workload measured with a few steps ('' = none)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="also print a per-segment device-time breakdown of a step to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
