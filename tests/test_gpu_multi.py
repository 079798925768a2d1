"""Multi-GPU paths under pytest -m gpu: every case spawns one process per GPU (torchrun, world 2) running
tools/check_sharded.py, which asserts that the sharded trajectory -- positions, velocities, forces, every force
component, both thermostat chains, energies -- is bit-identical to the single-GPU one. Skipped on boxes with one GPU.
The exchange over NVLink peer memory (csrc/xchg.cuh), the NCCL fallback, both sharding plans, the volume constraint,
explicit counterions, the WCA branch (a collision that only one rank's units see) and the host-state entry points."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_sharded(world, *args, timeout=600):
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + (os.getpid() + hash(args)) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py")] + [str(a) for a in args]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0 and "SHARDED_CHECK PASS" in out.stdout, out.stdout[-4000:]


CASES = {
    "replicated": ["--plan", "replicated"],
    "rows": ["--plan", "rows"],
    "rows_nccl": ["--plan", "rows", "--exchange", "nccl"],
    "replicated_gv": ["--plan", "replicated", "--gv"],
    "rows_gv": ["--plan", "rows", "--gv"],
    "replicated_ions": ["--plan", "replicated", "--ions", "100"],
    "rows_ions": ["--plan", "rows", "--ions", "100"],
    "replicated_collision": ["--plan", "replicated", "--collide", "--steps", "3"],
    "rows_collision": ["--plan", "rows", "--collide", "--steps", "3"],
    "rows_nccl_collision": ["--plan", "rows", "--exchange", "nccl", "--collide", "--steps", "3"],
    "replicated_host_state": ["--plan", "replicated", "--packed", "--steps", "10"],
    "rows_host_state": ["--plan", "rows", "--packed", "--steps", "10"],
}


@pytest.mark.parametrize("case", list(CASES))
def test_two_gpus_bit_identical_to_one(case):
    run_sharded(2, "--workload", "S32_disc", *CASES[case])


def test_host_state_entry_points_equal_resident_steps():
    """npsl_md_step_packed (1 step x k, k steps at once) and npsl_md_step_host against npsl_step on the resident state:
    bit for bit, and against the reference fixture (one GPU; the same comparison across two GPUs is above)."""
    import numpy as np
    from conftest import golden, relerr
    from test_gpu_parity import TOL, make_ctx
    g = golden("disc_D4")
    for mode in (1, 2):  # ordered and symmetric pair kernel
        with make_ctx(g) as ctx:
            ctx.set_pair_mode(mode)
            ctx.force_init()
            st0, th0 = ctx.download_state(), ctx.get_thermostat()
            ctx.step(2)
            want = ctx.download_state()
            assert relerr(want["force"], g["k2_force"]) < TOL and relerr(want["pos"], g["k2_pos"]) < 1e-13
            ke_want = ctx.energy()["kinetic"]
            hp, hv, hf, hk = ctx.host_state()
            for how in ("1+1", "2", "arrays"):
                ctx.set_thermostat(*th0)
                ctx.upload_state(st0["pos"], st0["vel"], None)
                ctx.force_init()
                if how == "arrays":
                    a = [np.ascontiguousarray(st0[k]) for k in ("pos", "vel", "force")]
                    ke = ctx.md_step_host(a[0], a[1], a[2], 2)
                else:
                    hp[:], hv[:], hf[:] = st0["pos"], st0["vel"], st0["force"]
                    if how == "2":
                        ctx.md_step_packed(2)
                    else:
                        ctx.md_step_packed(1)
                        ctx.md_step_packed(1)
                    a, ke = [hp, hv, hf], float(hk[0])
                for k, got in zip(("pos", "vel", "force"), a):
                    assert np.array_equal(got, want[k]), (mode, how, k)
                assert ke == ke_want, (mode, how)
