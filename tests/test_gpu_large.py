"""Parity of the CUDA path on the configurations bench.py actually runs -- BASELINE.json configs[3] (S64 disc,
40 962 vertices) and configs[4] (S128 rod, 163 842 vertices) -- against fixtures written by the compiled reference
(oracle/gen_golden.py: whole reference steps at S64; the reference's own force_calculation through its MPI row bounds on
three windows of 512 rows at S128). The pair kernel is the automatically selected one (k_pair_sym from 4096 vertices up).
Per field max|d|/max|ref| and per vertex |d_i| / max(|ref_i|, 1e-6 max|ref|), both to 1e-10."""
import os

import numpy as np
import pytest

from conftest import GOLD, bending_term_scale, golden, relerr, vertex_relerr
from np_shape_lab_b200 import capi
from np_shape_lab_b200 import workloads as W
from test_gpu_parity import TOL, check_against_fixture, make_ctx

pytestmark = pytest.mark.gpu


def one_ulp_response(ctx, pos):
    """How far the device's own force moves when every input coordinate moves by one ulp (random direction): the
    99.9th percentile over the vertices of |F(x~) - F(x)|. From step 1 on the reference's positions are x87 80-bit
    numbers; the fixture (and any FP64 implementation) holds them rounded to double, i.e. up to half an ulp off, and
    the bending term turns that into a force difference (stiffness ~ bkappa / edge^3 ~ 1e6 at S64: 1e6 x 1.1e-16 per
    coordinate, a few 1e-10 per vertex) that no FP64 evaluation can remove. It is negligible against the largest
    force (max-norm 1e-12) but not against a vertex whose own force is 1e-3."""
    rng = np.random.default_rng(1)
    moved = np.where(rng.random(pos.shape) < 0.5, np.nextafter(pos, np.inf), np.nextafter(pos, -np.inf))
    f1 = ctx.force_calculation(np.ascontiguousarray(moved))
    f0 = ctx.force_calculation(np.ascontiguousarray(pos))  # back to the state the trajectory is in
    return float(np.percentile(np.linalg.norm(f1 - f0, axis=1), 99.9)), f0


def test_s64_disc_every_field_at_steps_0_and_1():
    g = golden("S64_disc")
    assert g.nv() == 40962
    with make_ctx(g) as ctx:
        assert ctx.pair_mode() == 2  # symmetric kernel, chosen automatically
        ctx.force_init()
        done = 0
        for k in g.full_steps:
            ctx.step(k - done)
            done = k
            allow = 0.0
            if k > 0:
                st = ctx.download_state()
                allow, f_again = one_ulp_response(ctx, st["pos"])
                assert np.array_equal(f_again, st["force"])  # the force pass reproduces the step's force bit for bit
                assert allow < 1e-11 * float(np.max(np.abs(st["force"])))  # nothing in the max-norm
            check_against_fixture(ctx, g, k, vertex_abs=allow)
            comp = ctx.download_force_components()
            assert vertex_relerr(comp["pair"], g["k%d_pforce" % k]) < TOL
            cond = bending_term_scale(g["edges"], g["faces"], g["k%d_pos" % k], g.params["bkappa"])
            assert vertex_relerr(comp["bending"], g["k%d_bforce" % k], cond=cond, abs_allow=allow) < TOL


def lcg_displacements(n, amp):
    """--perturb of oracle/ref_driver.cpp (mode window)."""
    st, mask = 0x9E3779B97F4A7C15, (1 << 64) - 1
    out = np.empty(3 * n)
    for k in range(3 * n):
        st = (st * 6364136223846793005 + 1442695040888963407) & mask
        out[k] = amp * (2.0 * ((st >> 11) * (1.0 / 9007199254740992.0)) - 1.0)
    return out.reshape(n, 3)


@pytest.mark.parametrize("name", ["S128_rod", "S64_disc_perturbed"])
def test_large_mesh_row_windows(name):
    z = np.load(os.path.join(GOLD, name + "_windows.npz"))
    w = W.build(str(z["workload"]))
    n = w.n_vertices
    # the reference's own charges and derived parameters (the host's agree to rounding, oracle/gen_golden.py checks it)
    prm = dict(w.params)
    prm.update(zip([str(k) for k in z["param_names"]], [float(v) for v in z["params"]]))
    q = z["q"]
    P = capi.make_params(prm, w.mesh_bath, w.ions_bath)
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=n) as ctx:
        assert ctx.pair_mode() == 2
        for tag, amp in (("k0", 0.0), ("pert", float(z["perturb"]))):
            pos = w.mesh.pos if amp == 0 else w.mesh.pos + lcg_displacements(n, amp)
            ctx.upload_state(pos, np.zeros_like(pos), q)
            ctx.force_init()
            comp = ctx.download_force_components()
            rows = z[tag + "_rows"]
            assert len(rows) >= 3 * 512
            pref = z[tag + "_pforce"]
            pscale = float(np.max(np.abs(pref)))
            assert relerr(comp["pair"][rows], pref) < TOL, (name, tag, "pair")
            assert vertex_relerr(comp["pair"][rows], pref) < TOL, (name, tag, "pair, per vertex")
            for ck, gk in (("bending", "bforce"), ("stretching", "sforce"), ("tension", "tforce"), ("volume_tension", "vforce")):
                ref = z[tag + "_" + gk]
                got = comp[ck] if ref.shape[0] == n else comp[ck][rows]
                # zero terms (stretching and volume tension on the initial sphere) get an absolute floor, and the
                # volume tension the cancellation floor explained in test_gpu_parity.check_against_fixture
                floor = pscale if ck == "volume_tension" else 1e-3 * pscale
                assert relerr(got, ref, floor=floor) < TOL, (name, tag, ck)
            e = ctx.energy()
            assert abs(e["total_area"] - float(z[tag + "_total_area"])) <= TOL * float(z[tag + "_total_area"])
            assert abs(e["total_volume"] - float(z[tag + "_total_volume"])) <= TOL * float(z[tag + "_total_volume"])
            # the symmetric kernel applies each pair to both vertices: the ordered-pair sum vanishes
            assert np.abs(comp["pair"].sum(axis=0)).max() < 1e-12 * np.abs(comp["pair"]).sum()
