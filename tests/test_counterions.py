"""Explicit counterions on the device: mesh-ion and ion-ion forces and energies, the ion half-kicks /
drift and the ion Nose-Hoover chain, against the compiled reference run with 64 hand-placed monovalent ions
(tests/golden/disc_D4_ions.npz; oracle/ref_driver.cpp:place_ions -- the reference's own put_counterions cannot terminate
with its hard-wired box). Some ions sit inside the mesh-ion WCA cutoff and two overlap, so the LJ branches are live."""
import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def ion_ctx(g, **kw):
    from test_gpu_parity import make_ctx
    ctx = make_ctx(g, **kw)
    ctx.set_ions(g["ion_pos0"], None, g["ion_q"], float(g["ion_params"][0]), float(g["ion_params"][1]))
    return ctx


def check_ions(ctx, g, k):
    ions = ctx.download_ions()
    assert relerr(ions["force"], g["k%d_ion_force" % k]) < TOL, (k, "ion force")
    assert relerr(ions["pos"], g["k%d_ion_pos" % k]) < 1e-13, (k, "ion pos")
    assert relerr(ions["vel"], g["k%d_ion_vel" % k], floor=1e-9) < TOL, (k, "ion vel")
    e = ctx.energy()
    ref = float(g["k%d_netIonKE" % k])
    assert abs(e["ions_kinetic"] - ref) <= TOL * max(ref, 1e-9), (k, e["ions_kinetic"], ref)


def test_counterion_forces_energies_and_thermostat_match_reference():
    from test_gpu_parity import check_against_fixture
    g = golden("disc_D4_ions")
    with ion_ctx(g) as ctx:
        ctx.force_init()
        done = 0
        for k in g.full_steps:
            ctx.step(k - done)
            done = k
            check_against_fixture(ctx, g, k)  # mesh side: forces incl. the ions' pull, LJ / ES energies incl. the ion terms,
            check_ions(ctx, g, k)             # E = KE + ion KE + U, both thermostat chains; then the ions themselves


def test_counterion_trajectory_matches_reference():
    g = golden("disc_D4_ions")
    with ion_ctx(g) as ctx:
        ctx.force_init()
        done = 0
        for k in g.traj_steps:
            ctx.step(k - done)
            done = k
            e, s = ctx.energy(), g.scalars(k)
            assert abs(e["total_area"] - s["total_area"]) <= 1e-6 * s["total_area"]
            assert abs(e["potential"] - s["penergy"]) <= 1e-6 * abs(s["penergy"])
            assert abs(e["energy"] - s["energy"]) <= 1e-6 * abs(s["energy"])
            ref = float(g["k%d_netIonKE" % k])
            assert abs(e["ions_kinetic"] - ref) <= 1e-6 * ref
        assert relerr(ctx.download_ions()["pos"], g["k%d_ion_pos" % g.traj_steps[-1]]) < 1e-6


def test_counterions_with_the_symmetric_kernel_and_removal():
    """Above 4096 vertices the mesh-mesh forces come from k_pair_sym; the ion terms must not depend on that. Checked on
    the synthetic S16 mesh against the ordered kernel, then the ions are removed again."""
    from np_shape_lab_b200 import capi, workloads as W
    w = W.build("S32_disc")
    P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
    rng = np.random.default_rng(3)
    d = rng.standard_normal((200, 3))
    ipos = d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(1.02, 1.6, size=(200, 1))
    iq = -np.ones(200)
    res = []
    for mode in (2, 1):
        with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
            ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
            ctx.set_pair_mode(mode)
            ctx.set_ions(ipos, None, iq, 0.06, 0.05)
            ctx.force_init()
            ctx.step(5)
            res.append((ctx.download_state(), ctx.download_ions(), ctx.energy()))
            if mode == 2:
                ctx.set_ions(None, None, None, 0.06, 0.05)  # n = 0: back to a mesh-only context
                ctx.force_init()
                e0 = ctx.energy()
                assert e0["ions_kinetic"] == 0.0
    (s2, i2, e2), (s1, i1, e1) = res
    assert relerr(s2["force"], s1["force"]) < 1e-11 and relerr(i2["force"], i1["force"]) < 1e-11
    assert relerr(i2["pos"], i1["pos"]) < 1e-12
    for k in ("electrostatic", "lj", "ions_kinetic", "energy"):
        assert abs(e2[k] - e1[k]) <= 1e-11 * max(abs(e1[k]), 1e-9), k


def test_python_host_mirror_with_counterions():
    """md_interface / force_calculation of the Python mirror with a list of PARTICLEs."""
    from np_shape_lab_b200 import host as H, mesh as M
    g = golden("disc_D4_ions")
    inp = M.PhysicalInputs(q=600.0, c=0.005)
    b = H.INTERFACE.from_inputs(M.Mesh(g["pos0"], g["edges"], g["faces"]), inp)
    d_ion, b.lj_length_mesh_ions = float(g["ion_params"][0]), float(g["ion_params"][1])
    ions = [H.PARTICLE(k + 1, d_ion, 1, float(q), 1.0, p) for k, (p, q) in enumerate(zip(g["ion_pos0"], g["ion_q"]))]
    mesh_bath, ions_bath = H.make_baths(inp, g.nv(), len(ions))
    assert ions_bath[0].dof == 3 * len(ions)
    sf = g.params["scalefactor"]
    H.force_calculation_init(b, ions, sf, "n")
    assert relerr(b.forvec, g["k0_force"]) < TOL
    assert relerr(np.array([p.forvec for p in ions]), g["k0_ion_force"]) < TOL
    c = H.CONTROL()
    c.steps, c.writedata = 100, 100
    H.md_interface(b, ions, mesh_bath, ions_bath, c, "N", "n", "L", sf)
    s = g.scalars(100)
    assert abs(b.penergy - s["penergy"]) <= 1e-6 * abs(s["penergy"])
    assert abs(b.energy - s["energy"]) <= 1e-6 * abs(s["energy"])
    assert abs(b.netIonKE - float(g["k100_netIonKE"])) <= 1e-6 * float(g["k100_netIonKE"])
