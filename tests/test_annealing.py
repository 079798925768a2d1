"""Gradual annealing of the mesh thermostat (src/newmd.cpp:258-299): from step annealfreq on, for annealDuration steps,
T and Q of every link of the mesh chain are multiplied by fT = (1 / TAnnealFac)^(1 / annealDuration) and
fQ = (1 / QAnnealFac)^(1 / annealDuration) after each step; TAnnealFac is 1.25 in the first window, 2 in the next two, ...
The fixture comes from the compiled reference run with CONTROL::annealfreq = 400 and annealDuration = 100 (plain CONTROL
fields, src/control.h; the CLI defaults are 50 000 / 5 000, src/main.cpp:165-168): two windows inside 1000 steps. Both host
mirrors (the C++ driver and host.py) must follow it: T and Q of every link exactly as the reference's arithmetic gives
them, xi / eta and A / U / E to trajectory accuracy."""
import json
import subprocess

import numpy as np
import pytest

from conftest import golden
from np_shape_lab_b200 import mesh as M

FREQ, DURA = 400, 100


def expected_T_Q(k, T0=0.001, Q0=0.01):
    """The reference's schedule restated: which steps anneal and by how much (src/newmd.cpp:258-299)."""
    T, Q = T0, Q0
    fT = fQ = None
    for num in range(1, k + 1):
        if num > DURA and (num % FREQ == 0 or num % FREQ < DURA):
            if num % FREQ == 0:
                fac = 1.25 if num == FREQ else (2.0 if num <= 3 * FREQ else 4.0)
                fT = (1.0 / fac) ** (1.0 / DURA)
                fQ = (1.0 / (1.0 + fac / 10.0)) ** (1.0 / DURA)
            T, Q = fT * T, fQ * Q
    return T, Q


def test_fixture_follows_the_schedule():
    """CPU: the reference's own T and Q at the dumped steps equal the restated schedule (the dummy tail link keeps Q = 0)."""
    g = golden("disc_D4_anneal")
    for k in g.traj_steps:
        mb = g["k%d_mesh_bath" % k]
        T, Q = expected_T_Q(k)
        assert np.allclose(mb[:, 1], T, rtol=1e-13, atol=0), (k, mb[:, 1], T)
        assert np.allclose(mb[:-1, 0], Q, rtol=1e-13, atol=0) and mb[-1, 0] == 0.0, (k, mb[:, 0], Q)
    assert g["k399_mesh_bath"][0, 1] == 0.001 and g["k400_mesh_bath"][0, 1] < 0.001  # the first window opens at step 400
    assert abs(g["k500_mesh_bath"][0, 1] / (0.001 / 1.25) - 1) < 1e-12               # ... and divides T by 1.25 in total
    assert abs(g["k900_mesh_bath"][0, 1] / (0.001 / 1.25 / 2.0) - 1) < 1e-12         # the second window by 2
    assert np.array_equal(g["k500_mesh_bath"][:, :2], g["k799_mesh_bath"][:, :2])    # nothing in between


def check(g, k, A, U, ke_plus_pe, mesh_bath):
    """mesh_bath: rows (Q, T, dof, xi, eta) after the run. The fixture's bath energies were evaluated after md_interface
    returned, i.e. with T and Q as the annealing of the last step left them (oracle/ref_driver.cpp), so the conserved
    energy is assembled the same way here: KE + PE of the last write step + the bath terms from the final links."""
    s = g.scalars(k)
    got = np.asarray(mesh_bath)
    bath_ke = float(np.sum(0.5 * got[:, 0] * got[:, 3] ** 2))         # src/thermostat.h:61-70 (kB = 1)
    bath_pe = float(np.sum(got[:, 2] * got[:, 1] * got[:, 4]))
    cons = s["energy"] + s["mesh_bath_ke"] + s["mesh_bath_pe"] + s["ions_bath_ke"] + s["ions_bath_pe"]
    E = ke_plus_pe + bath_ke + bath_pe + s["ions_bath_ke"] + s["ions_bath_pe"]
    assert abs(A - s["total_area"]) <= 1e-6 * s["total_area"], (k, "A")
    assert abs(U - s["penergy"]) <= 1e-6 * abs(s["penergy"]), (k, "U")
    assert abs(E - cons) <= 1e-6 * abs(cons), (k, "E", E, cons)
    assert abs(bath_ke - s["mesh_bath_ke"]) <= 1e-6 * abs(s["mesh_bath_ke"]), (k, "bath KE")
    assert abs(bath_pe - s["mesh_bath_pe"]) <= 1e-6 * abs(s["mesh_bath_pe"]), (k, "bath PE")
    ref = g["k%d_mesh_bath" % k]  # Q, T, dof, xi, eta per link
    assert np.allclose(got[:, 0], ref[:, 0], rtol=1e-12, atol=0), (k, "Q")
    assert np.allclose(got[:, 1], ref[:, 1], rtol=1e-12, atol=0), (k, "T")
    assert np.allclose(got[:, 3], ref[:, 3], rtol=1e-6, atol=1e-9 * np.abs(ref[:, 3]).max()), (k, "xi")
    assert np.allclose(got[:, 4], ref[:, 4], rtol=1e-6, atol=1e-9 * np.abs(ref[:, 4]).max()), (k, "eta")


@pytest.mark.gpu
@pytest.mark.parametrize("k", [450, 1000])
def test_host_cpp_driver_anneals_like_the_reference(k, tmp_path):
    from test_host_cpp import BIN, build_host
    build_host()
    g = golden("disc_D4_anneal")
    M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
    r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", str(k), "-W", str(k), "--out", str(tmp_path / "o"),
                        "-q", "600", "-c", "0.005", "-f", str(FREQ), "-u", str(DURA)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    check(g, k, res["A"], res["U"], res["KE_plus_PE"], res["mesh_bath"])


@pytest.mark.gpu
def test_python_host_mirror_anneals_like_the_reference():
    from np_shape_lab_b200 import host as H
    g = golden("disc_D4_anneal")
    inp = M.PhysicalInputs(q=600.0, c=0.005)
    b = H.INTERFACE.from_inputs(M.Mesh(g["pos0"], g["edges"], g["faces"]), inp)
    mesh_bath, ions_bath = H.make_baths(inp, g.nv(), 0)
    c = H.CONTROL()
    c.steps, c.writedata, c.annealfreq, c.annealDuration = 850, 850, FREQ, DURA
    H.md_interface(b, [], mesh_bath, ions_bath, c, "N", "n", "L", g.params["scalefactor"])
    check(g, 850, b.total_area, b.penergy, b.energy, [[t.Q, t.T, t.dof, t.xi, t.eta] for t in mesh_bath])
