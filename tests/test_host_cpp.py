"""The C++ host mirror (np-shape-lab_b200/host_cpp): same call surface as the reference, bodies over
the C ABI. CPU: it builds, links against libnpsl.so and fails loudly without a device. GPU: the
README disc / Janus / buckled recipes run through its command-line driver reproduce the reference's
own trajectories (tests/golden, written by the compiled reference)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from np_shape_lab_b200 import mesh as M

HOST = os.path.join(ROOT, "np-shape-lab_b200", "host_cpp")
BIN = os.path.join(HOST, "np_shape_lab_b200")


def build_host():
    from np_shape_lab_b200 import build as B
    B.build()
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_host_cpp_builds_and_has_no_cpu_path(tmp_path):
    build_host()
    out = subprocess.run([BIN, "--help"], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "no CPU path" in out.stdout
    src = open(os.path.join(HOST, "npsl_host.hpp")).read()
    for name in ("force_calculation_init", "force_calculation", "compute_energy", "dressup", "md_interface",
                 "compute_kinetic_energy", "bath_kinetic_energy", "bath_potential_energy", "class VERTEX",
                 "class EDGE", "class FACE", "class INTERFACE", "class THERMOSTAT", "class CONTROL"):
        assert name in src, name
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        g = golden("disc_D4")
        M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
        r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", "1", "--out", str(tmp_path / "o")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr


RECIPES = {
    "disc_D4": ["-q", "600", "-c", "0.005"],
    "janus_D8": ["-q", "600", "-c", "0.005", "-N", "2", "-p", "0.5"],
    "buckled_D4": ["-q", "0", "-c", "0.005", "-t", "0", "-v", "0", "-b", "1", "-s", "1000", "-B", "y"],
    "disc_D4_GV": ["-q", "600", "-c", "0.005", "-G", "V"],
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(RECIPES))
def test_host_cpp_reproduces_reference_trajectory(name, tmp_path):
    build_host()
    g = golden(name)
    M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
    k = g.traj_steps[-1]
    r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", str(k), "--out", str(tmp_path / "outfiles")]
                       + RECIPES[name], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    s = g.scalars(k)
    cons = s["energy"] + s["mesh_bath_ke"] + s["mesh_bath_pe"] + s["ions_bath_ke"] + s["ions_bath_pe"]
    assert abs(res["A"] - s["total_area"]) <= 1e-6 * s["total_area"]
    assert abs(res["U"] - s["penergy"]) <= 1e-6 * abs(s["penergy"])
    assert abs(res["E"] - cons) <= 1e-6 * abs(cons)
    # the series files the reference writes (rows at 0, 1, 2 and every writedata steps)
    rows = [l.split() for l in open(tmp_path / "outfiles" / "energy_nanomembrane.dat")]
    assert [int(x[0]) for x in rows[:3]] == [0, 1, 2] and int(rows[-1][0]) == k
    s0 = g.scalars(0)
    assert abs(float(rows[0][2]) - s0["penergy"]) <= 2e-5 * abs(s0["penergy"])  # 6 printed digits
    parts = [l.split() for l in open(tmp_path / "outfiles" / "energy_in_parts_kE_bE_sE_tE_ljE_esE.dat")]
    assert len(parts) == len(rows) and len(parts[0]) == 8
    for f in ("area.dat", "volume.dat", "temperature.dat", "net_Energy_Drift.dat"):
        assert os.path.getsize(tmp_path / "outfiles" / f) > 0


# README.md:29-65 of the reference: final area A, potential energy U and conserved energy E after
# 25 000 steps on the 372-vertex mesh, "on order of a percent".
README_FINALS = {
    "disc_D4": (["-q", "600", "-c", "0.005"], 15.675, 2620.98, 2699.97),
    "rod_D4": (["-q", "600", "-c", "0.01"], 13.259, 1919.14, 1939.96),
    "janus_D4": (["-q", "600", "-c", "0.005", "-N", "2", "-p", "0.5"], 12.53, 1224.0, 1437.0),
    "buckled_D4": (["-q", "0", "-c", "0.005", "-t", "0", "-v", "0", "-b", "1", "-s", "1000", "-B", "y"],
                   12.3795, 26.86, 127.69),
    # README.md:55-65: the yin-yang (-N 2 -H y) and cube (-H c, -b 80) patterns
    "yinyang_D4": (["-q", "600", "-c", "0.005", "-N", "2", "-p", "0.5", "-H", "y"], 12.5099, 1218.48, 1433.25),
    "cube_D4": (["-q", "600", "-c", "0.005", "-b", "80", "-H", "c"], 12.712, 2852.72, 3626.28),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(README_FINALS))
def test_readme_final_values(name, tmp_path):
    """The six README recipes, 25 000 steps each, through the C++ host driver."""
    build_host()
    flags, A, U, E = README_FINALS[name]
    g = golden(name if name in ("disc_D4", "rod_D4", "janus_D4", "buckled_D4") else "disc_D4")  # all on infiles/3_4.dat
    M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
    r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", "25000", "--out", str(tmp_path / "outfiles")]
                       + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert abs(res["A"] - A) <= 0.01 * A, (res, A)
    assert abs(res["U"] - U) <= 0.01 * U, (res, U)
    assert abs(res["E"] - E) <= 0.01 * E, (res, E)
    rows = [l.split() for l in open(tmp_path / "outfiles" / "energy_nanomembrane.dat")]
    assert len(rows) == 3 + 25000 // 500  # steps 0,1,2 and every 500


@pytest.mark.gpu
def test_host_cpp_snapshot_writers_match_reference(tmp_path):
    """p.lammpstrj (interface_movie), <num>.off (interface_off) and the four local_*_E.off maps of a 5000-step README
    disc run against what the compiled reference wrote for the same run (tests/golden/writers_disc_D4.npz): same
    frames at the same steps, positions / vertex areas / charge densities / per-face energies to trajectory accuracy."""
    from oracle import refdump as R
    build_host()
    g = golden("disc_D4")
    w = np.load(os.path.join(ROOT, "tests", "golden", "writers_disc_D4.npz"))
    K = int(w["steps"])
    M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
    out = tmp_path / "outfiles"
    r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", str(K), "--out", str(out)] + RECIPES["disc_D4"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for label in ("electrostatic range", "elastic range", "stretching range", "bending range"):
        assert label in r.stdout  # the scale-bar lines of src/interface.cpp:1166,1221,1305-1308
    # movie: the driver adds the initial frame ("time step -1", src/main.cpp:417), then one frame per 1000 steps
    steps, frames = R.parse_lammpstrj(str(out / "p.lammpstrj"))
    assert steps.tolist() == [-1] + [int(x) for x in w["movie_all_timesteps"]]
    n = g.nv()
    assert np.allclose(frames[0][:, 2:5], g["pos0"], atol=1e-11) and frames.shape == (len(steps), n, 8)
    for ts, ref in zip(w["movie_timesteps"], w["movie_frames"]):
        mine = frames[steps.tolist().index(int(ts))]
        assert np.array_equal(mine[:, :2], ref[:, :2])                       # index, type
        assert np.max(np.abs(mine[:, 2:5] - ref[:, 2:5])) <= 1e-6           # x y z (unit-sphere scale)
        assert np.max(np.abs(mine[:, 5] - ref[:, 5])) <= 1e-11              # Vq
        assert np.max(np.abs(mine[:, 6] / ref[:, 6] - 1)) <= 1e-5           # Va
        assert np.max(np.abs(mine[:, 7] / ref[:, 7] - 1)) <= 1e-5           # Vq/a
    pos, faces, colours = R.parse_shape_off(str(out / ("%d.off" % K)))
    assert np.array_equal(faces, w["off_faces"]) and np.array_equal(colours, w["off_colours"])
    assert np.max(np.abs(pos - w["off_pos"])) <= 2e-5                        # 6 printed digits
    assert not (out / "2500.off").exists()  # only when both moviefreq and offfreq divide the step, src/newmd.cpp:231-237
    for key, fn in R.LOCAL_OFF.items():
        p2, e = R.parse_local_off(str(out / fn))
        ref = w["local_" + key]
        tol = {"stretching": 5e-3, "elastic": 5e-3, "bending": 1e-3, "electrostatic": 1e-4}[key]
        assert np.max(np.abs(e - ref)) <= tol * float(np.max(np.abs(ref))), (key, np.max(np.abs(e - ref)))
        assert np.max(np.abs(p2 - w["off_pos"])) <= 2e-5


@pytest.mark.gpu
def test_host_cpp_runs_with_explicit_counterions(tmp_path):
    """md_interface of the C++ mirror with a non-empty vector<PARTICLE> (read from a counterions.xyz in the layout
    INTERFACE::put_counterions writes): A / U / E after 500 steps against the compiled reference's run with the same 64
    ions (tests/golden/disc_D4_ions.npz), ions in the movie file."""
    from oracle import refdump as R
    build_host()
    g = golden("disc_D4_ions")
    M.write_dat(str(tmp_path / "m.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"]))
    ions = g["ion_pos0"]
    with open(tmp_path / "counterions.xyz", "w") as fh:
        fh.write("%d\ncounterions\n" % len(ions))
        for p in ions:
            fh.write("C %.17g %.17g %.17g %.17g\n" % (p[0], p[1], p[2], float(np.linalg.norm(p))))
    k = g.traj_steps[-1]
    out = tmp_path / "outfiles"
    r = subprocess.run([BIN, "--mesh", str(tmp_path / "m.dat"), "-S", str(k), "--out", str(out), "--ions",
                        str(tmp_path / "counterions.xyz")] + RECIPES["disc_D4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    s = g.scalars(k)
    cons = s["energy"] + s["mesh_bath_ke"] + s["mesh_bath_pe"] + s["ions_bath_ke"] + s["ions_bath_pe"]
    assert abs(res["A"] - s["total_area"]) <= 1e-6 * s["total_area"]
    assert abs(res["U"] - s["penergy"]) <= 1e-6 * abs(s["penergy"])
    assert abs(res["E"] - cons) <= 1e-6 * abs(cons)
    steps, frames = R.parse_lammpstrj(str(out / "p.lammpstrj"))
    assert frames.shape[1] == g.nv() + len(ions) and set(frames[0][g.nv():, 1].tolist()) == {2.0}  # ions are type 2
