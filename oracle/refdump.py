"""TEST INFRASTRUCTURE ONLY.

Runs ``oracle/_ref/npsl_ref`` (the reference's own translation units, see oracle/Makefile and
oracle/ref_driver.cpp) in a scratch working directory and parses its ``--mode dump`` output.
Imported only by ``oracle/gen_golden.py``, ``tests/`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "npsl_ref")

SCALARS = ["total_area", "total_volume", "netMeshKE", "penergy", "energy", "bE", "sE", "tE", "volE", "ljE",
           "esE", "vertex_ke", "mesh_bath_ke", "mesh_bath_pe", "ions_bath_ke", "ions_bath_pe"]
PARAMS = ["scalefactor", "em", "inv_kappa_out", "elj", "lj_length_mesh_mesh", "bkappa", "sconstant", "sigma_a",
          "sigma_v", "ref_area", "ref_volume", "avg_edge_length", "dt", "T", "buckling"]
VECS = {"pos": "pos", "vel": "vel", "for": "force", "bfo": "bforce", "sfo": "sforce", "tfo": "tforce",
        "vfo": "vforce", "pfo": "pforce"}


def have_ref() -> bool:
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def make_workdir(mesh_file: str, D: int, root: str | None = None) -> str:
    """Scratch cwd with ``infiles/3_<D>.dat`` and an empty ``outfiles/`` (the reference appends to
    its series files, so every run gets a fresh directory)."""
    wd = tempfile.mkdtemp(prefix="npsl_ref_", dir=root)
    os.makedirs(os.path.join(wd, "infiles"))
    os.makedirs(os.path.join(wd, "outfiles"))
    dst = os.path.join(wd, "infiles", "3_%d.dat" % D)
    if os.path.abspath(mesh_file) != dst:
        shutil.copy(mesh_file, dst)
    return wd


def reset_outfiles(wd: str) -> None:
    shutil.rmtree(os.path.join(wd, "outfiles"), ignore_errors=True)
    os.makedirs(os.path.join(wd, "outfiles"))


def run_ref(wd: str, args: list[str], threads: int | None = None, timeout: float | None = None) -> str:
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    out = subprocess.run([REF_BIN] + [str(a) for a in args], cwd=wd, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("npsl_ref failed (%d): %s" % (out.returncode, out.stderr[-2000:]))
    return out.stdout


def last_json_line(stdout: str) -> dict:
    for line in reversed(stdout.strip().splitlines()):
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            return json.loads(line)
    raise RuntimeError("no JSON line in npsl_ref output")


def parse_dump(path: str) -> dict:
    """Dump file -> dict of numpy arrays (long-double text is rounded once to float64)."""
    d: dict = {}
    with open(path) as fh:
        lines = fh.read().splitlines()
    n = e = f = 0
    vec: dict = {}
    baths: dict = {"mesh_bath": [], "ions_bath": []}
    edges, faces = [], []
    varea, lface = [], []
    ions = {"ionq": [], "ipos": [], "ivel": [], "ifor": []}
    q = None
    for line in lines:
        t = line.split()
        k = t[0]
        if k == "sizes":
            n, e, f = int(t[1]), int(t[2]), int(t[3])
            q = np.zeros(n)
            for name in VECS.values():
                vec[name] = np.zeros((n, 3))
        elif k in VECS:
            vec[VECS[k]][int(t[1])] = (float(t[2]), float(t[3]), float(t[4]))
        elif k == "q":
            q[int(t[1])] = float(t[2])
        elif k in baths:
            baths[k].append([float(t[2]), float(t[3]), float(t[4]), float(t[5]), float(t[6])])
        elif k == "edge":
            edges.append([int(t[2]), int(t[3]), float(t[4]), float(t[5])])
        elif k == "face":
            faces.append([int(t[2]), int(t[3]), int(t[4]), float(t[5]), float(t[6])])
        elif k in ions:
            ions[k].append([float(x) for x in t[2:]])
        elif k == "varea":
            varea.append([float(x) for x in t[2:6]])
        elif k == "lface":
            lface.append([float(x) for x in t[2:6]])
        elif len(t) == 2:
            d[k] = float(t[1])
    d.update(vec)
    d["q"] = q
    d["mesh_bath"] = np.array(baths["mesh_bath"])
    d["ions_bath"] = np.array(baths["ions_bath"])
    ea = np.array(edges)
    fa = np.array(faces)
    d["edges"] = ea[:, :2].astype(np.int32)
    d["l0"] = ea[:, 2].copy()
    d["itsS"] = ea[:, 3].copy()
    d["faces"] = fa[:, :3].astype(np.int32)
    d["face_area"] = fa[:, 3].copy()
    d["face_volume"] = fa[:, 4].copy()
    if varea:  # --local 1
        va = np.array(varea)
        d["vertex_area"], d["vertex_normal"] = va[:, 0].copy(), va[:, 1:4].copy()
        d["local_faces"] = np.array(lface)  # [F, 4]: electrostatic, elastic, bending, stretching
    if ions["ionq"]:  # --ions M
        d["ion_q"] = np.array(ions["ionq"])[:, 0].copy()
        d["ion_pos"], d["ion_vel"], d["ion_force"] = np.array(ions["ipos"]), np.array(ions["ivel"]), np.array(ions["ifor"])
    d["scalars"] = np.array([d[s] for s in SCALARS])
    d["params"] = np.array([d[s] for s in PARAMS])
    assert d["edges"].shape[0] == e and d["faces"].shape[0] == f
    return d


def dump_at_step(wd: str, D: int, flags: list[str], step: int, threads: int | None = None, writedata: int = 500,
                 local: bool = False) -> dict:
    reset_outfiles(wd)
    out = os.path.join(wd, "dump_%d.txt" % step)
    run_ref(wd, ["--mode", "dump", "--out", out, "-D", D, "--steps", step, "-W", writedata] + (["--local", 1] if local else []) + flags,
            threads=threads)
    d = parse_dump(out)
    if local:  # the reference's own files (6 significant digits)
        for key, fn in LOCAL_OFF.items():
            d["off_" + key] = parse_local_off(os.path.join(wd, "outfiles", fn))[1]
    return d


LOCAL_OFF = {"electrostatic": "local_electrostatic_E.off", "elastic": "local_elastic_E.off",
             "bending": "local_bending_E.off", "stretching": "local_stretching_E.off"}


def parse_local_off(path: str):
    """local_*_E.off of INTERFACE::compute_local_energies (src/interface.cpp:1140-1337): vertex positions and, per
    face, colormap(E) = (r, g, b, a) = (0, E, 1 - E, 1). Returns (positions [V,3], E per face [F])."""
    with open(path) as fh:
        t = fh.read().split()
    assert t[0] == "OFF"
    nv, nf = int(t[1]), int(t[2])
    pos = np.array(t[4:4 + 3 * nv], dtype=np.float64).reshape(nv, 3)
    rows = np.array(t[4 + 3 * nv:4 + 3 * nv + 8 * nf], dtype=np.float64).reshape(nf, 8)
    return pos, rows[:, 5].copy()


def parse_lammpstrj(path: str):
    """p.lammpstrj of interface_movie (src/functions.cpp:50-84) -> (timesteps [T], rows [T, N, 8]:
    index type x y z Vq Va Vq/a)."""
    steps, frames = [], []
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = 0
    while i < len(lines):
        assert lines[i].startswith("ITEM: TIMESTEP"), lines[i]
        steps.append(int(lines[i + 1]))
        n = int(lines[i + 3])
        rows = [[float(x) for x in lines[i + 9 + k].split()] for k in range(n)]
        frames.append(rows)
        i += 9 + n
    return np.array(steps), np.array(frames)


def parse_shape_off(path: str):
    """<num>.off of interface_off (src/functions.cpp:134-167) -> (positions [V,3], faces [F,3], colours [F,4])."""
    with open(path) as fh:
        t = fh.read().split()
    assert t[0] == "OFF"
    nv, nf = int(t[1]), int(t[2])
    pos = np.array(t[4:4 + 3 * nv], dtype=np.float64).reshape(nv, 3)
    rows = np.array(t[4 + 3 * nv:4 + 3 * nv + 8 * nf], dtype=np.float64).reshape(nf, 8)
    return pos, rows[:, 1:4].astype(np.int32), rows[:, 4:8].copy()
