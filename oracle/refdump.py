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
    d["scalars"] = np.array([d[s] for s in SCALARS])
    d["params"] = np.array([d[s] for s in PARAMS])
    assert d["edges"].shape[0] == e and d["faces"].shape[0] == f
    return d


def dump_at_step(wd: str, D: int, flags: list[str], step: int, threads: int | None = None, writedata: int = 500) -> dict:
    reset_outfiles(wd)
    out = os.path.join(wd, "dump_%d.txt" % step)
    run_ref(wd, ["--mode", "dump", "--out", out, "-D", D, "--steps", step, "-W", writedata] + flags, threads=threads)
    return parse_dump(out)
