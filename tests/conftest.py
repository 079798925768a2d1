"""Shared fixtures. GPU tests are marked ``gpu``; everything else runs on a CPU-only box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ["disc_D4", "rod_D4", "janus_D4", "buckled_D4", "janus_D8", "buckling_D8", "disc_ico6"]
CONSTRAINED_NAMES = ["disc_D4_GV", "janus_D8_GV"]  # rigid volume constraint (-G V)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One tests/golden/<name>.npz fixture: outputs of the reference's own code (oracle/gen_golden.py)."""

    def __init__(self, name):
        self.name = name
        self._mesh = None
        self.z = np.load(os.path.join(GOLD, name + ".npz"))
        self.params = dict(zip([str(x) for x in self.z["param_names"]], [float(x) for x in self.z["params"]]))
        self.scalar_names = [str(x) for x in self.z["scalar_names"]]
        self.full_steps = [int(x) for x in self.z["full_steps"]]
        self.traj_steps = [int(x) for x in self.z["traj_steps"]]

    def __getitem__(self, k):
        if k in self.z.files:
            return self.z[k]
        if "mesh_source" in self.z.files:  # large synthetic meshes: topology and initial state come from the generator
            m = self.mesh()
            if k in ("pos0", "k0_pos"):
                return m.pos
            if k == "k0_vel":
                return np.zeros_like(m.pos)
            if k in ("edges", "faces"):
                return getattr(m, k)
        raise KeyError(k)

    def mesh(self):
        if self._mesh is None:
            from np_shape_lab_b200 import mesh as M
            src = str(self.z["mesh_source"])
            assert src.startswith("ico:")
            self._mesh = M.icosphere(int(src[4:]))
        return self._mesh

    def scalars(self, step):
        return dict(zip(self.scalar_names, [float(x) for x in self.z["k%d_scalars" % step]]))

    def nv(self):
        return int(self["pos0"].shape[0])


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


@pytest.fixture(params=GOLDEN_NAMES)
def gold(request):
    return golden(request.param)


def relerr(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor) -- the per-field metric of SURVEY 8d."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))) if b.size else 0.0, floor)
    num = float(np.max(np.abs(a - b))) if b.size else 0.0
    return num / den if den > 0 else num


def vertex_relerr(a, b, floor_frac=1e-6, cond=None, abs_allow=0.0):
    """Per-vertex check: max_i |a_i - b_i| / max(|b_i|, floor_frac * max_j |b_j|, 1e-3 * cond_i) over [N,3] vector
    fields (north_star: per-vertex forces to 1e-10; the floor keeps vertices whose force nearly cancels meaningful).
    cond_i, where given, is the magnitude of the terms whose sum is b_i (bending_term_scale): a sum of terms of size
    cond_i cannot be reproduced to better than ~1e-16 cond_i in FP64 -- the reference itself, which works in x87
    extended precision but stores EDGE::itsS and the face areas as doubles, moves by ~6e-12 at S64 when its own
    bending sum is merely re-ordered -- so the tolerance 1e-10 is applied to max(|b_i|, 1e-3 cond_i), i.e. 1e-13 of
    the summed terms. abs_allow: an absolute allowance subtracted from every vertex's deviation first (the measured
    response of the force to a one-ulp move of the input positions, where the reference's own inputs are not
    representable in FP64: tests/test_gpu_large.py)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b, axis=1)
    den = np.maximum(nb, floor_frac * float(nb.max()) if nb.size else 0.0)
    if cond is not None:
        den = np.maximum(den, 1e-3 * np.asarray(cond))
    den = np.where(den > 0, den, 1.0)
    return float(np.max(np.maximum(np.linalg.norm(a - b, axis=1) - abs_allow, 0.0) / den)) if nb.size else 0.0


def bending_term_scale(edges, faces, pos, bkappa):
    """Per vertex, the summed magnitude of its bending-force contributions (EDGE::bending_forces adds one term per
    incident / opposite edge, src/edge.cpp:38-58): on a smooth closed surface these terms are O(bkappa / edge) each
    and cancel to a few per mille, which is what limits how well any FP64 evaluation can match the sum."""
    pos = np.asarray(pos, dtype=np.float64)
    third = {}
    for f in np.asarray(faces).tolist():
        for k in range(3):
            third[(f[k], f[(k + 1) % 3])] = f[(k + 2) % 3]
    rec = np.array([[a, b, third[(a, b)], third[(b, a)]] for a, b in np.asarray(edges).tolist()])
    p0, p1, pa, pb = [pos[rec[:, k]] for k in range(4)]
    dot = lambda u, v: np.einsum("ij,ij->i", u, v)
    e01 = p1 - p0
    dir0, dir1 = np.cross(e01, pa - p1), np.cross(p0 - p1, pb - p0)
    S, n0, n1 = dot(dir0, dir1), dot(dir0, dir0), dot(dir1, dir1)
    kb = bkappa / np.sqrt(n0 * n1)
    w0, w1 = dir1 - (S / n0)[:, None] * dir0, dir0 - (S / n1)[:, None] * dir1
    slots = [kb[:, None] * (np.cross(p1 - pa, w0) + np.cross(pb - p1, w1)),
             kb[:, None] * (np.cross(pa - p0, w0) + np.cross(p0 - pb, w1)),
             kb[:, None] * np.cross(p0 - p1, w0), kb[:, None] * np.cross(e01, w1)]
    out = np.zeros(pos.shape[0])
    for k, t in enumerate(slots):
        np.add.at(out, rec[:, k], np.linalg.norm(t, axis=1))
    return out


def make_oracle(g):
    from oracle import oracle as O
    o = O.Oracle(g["edges"], g["faces"], g["l0"], g.params)
    o.set_state(g["pos0"], np.zeros_like(g["pos0"]), g["q"])
    o.set_baths(g["k0_mesh_bath"], g["k0_ions_bath"])
    return o
