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


def vertex_relerr(a, b, floor_frac=1e-6):
    """Per-vertex check: max_i |a_i - b_i| / max(|b_i|, floor_frac * max_j |b_j|) over [N,3] vector fields
    (north_star: per-vertex forces to 1e-10; the floor keeps vertices whose force nearly cancels meaningful)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b, axis=1)
    den = np.maximum(nb, floor_frac * float(nb.max()) if nb.size else 0.0)
    den = np.where(den > 0, den, 1.0)
    return float(np.max(np.linalg.norm(a - b, axis=1) / den)) if nb.size else 0.0


def make_oracle(g):
    from oracle import oracle as O
    o = O.Oracle(g["edges"], g["faces"], g["l0"], g.params)
    o.set_state(g["pos0"], np.zeros_like(g["pos0"]), g["q"])
    o.set_baths(g["k0_mesh_bath"], g["k0_ions_bath"])
    return o
