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
        self.z = np.load(os.path.join(GOLD, name + ".npz"))
        self.params = dict(zip([str(x) for x in self.z["param_names"]], [float(x) for x in self.z["params"]]))
        self.scalar_names = [str(x) for x in self.z["scalar_names"]]
        self.full_steps = [int(x) for x in self.z["full_steps"]]
        self.traj_steps = [int(x) for x in self.z["traj_steps"]]

    def __getitem__(self, k):
        return self.z[k]

    def scalars(self, step):
        return dict(zip(self.scalar_names, [float(x) for x in self.z["k%d_scalars" % step]]))

    def nv(self):
        return int(self.z["pos0"].shape[0])


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


def make_oracle(g):
    from oracle import oracle as O
    o = O.Oracle(g["edges"], g["faces"], g["l0"], g.params)
    o.set_state(g["pos0"], np.zeros_like(g["pos0"]), g["q"])
    o.set_baths(g["k0_mesh_bath"], g["k0_ions_bath"])
    return o
