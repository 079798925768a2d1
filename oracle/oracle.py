"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libnpsl_oracle.so (the C restatement,
oracle/npsl_oracle.c). Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnpsl_oracle.so")

PARAM_ORDER = ["scalefactor", "em", "inv_kappa_out", "elj", "lj_length_mesh_mesh", "bkappa", "sconstant", "sigma_a",
               "sigma_v", "ref_area", "ref_volume", "avg_edge_length", "dt", "T", "buckling"]

_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "npsl_oracle.c")):
            build()
        L = C.CDLL(LIB)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.npo_create.restype = vp
        L.npo_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp]
        L.npo_destroy.argtypes = [vp]
        L.npo_set_state.argtypes = [vp, dp, dp, dp]
        L.npo_set_baths.argtypes = [vp, C.c_int, dp, dp]
        L.npo_get_baths.argtypes = [vp, dp, dp]
        L.npo_dressup.argtypes = [vp]
        L.npo_force.argtypes = [vp, C.c_int, C.c_int]
        L.npo_energy.argtypes = [vp, C.c_int, C.c_int, dp]
        L.npo_md_init.argtypes = [vp]
        L.npo_set_constraint.argtypes = [vp, C.c_int]
        L.npo_constraint_init.argtypes = [vp]
        L.npo_refresh_ke.argtypes = [vp]
        L.npo_md_step.argtypes = [vp]
        L.npo_md_steps.argtypes = [vp, C.c_int]
        L.npo_bath_energies.argtypes = [vp, dp]
        L.npo_get_vec.argtypes = [vp, C.c_int, dp]
        L.npo_get_scalars.argtypes = [vp, dp]
        L.npo_get_edge_S.argtypes = [vp, dp]
        L.npo_set_ions.argtypes = [vp, C.c_int, dp, dp, dp, C.c_double, C.c_double]
        L.npo_get_ions.argtypes = [vp, C.c_int, dp]
        L.npo_ions_ke.argtypes = [vp]
        L.npo_ions_ke.restype = C.c_double
        L.npo_vertex_geometry.argtypes = [vp, dp, dp]
        L.npo_local_energies.argtypes = [vp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


VEC = {"pos": 0, "vel": 1, "force": 2, "pforce": 3, "bforce": 4, "sforce": 5, "tforce": 6, "vforce": 7}
ENERGY_NAMES = ["KE", "bE", "sE", "tE", "ljE", "esE", "volE", "penergy", "energy"]


class Oracle:
    """CPU checker over flat arrays; mirrors force_calculation / compute_energy / md_interface."""

    def __init__(self, edges, faces, l0, params: dict):
        self.L = lib()
        self.edges = np.ascontiguousarray(edges, dtype=np.int32)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32)
        self.nv = int(max(self.edges.max(), self.faces.max())) + 1
        self.ne, self.nf = self.edges.shape[0], self.faces.shape[0]
        l0 = np.ascontiguousarray(l0, dtype=np.float64)
        p = np.array([float(params[k]) for k in PARAM_ORDER], dtype=np.float64)
        self.h = self.L.npo_create(self.nv, self.ne, self.nf, _ip(self.edges), _ip(self.faces), _dp(l0), _dp(p))
        if not self.h:
            raise RuntimeError("npo_create failed (bad topology)")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.npo_destroy(self.h)
            self.h = None

    def set_state(self, pos=None, vel=None, q=None):
        def prep(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        pos, vel, q = prep(pos), prep(vel), prep(q)
        self.L.npo_set_state(self.h, _dp(pos) if pos is not None else None, _dp(vel) if vel is not None else None,
                             _dp(q) if q is not None else None)

    def set_baths(self, mesh_bath, ions_bath):
        m = np.ascontiguousarray(mesh_bath, dtype=np.float64)
        i = np.ascontiguousarray(ions_bath, dtype=np.float64)
        self.chain = m.shape[0]
        self.L.npo_set_baths(self.h, self.chain, _dp(m), _dp(i))

    def get_baths(self):
        m = np.zeros((self.chain, 5))
        i = np.zeros((self.chain, 5))
        self.L.npo_get_baths(self.h, _dp(m), _dp(i))
        return m, i

    def dressup(self):
        self.L.npo_dressup(self.h)

    def force(self, lo=0, hi=None):
        self.L.npo_force(self.h, lo, self.nv - 1 if hi is None else hi)

    def energy(self, lo=0, hi=None) -> dict:
        out = np.zeros(9)
        self.L.npo_energy(self.h, lo, self.nv - 1 if hi is None else hi, _dp(out))
        return dict(zip(ENERGY_NAMES, out.tolist()))

    def md_init(self):
        self.L.npo_md_init(self.h)

    def set_constraint(self, on: bool = True):
        """Rigid volume constraint (-G V, linear form): SHAKE / RATTLE inside md_steps."""
        self.L.npo_set_constraint(self.h, 1 if on else 0)

    def constraint_init(self):
        self.L.npo_constraint_init(self.h)

    def refresh_ke(self):
        self.L.npo_refresh_ke(self.h)

    def md_steps(self, n: int):
        self.L.npo_md_steps(self.h, int(n))

    def bath_energies(self):
        out = np.zeros(4)
        self.L.npo_bath_energies(self.h, _dp(out))
        return out

    def vec(self, name: str) -> np.ndarray:
        out = np.zeros((self.nv, 3))
        self.L.npo_get_vec(self.h, VEC[name], _dp(out))
        return out

    def scalars(self) -> dict:
        out = np.zeros(3)
        self.L.npo_get_scalars(self.h, _dp(out))
        return {"total_area": out[0], "total_volume": out[1], "vertex_ke": out[2]}

    def edge_S(self) -> np.ndarray:
        out = np.zeros(self.ne)
        self.L.npo_get_edge_S(self.h, _dp(out))
        return out

    def vertex_geometry(self) -> dict:
        """VERTEX::compute_area_normal for every vertex (after dressup)."""
        area, normal = np.zeros(self.nv), np.zeros((self.nv, 3))
        self.L.npo_vertex_geometry(self.h, _dp(area), _dp(normal))
        return {"area": area, "normal": normal}

    def local_energies(self) -> dict:
        """Per-face maps of INTERFACE::compute_local_energies(_by_component) (after force)."""
        out = np.zeros((self.nf, 4))
        self.L.npo_local_energies(self.h, _dp(out))
        return {"electrostatic": out[:, 0].copy(), "elastic": out[:, 1].copy(), "bending": out[:, 2].copy(),
                "stretching": out[:, 3].copy()}

    def set_ions(self, pos, vel, q, diameter: float, lj_length_mesh_ions: float):
        """Explicit counterions (vector<PARTICLE>): positions, velocities (or None), charges."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self.n_ions = int(pos.shape[0])
        self.L.npo_set_ions(self.h, self.n_ions, _dp(pos), _dp(vel) if vel is not None else None, _dp(q), float(diameter),
                            float(lj_length_mesh_ions))

    def ions(self, what: str) -> np.ndarray:
        out = np.zeros((self.n_ions, 3))
        self.L.npo_get_ions(self.h, {"pos": 0, "vel": 1, "force": 2}[what], _dp(out))
        return out

    def ions_ke(self) -> float:
        """INTERFACE::netIonKE of the last energy() call."""
        return float(self.L.npo_ions_ke(self.h))
