"""ctypes binding of ``libnpsl.so`` (include/npsl.h), one Python method per C entry point.

This is plumbing only: every method forwards plain pointers and sizes to the C ABI and turns a
negative status into :class:`NpslError`. There is no CPU path here or behind it -- if the
library is missing, or the box has no CUDA device, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# NPSL_LIB: measurement only (A/B of two builds of the library on the GPU box); the product path is libnpsl.so in-tree
LIB_PATH = os.environ.get("NPSL_LIB") or os.path.join(PKG, "libnpsl.so")

NPSL_MAX_CHAIN = 16
NPSL_NCCL_ID_BYTES = 128
ABI_VERSION = 2

# every symbol include/npsl.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "npsl_create", "npsl_create_sharded", "npsl_nccl_unique_id", "npsl_destroy", "npsl_last_error",
    "npsl_abi_version", "npsl_upload_state", "npsl_download_state", "npsl_download_force_components",
    "npsl_download_mesh_fields", "npsl_get_thermostat", "npsl_set_thermostat", "npsl_dressup", "npsl_force_init",
    "npsl_force", "npsl_step", "npsl_energy", "npsl_sync", "npsl_launch_count", "npsl_pair_timing",
    "npsl_pair_timing_read", "npsl_owned_rows", "npsl_pair_geometry", "npsl_fp64_peak", "npsl_md_run",
    "npsl_step_timing", "npsl_step_timing_list", "npsl_set_pair_plan", "npsl_force_calculation", "npsl_pair_poly_eval", "npsl_pair_fast_eval", "npsl_pair_math", "npsl_pair_kernel_info", "npsl_md_step_host", "npsl_host_state", "npsl_md_step_packed", "npsl_set_pair_mode",
    "npsl_pair_mode", "npsl_exchange_mode", "npsl_shard_plan", "npsl_vertex_geometry", "npsl_local_energies", "npsl_step_profile", "npsl_set_volume_constraint", "npsl_constraint_init",
    "npsl_constraint_multipliers", "npsl_set_ions", "npsl_download_ions", "npsl_plan_pair_units",
]


class NpslError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("npsl status %d: %s" % (status, message))
        self.status = status


class MeshDesc(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_edges", C.c_int32), ("n_faces", C.c_int32),
                ("edge_v", C.POINTER(C.c_int32)), ("face_v", C.POINTER(C.c_int32)), ("l0", C.POINTER(C.c_double))]


class BathLink(C.Structure):
    _fields_ = [("Q", C.c_double), ("T", C.c_double), ("dof", C.c_double), ("xi", C.c_double), ("eta", C.c_double)]


class Params(C.Structure):
    _fields_ = [("scalefactor", C.c_double), ("em", C.c_double), ("inv_kappa_out", C.c_double), ("elj", C.c_double),
                ("lj_length_mesh_mesh", C.c_double), ("bkappa", C.c_double), ("sconstant", C.c_double),
                ("sigma_a", C.c_double), ("sigma_v", C.c_double), ("ref_volume", C.c_double),
                ("avg_edge_length", C.c_double), ("dt", C.c_double), ("buckling", C.c_int32), ("chain", C.c_int32),
                ("mesh_bath", BathLink * NPSL_MAX_CHAIN), ("ions_bath", BathLink * NPSL_MAX_CHAIN)]


ENERGY_FIELDS = ["kinetic", "bending", "stretching", "tension", "lj", "electrostatic", "volume", "potential",
                 "energy", "mesh_bath_ke", "mesh_bath_pe", "ions_bath_ke", "ions_bath_pe", "total_area",
                 "total_volume", "ions_kinetic"]


class Energies(C.Structure):
    _fields_ = [(n, C.c_double) for n in ENERGY_FIELDS]

    def as_dict(self) -> dict:
        return {n: float(getattr(self, n)) for n in ENERGY_FIELDS}


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library. Raises if it has not been built (``python -m
    np_shape_lab_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NpslError(-2, "libnpsl.so is not built (%s); run __graft_entry__.build() -- there is no CPU path"
                        % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.npsl_abi_version.restype = C.c_int
    L.npsl_last_error.restype = C.c_char_p
    L.npsl_last_error.argtypes = [vp]
    L.npsl_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(Params), C.POINTER(vp)]
    L.npsl_create_sharded.argtypes = [C.POINTER(MeshDesc), C.POINTER(Params), C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.npsl_nccl_unique_id.argtypes = [vp]
    L.npsl_destroy.argtypes = [vp]
    L.npsl_destroy.restype = None
    L.npsl_upload_state.argtypes = [vp, dp, dp, dp]
    L.npsl_download_state.argtypes = [vp, dp, dp, dp]
    L.npsl_download_force_components.argtypes = [vp, dp, dp, dp, dp, dp]
    L.npsl_download_mesh_fields.argtypes = [vp, dp, dp, dp, dp]
    L.npsl_get_thermostat.argtypes = [vp, C.POINTER(BathLink), C.POINTER(BathLink)]
    L.npsl_set_thermostat.argtypes = [vp, C.POINTER(BathLink), C.POINTER(BathLink)]
    L.npsl_dressup.argtypes = [vp, dp, dp]
    L.npsl_force_init.argtypes = [vp]
    L.npsl_force.argtypes = [vp]
    L.npsl_step.argtypes = [vp, C.c_int]
    L.npsl_energy.argtypes = [vp, C.POINTER(Energies)]
    L.npsl_sync.argtypes = [vp]
    L.npsl_launch_count.argtypes = [vp]
    L.npsl_launch_count.restype = C.c_int64
    L.npsl_pair_timing.argtypes = [vp, C.c_int]
    L.npsl_pair_timing_read.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.npsl_owned_rows.argtypes = [vp, ip, ip]
    L.npsl_pair_geometry.argtypes = [vp, ip, ip, ip, ip]
    L.npsl_fp64_peak.argtypes = [dp, dp]
    L.npsl_md_run.argtypes = [vp, dp, dp, dp, C.c_int, C.POINTER(Energies), dp, dp]
    L.npsl_step_timing.argtypes = [vp, C.c_int, C.c_int, dp]
    L.npsl_step_timing_list.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    L.npsl_force_calculation.argtypes = [vp, dp, dp]
    L.npsl_md_step_host.argtypes = [vp, dp, dp, dp, C.c_int, dp]
    L.npsl_host_state.argtypes = [vp] + [C.POINTER(dp)] * 4
    L.npsl_pair_poly_eval.argtypes = [C.c_double, dp, dp, dp, C.c_int]
    L.npsl_pair_fast_eval.argtypes = [dp, dp, dp, C.c_int, C.c_int]
    L.npsl_pair_math.argtypes = [vp]
    L.npsl_pair_kernel_info.argtypes = [vp, ip, ip, dp]
    L.npsl_md_step_packed.argtypes = [vp, C.c_int]
    L.npsl_set_pair_plan.argtypes = [vp, C.c_int, C.c_int]
    L.npsl_set_pair_mode.argtypes = [vp, C.c_int]
    L.npsl_pair_mode.argtypes = [vp]
    L.npsl_exchange_mode.argtypes = [vp]
    L.npsl_shard_plan.argtypes = [vp]
    L.npsl_plan_pair_units.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_int, ip]
    L.npsl_set_ions.argtypes = [vp, C.c_int, dp, dp, dp, C.c_double, C.c_double]
    L.npsl_download_ions.argtypes = [vp, dp, dp, dp]
    L.npsl_vertex_geometry.argtypes = [vp, dp, dp]
    L.npsl_local_energies.argtypes = [vp, dp, dp, dp, dp]
    L.npsl_set_volume_constraint.argtypes = [vp, C.c_int]
    L.npsl_constraint_init.argtypes = [vp]
    L.npsl_constraint_multipliers.argtypes = [vp, dp, dp]
    L.npsl_step_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    if L.npsl_abi_version() != ABI_VERSION:
        raise NpslError(-1, "libnpsl.so ABI version %d, binding expects %d" % (L.npsl_abi_version(), ABI_VERSION))
    _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def fill_bath(dst, rows) -> None:
    rows = np.asarray(rows, dtype=np.float64)
    for j in range(rows.shape[0]):
        dst[j].Q, dst[j].T, dst[j].dof, dst[j].xi, dst[j].eta = [float(x) for x in rows[j]]


def make_params(p: dict, mesh_bath, ions_bath) -> Params:
    """``p``: the dict of :func:`mesh.derive_params` (or the golden fixtures' parameter list)."""
    P = Params()
    for k in ["scalefactor", "em", "inv_kappa_out", "elj", "lj_length_mesh_mesh", "bkappa", "sconstant", "sigma_a",
              "sigma_v", "ref_volume", "avg_edge_length", "dt"]:
        setattr(P, k, float(p[k]))
    P.buckling = int(p["buckling"])
    mesh_bath = np.asarray(mesh_bath, dtype=np.float64)
    ions_bath = np.asarray(ions_bath, dtype=np.float64)
    if mesh_bath.shape != ions_bath.shape or mesh_bath.shape[1] != 5 or mesh_bath.shape[0] > NPSL_MAX_CHAIN:
        raise ValueError("baths must both be [C<=%d,5] arrays of (Q,T,dof,xi,eta)" % NPSL_MAX_CHAIN)
    P.chain = int(mesh_bath.shape[0])
    fill_bath(P.mesh_bath, mesh_bath)
    fill_bath(P.ions_bath, ions_bath)
    return P


def plan_pair_units(n_vertices: int, rank: int = 0, world: int = 1):
    """Units of the symmetric pair kernel for one rank, without a device: (units [k,4] = row block, chunk, first column,
    one past the last column; geometry dict)."""
    L = load()
    geo = np.zeros(6, dtype=np.int32)
    n = L.npsl_plan_pair_units(n_vertices, rank, world, None, 0, geo.ctypes.data_as(C.POINTER(C.c_int32)))
    if n < 0:
        raise NpslError(n, (L.npsl_last_error(None) or b"").decode())
    units = np.zeros((n, 4), dtype=np.int32)
    L.npsl_plan_pair_units(n_vertices, rank, world, units.ctypes.data_as(C.POINTER(C.c_int32)), n, None)
    return units, dict(zip(["rows_per_block", "column_granularity", "chunks", "full_chunks", "tiles_full", "tiles_half"],
                           [int(x) for x in geo]))


def pair_fast_eval(s: np.ndarray, seed_bits: int = 20):
    """(device-order value, extended-precision value) of e^{-r}(1/r^3 + 1/r^2) at s = r^2 (lambda units) for the default
    reduced-accuracy scheme of k_pair_sym with a seed of seed_bits bits (npsl_pair_fast_eval; host only)."""
    L = load()
    s = np.ascontiguousarray(s, dtype=np.float64)
    v, ev = np.zeros_like(s), np.zeros_like(s)
    st = L.npsl_pair_fast_eval(_dp(s), _dp(v), _dp(ev), int(s.size), int(seed_bits))
    if st != 0:
        raise NpslError(st, "npsl_pair_fast_eval")
    return v, ev


def nccl_unique_id() -> bytes:
    L = load()
    buf = C.create_string_buffer(NPSL_NCCL_ID_BYTES)
    st = L.npsl_nccl_unique_id(buf)
    if st != 0:
        raise NpslError(st, (L.npsl_last_error(None) or b"").decode())
    return buf.raw


def pair_poly_eval(inv_kappa_out: float, s: np.ndarray):
    """(table value, extended-precision value) of the pair kernel's radial factor at s (npsl_pair_poly_eval; host only)."""
    L = load()
    s = np.ascontiguousarray(s, dtype=np.float64)
    tv, ev = np.empty_like(s), np.empty_like(s)
    st = L.npsl_pair_poly_eval(float(inv_kappa_out), _dp(s), _dp(tv), _dp(ev), int(s.size))
    if st != 0:
        raise NpslError(st, "npsl_pair_poly_eval")
    return tv, ev


def fp64_peak() -> tuple:
    """(DFMA instructions/s, nominal SM clock MHz) from the pure-DFMA probe kernel."""
    L = load()
    r, f = C.c_double(), C.c_double()
    st = L.npsl_fp64_peak(C.byref(r), C.byref(f))
    if st != 0:
        raise NpslError(st, (L.npsl_last_error(None) or b"").decode())
    return r.value, f.value


class Context:
    """Owner of one ``npsl_ctx``. ``edges`` [E,2], ``faces`` [F,3] (stored winding), ``l0`` [E]."""

    def __init__(self, edges, faces, l0, params: Params, n_vertices: int | None = None, rank: int = 0,
                 world: int = 1, nccl_id: bytes | None = None):
        self.L = load()
        self.h = C.c_void_p()
        self._edges = np.ascontiguousarray(edges, dtype=np.int32)
        self._faces = np.ascontiguousarray(faces, dtype=np.int32)
        self._l0 = np.ascontiguousarray(l0, dtype=np.float64)
        self.nv = int(n_vertices) if n_vertices is not None else int(max(self._edges.max(), self._faces.max())) + 1
        self.ne, self.nf = int(self._edges.shape[0]), int(self._faces.shape[0])
        self.chain = int(params.chain)
        md = MeshDesc(self.nv, self.ne, self.nf, self._edges.ctypes.data_as(C.POINTER(C.c_int32)),
                      self._faces.ctypes.data_as(C.POINTER(C.c_int32)), _dp(self._l0))
        if world == 1:
            st = self.L.npsl_create(C.byref(md), C.byref(params), C.byref(self.h))
        else:
            idbuf = C.create_string_buffer(nccl_id, NPSL_NCCL_ID_BYTES) if nccl_id is not None else None
            st = self.L.npsl_create_sharded(C.byref(md), C.byref(params), rank, world, idbuf, C.byref(self.h))
        if st != 0:
            self.h = C.c_void_p()
            raise NpslError(st, (self.L.npsl_last_error(None) or b"").decode())
        self.rank, self.world = rank, world

    # -- plumbing
    def _ck(self, st: int) -> None:
        if st != 0:
            raise NpslError(st, (self.L.npsl_last_error(self.h) or b"").decode())

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.npsl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state
    def upload_state(self, pos=None, vel=None, q=None) -> None:
        pos, vel, q = _f64(pos, (self.nv, 3)), _f64(vel, (self.nv, 3)), _f64(q, (self.nv,))
        self._ck(self.L.npsl_upload_state(self.h, _dp(pos), _dp(vel), _dp(q)))

    def download_state(self, pos=True, vel=True, force=True) -> dict:
        out = {k: np.empty((self.nv, 3)) for k, want in (("pos", pos), ("vel", vel), ("force", force)) if want}
        self._ck(self.L.npsl_download_state(self.h, _dp(out.get("pos")), _dp(out.get("vel")), _dp(out.get("force"))))
        return out

    def download_force_components(self) -> dict:
        names = ["pair", "bending", "stretching", "tension", "volume_tension"]
        out = {k: np.empty((self.nv, 3)) for k in names}
        self._ck(self.L.npsl_download_force_components(self.h, *[_dp(out[k]) for k in names]))
        return out

    def download_mesh_fields(self) -> dict:
        out = {"face_area": np.empty(self.nf), "face_volume": np.empty(self.nf), "face_dir": np.empty((self.nf, 3)),
               "edge_S": np.empty(self.ne)}
        self._ck(self.L.npsl_download_mesh_fields(self.h, _dp(out["face_area"]), _dp(out["face_volume"]),
                                                   _dp(out["face_dir"]), _dp(out["edge_S"])))
        return out

    def get_thermostat(self):
        m = (BathLink * NPSL_MAX_CHAIN)()
        i = (BathLink * NPSL_MAX_CHAIN)()
        self._ck(self.L.npsl_get_thermostat(self.h, m, i))
        conv = lambda a: np.array([[a[j].Q, a[j].T, a[j].dof, a[j].xi, a[j].eta] for j in range(self.chain)])
        return conv(m), conv(i)

    def set_thermostat(self, mesh_bath, ions_bath) -> None:
        m = (BathLink * NPSL_MAX_CHAIN)()
        i = (BathLink * NPSL_MAX_CHAIN)()
        fill_bath(m, mesh_bath)
        fill_bath(i, ions_bath)
        self._ck(self.L.npsl_set_thermostat(self.h, m, i))

    # -- the path
    def dressup(self):
        a, v = C.c_double(), C.c_double()
        self._ck(self.L.npsl_dressup(self.h, C.byref(a), C.byref(v)))
        return a.value, v.value

    def force_init(self) -> None:
        self._ck(self.L.npsl_force_init(self.h))

    def force(self) -> None:
        self._ck(self.L.npsl_force(self.h))

    def step(self, n: int = 1) -> None:
        self._ck(self.L.npsl_step(self.h, int(n)))

    def energy(self) -> dict:
        e = Energies()
        self._ck(self.L.npsl_energy(self.h, C.byref(e)))
        return e.as_dict()

    def sync(self) -> None:
        self._ck(self.L.npsl_sync(self.h))

    def md_run(self, pos, vel, q, n_steps: int):
        """Host-buffer entry point (npsl_md_run): upload, force_init, n steps, energy, download."""
        pos, vel, q = _f64(pos, (self.nv, 3)), _f64(vel, (self.nv, 3)), _f64(q, (self.nv,))
        e = Energies()
        pos_out, vel_out = np.empty((self.nv, 3)), np.empty((self.nv, 3))
        self._ck(self.L.npsl_md_run(self.h, _dp(pos), _dp(vel), _dp(q), int(n_steps), C.byref(e), _dp(pos_out),
                                    _dp(vel_out)))
        return e.as_dict(), pos_out, vel_out

    # -- introspection
    def launch_count(self) -> int:
        return int(self.L.npsl_launch_count(self.h))

    def pair_timing(self, enable: bool) -> None:
        self._ck(self.L.npsl_pair_timing(self.h, 1 if enable else 0))

    def pair_timing_read(self):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.L.npsl_pair_timing_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def step_timing(self, n_steps: int, flush_l2: bool = False) -> float:
        """Device time (ms, CUDA events on the context's stream) of ``n_steps`` graph-launched steps."""
        ms = C.c_double()
        self._ck(self.L.npsl_step_timing(self.h, int(n_steps), 1 if flush_l2 else 0, C.byref(ms)))
        return ms.value

    def step_timing_list(self, n_steps: int, flush_l2: bool = True, lead_in: int = 1) -> np.ndarray:
        """Device time (ms) of each of ``n_steps`` graph-launched steps after ``lead_in`` untimed ones."""
        out = np.zeros(int(n_steps))
        self._ck(self.L.npsl_step_timing_list(self.h, int(n_steps), 1 if flush_l2 else 0, int(lead_in), _dp(out)))
        return out

    def force_calculation(self, pos, force_out: np.ndarray | None = None):
        """Host-buffer force_calculation: positions in, total forces out (npsl_force_calculation).
        ``pos`` / ``force_out`` must be C-contiguous float64 [n,3]; no copies are made here."""
        if pos is not None and (pos.dtype != np.float64 or not pos.flags.c_contiguous or pos.shape != (self.nv, 3)):
            raise ValueError("pos must be C-contiguous float64 [n,3]")
        if force_out is None:
            force_out = np.empty((self.nv, 3))
        self._ck(self.L.npsl_force_calculation(self.h, _dp(pos), _dp(force_out)))
        return force_out

    def md_step_host(self, pos: np.ndarray, vel: np.ndarray, force: np.ndarray, n_steps: int = 1) -> float:
        """Loop body of md_interface over host-resident state (npsl_md_step_host): the three
        C-contiguous float64 [n,3] arrays are updated in place; returns the kinetic energy."""
        for a in (pos, vel, force):
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != (self.nv, 3):
                raise ValueError("pos, vel, force must be C-contiguous float64 [n,3]")
        ke = C.c_double()
        self._ck(self.L.npsl_md_step_host(self.h, _dp(pos), _dp(vel), _dp(force), int(n_steps), C.byref(ke)))
        return ke.value

    def host_state(self):
        """The context's page-locked state block as numpy views (npsl_host_state): pos, vel, force [n,3] and the
        kinetic energy [1]. Valid until the context is closed."""
        ptr = [C.POINTER(C.c_double)() for _ in range(4)]
        self._ck(self.L.npsl_host_state(self.h, *[C.byref(p) for p in ptr]))
        views = [np.ctypeslib.as_array(p, shape=(self.nv, 3)) for p in ptr[:3]]
        return views[0], views[1], views[2], np.ctypeslib.as_array(ptr[3], shape=(1,))

    def md_step_packed(self, n_steps: int = 1) -> None:
        """Loop body of md_interface over the page-locked state block (npsl_md_step_packed)."""
        self._ck(self.L.npsl_md_step_packed(self.h, int(n_steps)))

    def pair_kernel_info(self) -> dict:
        r, t, f = C.c_int32(), C.c_int32(), C.c_double()
        self._ck(self.L.npsl_pair_kernel_info(self.h, C.byref(r), C.byref(t), C.byref(f)))
        return {"rows_per_thread": r.value, "threads": t.value, "fp64_instr_per_pair": f.value}

    def pair_math(self) -> int:
        """0 exact, 1 polynomial table, 2 reduced-accuracy exact scheme (npsl_pair_math)."""
        return self.L.npsl_pair_math(self.h)

    def owned_rows(self):
        lo, hi = C.c_int32(), C.c_int32()
        self._ck(self.L.npsl_owned_rows(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def pair_geometry(self) -> dict:
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.L.npsl_pair_geometry(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"rows_per_thread": a.value, "row_tiles": b.value, "col_splits": c.value, "cols_per_cta": d.value}

    def set_pair_mode(self, mode: int) -> None:
        """0 automatic, 1 ordered (k_pair), 2 symmetric (k_pair_sym)."""
        self._ck(self.L.npsl_set_pair_mode(self.h, int(mode)))

    def pair_mode(self) -> int:
        return int(self.L.npsl_pair_mode(self.h))

    def step_profile(self, n_steps: int = 20) -> str:
        buf = C.create_string_buffer(4096)
        self._ck(self.L.npsl_step_profile(self.h, int(n_steps), buf, 4096))
        return buf.value.decode()

    def set_volume_constraint(self, on: bool = True) -> None:
        """Rigid volume constraint (-G V): SHAKE / RATTLE inside every step."""
        self._ck(self.L.npsl_set_volume_constraint(self.h, 1 if on else 0))

    def constraint_init(self) -> None:
        """SHAKE + RATTLE before the first step (src/newmd.cpp:33-38)."""
        self._ck(self.L.npsl_constraint_init(self.h))

    def constraint_multipliers(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.npsl_constraint_multipliers(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_ions(self, pos, vel, q, diameter: float, lj_length_mesh_ions: float) -> None:
        """Explicit counterions (vector<PARTICLE>): pos [n,3], vel [n,3] or None, q [n]; call before force_init."""
        pos = _f64(pos)
        q = _f64(q)
        vel = None if vel is None else _f64(vel)
        self.n_ions = int(pos.shape[0]) if pos is not None else 0
        self._ck(self.L.npsl_set_ions(self.h, self.n_ions, _dp(pos), _dp(vel), _dp(q), float(diameter), float(lj_length_mesh_ions)))

    def download_ions(self) -> dict:
        n = getattr(self, "n_ions", 0)
        out = {k: np.zeros((n, 3)) for k in ("pos", "vel", "force")}
        self._ck(self.L.npsl_download_ions(self.h, _dp(out["pos"]), _dp(out["vel"]), _dp(out["force"])))
        return out

    def vertex_geometry(self) -> dict:
        """VERTEX::itsarea / itsnormal of every vertex at the positions on the device (src/vertex.cpp:7-18)."""
        area, normal = np.zeros(self.nv), np.zeros((self.nv, 3))
        self._ck(self.L.npsl_vertex_geometry(self.h, _dp(area), _dp(normal)))
        return {"area": area, "normal": normal}

    def local_energies(self) -> dict:
        """Per-face energy maps of INTERFACE::compute_local_energies(_by_component), src/interface.cpp:1140-1337."""
        out = {k: np.zeros(self.nf) for k in ("electrostatic", "elastic", "bending", "stretching")}
        self._ck(self.L.npsl_local_energies(self.h, *[_dp(out[k]) for k in ("electrostatic", "elastic", "bending", "stretching")]))
        return out

    def shard_plan(self) -> int:
        """0 single GPU, 1 rows (owned vertex slices), 2 replicated integrator (only the pair term is split)."""
        return int(self.L.npsl_shard_plan(self.h))

    def exchange_mode(self) -> int:
        """0 single GPU, 1 NCCL all-gathers, 2 stores into peer memory over NVLink."""
        return int(self.L.npsl_exchange_mode(self.h))

    def set_pair_plan(self, rows_per_thread: int, col_splits: int) -> None:
        self._ck(self.L.npsl_set_pair_plan(self.h, int(rows_per_thread), int(col_splits)))
