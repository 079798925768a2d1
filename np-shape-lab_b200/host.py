"""Host-side mirror of the reference's call surface for the force/energy path.

Same names, argument meaning and side effects as the reference's in-process C++ surface, with
the bodies replaced by calls over the C ABI (include/npsl.h -> csrc/):

    force_calculation_init(boundary, counterions, scalefactor, bucklingFlag)   src/newforces.cpp:4
    force_calculation(boundary, counterions, scalefactor, bucklingFlag)        src/newforces.cpp:137
    INTERFACE.compute_energy(num, counterions, scalefactor, bucklingFlag)      src/interface.cpp:964
    INTERFACE.dressup(lambda_a, lambda_v)                                      src/interface.cpp:33
    md_interface(boundary, counterions, mesh_bath, ions_bath, mdremote, geomConstraint,
                 bucklingFlag, constraintForm, scalefactor, box_radius)        src/newmd.cpp:9

Results are delivered the way the reference delivers them -- by mutating ``boundary`` (per-vertex
``forvec/bforce/sforce/TForce/VolTForce``, ``total_area``, ``penergy``, ...) and the thermostat
objects -- so tests written against the reference's semantics read the same here. The arrays
replace the reference's pointer-linked VERTEX/EDGE/FACE objects one field per array.

There is no CPU implementation behind these functions: without libnpsl.so and a CUDA device they
raise :class:`capi.NpslError`.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional

import numpy as np

from . import capi
from . import mesh as M


@dataclasses.dataclass
class THERMOSTAT:
    """src/thermostat.h:8-71."""
    Q: float
    T: float
    dof: float
    xi: float = 0.0
    eta: float = 0.0

    def row(self):
        return [self.Q, self.T, self.dof, self.xi, self.eta]


@dataclasses.dataclass
class CONTROL:
    """src/control.h:6-42 (the fields md_interface reads)."""
    timestep: float = 0.001
    steps: int = 1000
    writedata: int = 500
    moviefreq: int = 1000
    offfreq: int = 2500
    anneal: str = "y"
    annealfreq: int = 50000
    annealDuration: int = 5000
    TAnnealFac: float = 1.25
    QAnnealFac: float = 1.0 + 0.25 / 10.0


def make_baths(inp: M.PhysicalInputs, n_vertices: int, n_ions: int = 0):
    """src/main.cpp:276-290 as lists of THERMOSTAT."""
    mb, ib = M.make_thermostats(inp, n_vertices, n_ions)
    conv = lambda a: [THERMOSTAT(*[float(x) for x in r]) for r in a]
    return conv(mb), conv(ib)


class INTERFACE:
    """The triangulated membrane (src/interface.h:20-150), array-per-field."""

    def __init__(self):
        self.ein = self.eout = self.em = 0.0
        self.lB_out = self.inv_kappa_out = 0.0
        self.number_of_vertices = 0
        # VERTEX fields
        self.posvec = self.velvec = self.forvec = None
        self.q = None
        self.bforce = self.sforce = self.TForce = self.VolTForce = self.pairforce = None
        # EDGE / FACE fields
        self.edges = self.faces = None
        self.l0 = None
        self.itsS = None
        self.face_area = self.face_volume = self.face_direction = None
        # scalars
        self.total_area = self.total_volume = 0.0
        self.ref_area = self.ref_volume = 0.0
        self.avg_edge_length = 0.0
        self.lambda_a = self.lambda_v = 0.0
        self.bkappa = self.sconstant = self.sigma_a = self.sigma_v = 0.0
        self.elj = 1.0
        self.lj_length_mesh_mesh = 0.0
        self.lj_length_mesh_ions = 0.0  # (2 * 0.5 avg_edge + counterion diameter) / 2, src/main.cpp:271-272
        self.netMeshKE = self.netIonKE = self.penergy = self.energy = 0.0
        self.energy_parts = {}
        self.dt = 0.001
        self._ctx: Optional[capi.Context] = None
        self._ctx_key = None
        self._dev_pos_current = False

    # ---- set-up (host only, src/main.cpp:196-294)
    def set_up(self, unit_radius_sphere: float, ein: float = M.EPSILON_WATER, eout: float = M.EPSILON_WATER):
        """src/interface.cpp:20-29."""
        self.ein, self.eout = ein, eout
        self.em = 0.5 * (ein + eout)
        self.lB_out = (M.LB_WATER * M.EPSILON_WATER / eout) / unit_radius_sphere

    def discretize(self, mesh: M.Mesh):
        """src/interface.cpp:70-256 with the file already parsed (mesh.load_dat / mesh.icosphere)."""
        n = mesh.n_vertices
        self.number_of_vertices = n
        self.posvec = np.array(mesh.pos, dtype=np.float64)
        self.velvec = np.zeros((n, 3))
        self.forvec = np.zeros((n, 3))
        self.q = np.zeros(n)
        self.m = np.ones(n)
        self.edges = np.array(mesh.edges, dtype=np.int32)
        self.faces = np.array(mesh.faces, dtype=np.int32)
        self.l0 = M.edge_lengths(mesh)
        s = 0.0
        for v in self.l0:
            s += v
        self.avg_edge_length = s / mesh.n_edges
        self._ctx = None

    @classmethod
    def from_inputs(cls, mesh: M.Mesh, inp: M.PhysicalInputs) -> "INTERFACE":
        """Everything main() does before md_interface (src/main.cpp:196-294), deterministic flags."""
        b = cls()
        b.set_up(inp.R)
        b.discretize(mesh)
        p = M.derive_params(inp, mesh)
        b.inv_kappa_out = p["inv_kappa_out"]
        b.bkappa, b.sconstant, b.sigma_a, b.sigma_v = p["bkappa"], p["sconstant"], p["sigma_a"], p["sigma_v"]
        b.ref_area, b.ref_volume = p["ref_area"], p["ref_volume"]
        b.total_area, b.total_volume = p["ref_area"], p["ref_volume"]
        b.lj_length_mesh_mesh, b.elj = p["lj_length_mesh_mesh"], p["elj"]
        b.dt = inp.dt
        b.q = M.assign_charges(mesh, inp.q, inp.N, inp.p, 1.0, inp.F, inp.H)
        b.scalefactor = p["scalefactor"]
        return b

    # ---- device context, created lazily on first call
    def _params_dict(self, scalefactor: float, bucklingFlag: str) -> dict:
        return dict(scalefactor=scalefactor, em=self.em, inv_kappa_out=self.inv_kappa_out, elj=self.elj,
                    lj_length_mesh_mesh=self.lj_length_mesh_mesh, bkappa=self.bkappa, sconstant=self.sconstant,
                    sigma_a=self.sigma_a, sigma_v=self.sigma_v, ref_volume=self.ref_volume,
                    avg_edge_length=self.avg_edge_length, dt=self.dt, buckling=1 if bucklingFlag == "y" else 0)

    def context(self, scalefactor: float, bucklingFlag: str, mesh_bath=None, ions_bath=None) -> capi.Context:
        key = (float(scalefactor), bucklingFlag, self.bkappa, self.sconstant, self.sigma_a, self.sigma_v,
               self.ref_volume, self.inv_kappa_out, self.dt)
        if self._ctx is None or key != self._ctx_key:
            if self._ctx is not None:
                self._ctx.close()
            dummy = [[0.0, 0.0, 0.0, 0.0, 0.0]]
            mb = [t.row() for t in mesh_bath] if mesh_bath else dummy
            ib = [t.row() for t in ions_bath] if ions_bath else [[0.0] * 5 for _ in mb]
            P = capi.make_params(self._params_dict(scalefactor, bucklingFlag), mb, ib)
            self._ctx = capi.Context(self.edges, self.faces, self.l0, P, n_vertices=self.number_of_vertices)
            self._ctx_key = key
            self._ctx.upload_state(self.posvec, self.velvec, self.q)
        return self._ctx

    def push_state(self, ctx: capi.Context) -> None:
        ctx.upload_state(self.posvec, self.velvec, self.q)

    def pull_forces(self, ctx: capi.Context) -> None:
        st = ctx.download_state(pos=False, vel=False, force=True)
        self.forvec = st["force"]
        comp = ctx.download_force_components()
        self.pairforce, self.bforce, self.sforce = comp["pair"], comp["bending"], comp["stretching"]
        self.TForce, self.VolTForce = comp["tension"], comp["volume_tension"]
        mf = ctx.download_mesh_fields()
        self.itsS, self.face_area, self.face_volume = mf["edge_S"], mf["face_area"], mf["face_volume"]
        self.face_direction = mf["face_dir"]

    # ---- the reference's member functions on the path
    def dressup(self, lambda_a: float = 0.0, lambda_v: float = 0.0) -> None:
        """src/interface.cpp:33-67."""
        self.lambda_a, self.lambda_v = lambda_a, lambda_v
        ctx = self._ctx
        if ctx is None:
            raise capi.NpslError(-4, "dressup before the device context exists; call force_calculation_init first")
        ctx.upload_state(pos=self.posvec)
        self.total_area, self.total_volume = ctx.dressup()
        mf = ctx.download_mesh_fields()
        self.face_area, self.face_volume, self.face_direction = mf["face_area"], mf["face_volume"], mf["face_dir"]

    def compute_energy(self, num: int, counterions: list, scalefactor: float, bucklingFlag: str,
                       outdir: Optional[str] = None) -> None:
        """src/interface.cpp:964-1137; appends the ``energy_in_parts`` row when ``outdir`` is given."""
        ctx = self.context(scalefactor, bucklingFlag)  # the counterions are on the device since the last force call
        e = ctx.energy()
        self.netIonKE = e["ions_kinetic"]
        self.energy_parts = e
        self.netMeshKE, self.penergy, self.energy = e["kinetic"], e["potential"], e["energy"]
        self.total_area, self.total_volume = e["total_area"], e["total_volume"]
        if outdir is not None:
            with open(os.path.join(outdir, "energy_in_parts_kE_bE_sE_tE_ljE_esE.dat"), "a") as fh:
                fh.write("%d %g %g %g %g %g %g %g\n" % (num, e["kinetic"], e["bending"], e["stretching"], e["tension"],
                                                        e["lj"], e["electrostatic"], e["volume"]))


    def compute_vertex_area_normals(self) -> None:
        """VERTEX::compute_area_normal for every vertex (src/vertex.cpp:7-18; src/newmd.cpp:232-235)."""
        if self._ctx is None:
            raise capi.NpslError(-4, "compute_vertex_area_normals before the device context exists")
        geo = self._ctx.vertex_geometry()
        self.itsarea, self.itsnormal = geo["area"], geo["normal"]

    def compute_local_energies(self, scalefactor: float, outdir: Optional[str] = None) -> dict:
        """src/interface.cpp:1140-1232: per-face electrostatic and elastic energy maps; with ``outdir`` also the
        reference's ``local_electrostatic_E.off`` / ``local_elastic_E.off`` (colormap = (0, E, 1 - E, 1))."""
        return self._local_maps(("electrostatic", "elastic"), outdir)

    def compute_local_energies_by_component(self, outdir: Optional[str] = None) -> dict:
        """src/interface.cpp:1235-1337: the stretching and bending maps separately."""
        return self._local_maps(("stretching", "bending"), outdir)

    def _local_maps(self, keys, outdir) -> dict:
        if self._ctx is None:
            raise capi.NpslError(-4, "local energy maps before the first force call")
        loc = self._ctx.local_energies()
        pos = self._ctx.download_state(vel=False, force=False)["pos"]
        for k in keys:
            print("%s range\t%g\t%g" % (k, loc[k].min(), loc[k].max()))
            if outdir is not None:
                with open(os.path.join(outdir, "local_%s_E.off" % k), "w") as fh:
                    fh.write("OFF\n%d %d %d\n" % (pos.shape[0], self.faces.shape[0], self.edges.shape[0]))
                    for p in pos:
                        fh.write("%g %g %g\n" % tuple(p))
                    for f, e in zip(self.faces, loc[k]):
                        fh.write("3 %d %d %d 0 %g %g 1\n" % (f[0], f[1], f[2], e, 1 - e))
        return {k: loc[k] for k in keys}


class PARTICLE:
    """src/particle.h:13-91: one explicit counterion (mass 1)."""

    def __init__(self, id: int = 0, diameter: float = 0.0, valency: int = 0, q: float = 0.0, m: float = 1.0, posvec=(0.0, 0.0, 0.0)):
        self.id, self.diameter, self.valency, self.q, self.m = id, diameter, valency, q, m
        self.posvec = np.array(posvec, dtype=np.float64)
        self.velvec = np.zeros(3)
        self.forvec = np.zeros(3)


def push_ions(boundary: "INTERFACE", ctx: capi.Context, counterions: list) -> None:
    """vector<PARTICLE> -> device (positions, velocities, charges; ion diameter and INTERFACE::lj_length_mesh_ions)."""
    if not counterions:
        if getattr(ctx, "n_ions", 0):
            ctx.set_ions(None, None, None, 1.0, 1.0)
        return
    ctx.set_ions(np.array([p.posvec for p in counterions]), np.array([p.velvec for p in counterions]),
                 np.array([p.q for p in counterions]), counterions[0].diameter, boundary.lj_length_mesh_ions)


def pull_ions(ctx: capi.Context, counterions: list) -> None:
    if not counterions:
        return
    st = ctx.download_ions()
    for k, p in enumerate(counterions):
        p.posvec, p.velvec, p.forvec = st["pos"][k].copy(), st["vel"][k].copy(), st["force"][k].copy()


def force_calculation_init(boundary: INTERFACE, counterions: list, scalefactor: float, bucklingFlag: str) -> None:
    """src/newforces.cpp:4-135."""
    ctx = boundary.context(scalefactor, bucklingFlag)
    boundary.push_state(ctx)
    push_ions(boundary, ctx, counterions)
    ctx.force_init()
    boundary.pull_forces(ctx)
    pull_ions(ctx, counterions)


def force_calculation(boundary: INTERFACE, counterions: list, scalefactor: float, bucklingFlag: str) -> None:
    """src/newforces.cpp:137-289."""
    ctx = boundary.context(scalefactor, bucklingFlag)
    boundary.push_state(ctx)
    push_ions(boundary, ctx, counterions)
    ctx.force()
    boundary.pull_forces(ctx)
    pull_ions(ctx, counterions)


def compute_kinetic_energy(boundary: INTERFACE) -> float:
    """src/newenergies.h:8-16 (host arrays; the device computes its own inside the step)."""
    v = boundary.velvec
    return float(np.sum(0.5 * boundary.m * np.einsum("ij,ij->i", v, v)))


def bath_kinetic_energy(bath: List[THERMOSTAT]) -> float:
    return float(sum(0.5 * t.Q * t.xi * t.xi for t in bath))  # src/thermostat.h:67-70


def bath_potential_energy(bath: List[THERMOSTAT]) -> float:
    return float(sum(t.dof * M.KB * t.T * t.eta for t in bath))  # src/thermostat.h:61-64


def _sync_baths(ctx, mesh_bath, ions_bath) -> None:
    m, i = ctx.get_thermostat()
    for j, t in enumerate(mesh_bath):
        t.Q, t.T, t.dof, t.xi, t.eta = [float(x) for x in m[j]]
    for j, t in enumerate(ions_bath):
        t.Q, t.T, t.dof, t.xi, t.eta = [float(x) for x in i[j]]


def md_interface(boundary: INTERFACE, counterions: list, mesh_bath: List[THERMOSTAT], ions_bath: List[THERMOSTAT],
                 mdremote: CONTROL, geomConstraint: str = "N", bucklingFlag: str = "n", constraintForm: str = "L",
                 scalefactor: float = 0.0, box_radius: float = 1.0, outdir: Optional[str] = None,
                 on_write=None) -> None:
    """src/newmd.cpp:9-323. The loop body runs on the device (one CUDA-graph launch per step);
    the host touches the state only on write steps (``num < 3`` or ``num % writedata == 0``) and for
    annealing. Series files are written to ``outdir`` in the reference's formats."""
    if geomConstraint not in ("N", "V") or (geomConstraint == "V" and constraintForm != "L"):
        raise NotImplementedError("only the rigid volume constraint in its linear form (-G V) is on the device path")
    boundary.dt = mdremote.timestep
    boundary.velvec = np.zeros_like(boundary.posvec)  # initialize_velocities_to_zero, src/newmd.cpp:16
    boundary._ctx = None
    ctx = boundary.context(scalefactor, bucklingFlag, mesh_bath, ions_bath)
    if geomConstraint == "V":
        ctx.set_volume_constraint(True)
    for p in counterions or []:
        p.velvec = np.zeros(3)
    push_ions(boundary, ctx, counterions)
    ctx.force_init()  # src/newmd.cpp:19
    if geomConstraint == "V":
        ctx.constraint_init()  # SHAKE + RATTLE prior to MD, src/newmd.cpp:33-38
    n3 = 3.0 * boundary.number_of_vertices

    def write_row(num: int) -> dict:
        boundary.compute_energy(num, counterions, scalefactor, bucklingFlag, outdir)
        _sync_baths(ctx, mesh_bath, ions_bath)
        e = boundary.energy_parts
        ke, pe = e["kinetic"], e["potential"]
        bke = e["mesh_bath_ke"] + e["ions_bath_ke"]
        bpe = e["mesh_bath_pe"] + e["ions_bath_pe"]
        row = dict(num=num, KE=ke, PE=pe, E=ke + pe, conserved=ke + pe + bke + bpe, bathKE=bke, bathPE=bpe,
                   area=e["total_area"], volume=e["total_volume"], parts=dict(e))
        if outdir is not None:
            with open(os.path.join(outdir, "energy_nanomembrane.dat"), "a") as fh:
                fh.write("%15d%15g%15g%15g%15g%15g%15g\n" % (num, ke, pe, ke + pe, row["conserved"], bke, bpe))
            with open(os.path.join(outdir, "area.dat"), "a") as fh:
                fh.write("%d\t%g\n" % (num, e["total_area"]))
            with open(os.path.join(outdir, "volume.dat"), "a") as fh:
                fh.write("%d\t%g\n" % (num, e["total_volume"]))
            with open(os.path.join(outdir, "temperature.dat"), "a") as fh:
                fh.write("%d\t%g\n" % (num, 2.0 * ke / (n3 * M.KB)))
        if on_write is not None:
            on_write(row)
        return row

    if outdir is not None:
        os.makedirs(outdir, exist_ok=True)
    first = write_row(0)  # src/newmd.cpp:52-79
    e0 = first["KE"] + first["PE"] + first["parts"]["mesh_bath_ke"] + first["parts"]["mesh_bath_pe"]  # initNetEnergy
    c = mdremote
    fT = (1.0 / c.TAnnealFac) ** (1.0 / c.annealDuration)  # src/newmd.cpp:84-87
    fQ = (1.0 / c.QAnnealFac) ** (1.0 / c.annealDuration)
    abort_counter = 0

    def anneal_step(n: int) -> bool:  # the condition of src/newmd.cpp:258-260 without the T floor
        return c.anneal == "y" and n > c.annealDuration and (n % c.annealfreq == 0 or n % c.annealfreq < c.annealDuration)

    num = 0
    while num < c.steps:
        # run on the device up to the next step at which the host has something to do
        cand = [c.steps, num + 1 if num < 2 else (num // c.writedata + 1) * c.writedata]
        if c.anneal == "y":
            if anneal_step(num + 1):
                cand.append(num + 1)
            cand.append((num // c.annealfreq + 1) * c.annealfreq)
            if num < c.annealDuration + 1:
                cand.append(c.annealDuration + 1)
        nxt = min(cand)
        ctx.step(nxt - num)
        num = nxt
        if num % c.writedata == 0 or num < 3:  # src/newmd.cpp:174
            row = write_row(num)
            ratio = row["conserved"] / e0 if e0 != 0 else 1.0
            if outdir is not None:
                with open(os.path.join(outdir, "net_Energy_Drift.dat"), "a") as fh:
                    fh.write("%d\t%g\t%g\t%g\t%d\n" % (num, ratio, row["conserved"], e0, abort_counter))
            # src/newmd.cpp:205-220: only the downward drift before annealing counts; 10 strikes abort.
            if not (1.05 < ratio) and num < c.annealfreq and ratio < 0.95:
                abort_counter += 1
                if abort_counter == 10:
                    raise capi.NpslError(-5, "aborting: >5 %% net energy drift downwards prior to annealing (step %d)" % num)
        # gradual annealing, src/newmd.cpp:258-299
        if anneal_step(num):
            _sync_baths(ctx, mesh_bath, ions_bath)
            if mesh_bath[0].T >= 1e-15:
                if num % c.annealfreq == 0:
                    if num == c.annealfreq:
                        c.TAnnealFac = 1.25
                    elif num <= 3 * c.annealfreq:
                        c.TAnnealFac = 2.0
                    elif num <= 6 * c.annealfreq:
                        c.TAnnealFac = 4.0
                    elif num <= 7 * c.annealfreq:
                        c.TAnnealFac = 8.0
                    else:
                        c.TAnnealFac = 10.0
                    c.QAnnealFac = 1.0 + c.TAnnealFac / 10.0
                    fT = (1.0 / c.TAnnealFac) ** (1.0 / c.annealDuration)
                    fQ = (1.0 / c.QAnnealFac) ** (1.0 / c.annealDuration)
                for t in mesh_bath:
                    t.T = fT * t.T
                    t.Q = fQ * t.Q
                ctx.set_thermostat([t.row() for t in mesh_bath], [t.row() for t in ions_bath])
    st = ctx.download_state()
    boundary.posvec, boundary.velvec, boundary.forvec = st["pos"], st["vel"], st["force"]
    pull_ions(ctx, counterions)
    _sync_baths(ctx, mesh_bath, ions_bath)
