"""Named workloads of BASELINE.json (``configs``) as inputs to the C ABI.

Host-side set-up only (the part of ``main`` before ``md_interface``, src/main.cpp:196-294): mesh,
reduced-unit parameters, the deterministic (``-F n``) charge assignment and the thermostat chains.

Time step of the synthetic meshes. The reference's default ``-d 0.001`` is tuned to its shipped
(3,D) meshes (edge ~0.2). The bending stiffness seen by one unit-mass vertex grows like
``bkappa / edge^2``, i.e. 100x from D4 to the class-I D=64 icosphere and 400x to D=128, which puts
dt = 1e-3 beyond the stability limit of velocity Verlet there (measured on the B200: NaN within
100 steps at D=64; stable with |dE/E| ~ 1e-5 over 1000 steps at 3e-4 and below, and at 6e-5 and
below for D=128). The synthetic workloads therefore pass ``-d 2e-4`` (D=64) and ``-d 5e-5``
(D=128); everything else is the README disc / rod recipe. Throughput does not depend on dt.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import mesh as M

@dataclasses.dataclass
class Workload:
    name: str
    mesh: M.Mesh
    inp: M.PhysicalInputs
    params: dict          # keys of npsl_params (+ l0, ref_area, T)
    q: np.ndarray         # [N]
    mesh_bath: np.ndarray  # [C,5]
    ions_bath: np.ndarray  # [C,5]
    cli: str              # the equivalent reference command line (for the record)
    D: int                # what the reference's -D would be

    @property
    def n_vertices(self) -> int:
        return self.mesh.n_vertices

    @property
    def ordered_pairs(self) -> int:
        """Pair evaluations per force_calculation call, counted as the reference evaluates them
        (ordered, i != j; src/newforces.cpp:156-157)."""
        n = self.mesh.n_vertices
        return n * (n - 1)


# name -> (mesh kind, D, PhysicalInputs overrides, CLI string)
_TABLE = {
    # BASELINE.json configs[3]: the configuration the headline metric is quoted on
    "S64_disc": ("ico", 64, dict(q=600.0, c=0.005, dt=2e-4),
                 "-R 10 -q 600 -c 0.005 -t 1 -v 1 -b 40 -s 40 -F n -d 0.0002 -D 64 (class-I icosphere as infiles/3_64.dat)"),
    # BASELINE.json configs[4]
    "S128_rod": ("ico", 128, dict(q=600.0, c=0.01, dt=5e-5),
                 "-R 10 -q 600 -c 0.01 -t 1 -v 1 -b 40 -s 40 -F n -d 0.00005 -D 128 (class-I icosphere as infiles/3_128.dat)"),
    # smaller synthetic meshes (smoke tests, CPU-affordable parity)
    "S32_disc": ("ico", 32, dict(q=600.0, c=0.005, dt=1e-3),
                 "-R 10 -q 600 -c 0.005 -t 1 -v 1 -b 40 -s 40 -F n -D 32 (class-I icosphere)"),
    "S8_disc": ("ico", 8, dict(q=600.0, c=0.005, dt=1e-3),
                "-R 10 -q 600 -c 0.005 -t 1 -v 1 -b 40 -s 40 -F n -D 8 (class-I icosphere)"),
}


def names():
    return list(_TABLE)


def build(name: str, mesh: Optional[M.Mesh] = None) -> Workload:
    kind, D, over, cli = _TABLE[name]
    if mesh is None:
        mesh = M.icosphere(D)
    inp = M.PhysicalInputs(**over)
    p = M.derive_params(inp, mesh)
    q = M.assign_charges(mesh, inp.q, inp.N, inp.p)
    mb, ib = M.make_thermostats(inp, mesh.n_vertices)
    return Workload(name=name, mesh=mesh, inp=inp, params=p, q=q, mesh_bath=mb, ions_bath=ib, cli=cli, D=D)


def reference_flags(w: Workload) -> list:
    """The workload as a command line in the reference's own CLI letters (src/main.cpp:109-185)."""
    i = w.inp
    return ["-R", i.R, "-q", i.q, "-c", i.c, "-t", i.t, "-v", i.v, "-b", i.b, "-s", i.s, "-B", i.B, "-F", "n",
            "-N", i.N, "-p", i.p, "-d", i.dt, "-T", i.T, "-Q", i.Q_mesh, "-Y", i.Q_ions, "-C", i.chain, "-D", w.D]
