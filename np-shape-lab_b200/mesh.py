"""Host-side mesh set-up for the force/energy path.

Everything here is one-off set-up that happens before the time loop (reference:
``src/main.cpp:196-294``): reading / writing the reference's ``infiles/<h>_<k>.dat`` mesh
format, building synthetic class-I icospheres in that format, the deterministic (``-F n``)
charge assignment, and the derived reduced-unit parameters handed over the C ABI.
None of it is on the per-step hot path; the per-step work lives in ``csrc/``.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Tuple

import numpy as np

# physical constants of the reference, src/utility.h:35-49
DCUT2 = 1.25992104989487316476
LB_WATER = 0.714
EPSILON_WATER = 80.0
ROOM_TEMPERATURE = 298.0
UNITENERGY = 1.3807 * math.pow(10.0, -16) * ROOM_TEMPERATURE
KB = 1.0


@dataclasses.dataclass
class Mesh:
    """Triangle mesh in the reference's file order.

    ``faces`` hold the winding the reference *uses* (src/interface.cpp:121-129: a face whose
    sign column is -1 is stored reversed), not the raw file columns.
    """

    pos: np.ndarray    # [N,3] float64
    edges: np.ndarray  # [E,2] int32, file order of the two end points
    faces: np.ndarray  # [F,3] int32, stored winding
    file_q: np.ndarray | None = None  # [N] fifth column of the vertex records (VERTEX::q until a pattern is assigned)

    @property
    def n_vertices(self) -> int:
        return int(self.pos.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.edges.shape[0])

    @property
    def n_faces(self) -> int:
        return int(self.faces.shape[0])


# --------------------------------------------------------------------------- file format


def load_dat(path: str) -> Mesh:
    """Parse ``infiles/<h>_<k>.dat`` the way ``INTERFACE::discretize`` does
    (src/interface.cpp:83-143): three sections, each opened by six header tokens
    (``# Number of X= n <name>``), whitespace separated records."""
    with open(path, "r") as fh:
        tok = fh.read().split()
    p = 0

    def header() -> int:
        nonlocal p
        n = int(tok[p + 4])
        p += 6
        return n

    nv = header()
    pos = np.zeros((nv, 3), dtype=np.float64)
    file_q = np.zeros(nv, dtype=np.float64)
    for _ in range(nv):
        idx = int(tok[p])
        pos[idx] = (float(tok[p + 1]), float(tok[p + 2]), float(tok[p + 3]))
        file_q[idx] = float(tok[p + 4])
        p += 5
    ne = header()
    edges = np.zeros((ne, 2), dtype=np.int32)
    for _ in range(ne):
        idx = int(tok[p])
        edges[idx] = (int(tok[p + 1]), int(tok[p + 2]))
        p += 4
    nf = header()
    faces = np.zeros((nf, 3), dtype=np.int32)
    for _ in range(nf):
        idx = int(tok[p])
        a, b, c, sign = int(tok[p + 1]), int(tok[p + 2]), int(tok[p + 3]), int(tok[p + 4])
        faces[idx] = (a, b, c) if sign == 1 else (c, b, a)
        p += 6
    return Mesh(pos, edges, faces, file_q)


def write_dat(path: str, mesh: Mesh) -> None:
    """Write a mesh in the reference's ``.dat`` layout (full ``%.17g`` coordinates, sign
    column always +1 because ``mesh.faces`` already holds the stored winding)."""
    with open(path, "w") as fh:
        fh.write("# Number of Vertices= %d\n\nvertices\n\n" % mesh.n_vertices)
        for i, (x, y, z) in enumerate(mesh.pos):
            fh.write("%d\t%.17g\t%.17g\t%.17g\t%.17g\n" % (i, x, y, z, 1.0 if mesh.file_q is None else mesh.file_q[i]))
        fh.write("\n\n\n# Number of Edges= %d\n\nedges\n\n" % mesh.n_edges)
        for i, (a, b) in enumerate(mesh.edges):
            fh.write("%d\t%d\t%d\t1\n" % (i, a, b))
        fh.write("\n\n\n# Number of Triangles= %d\n\ntriangles\n\n" % mesh.n_faces)
        for i, (a, b, c) in enumerate(mesh.faces):
            fh.write("%d\t%d\t%d\t%d\t1\t1\n" % (i, a, b, c))


# --------------------------------------------------------------------------- synthetic meshes


def _icosahedron() -> Tuple[np.ndarray, np.ndarray]:
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
         [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
         [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
         [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def icosphere(freq: int) -> Mesh:
    """Class-I geodesic icosphere of frequency ``freq``: 10f^2+2 vertices, 30f^2 edges, 20f^2
    faces on the unit sphere, outward counter-clockwise winding. No RNG. Shared vertices on
    icosahedron edges/corners are merged by their exact integer barycentric weights, so the
    result is bit-reproducible."""
    if freq < 1:
        raise ValueError("freq must be >= 1")
    iv, ifc = _icosahedron()
    D = freq
    # local lattice of one icosahedron face: (i,j) with i+j<=D, weights (D-i-j, i, j)
    ii, jj = np.meshgrid(np.arange(D + 1), np.arange(D + 1), indexing="ij")
    keep = (ii + jj) <= D
    li, lj = ii[keep], jj[keep]
    local_id = -np.ones((D + 1, D + 1), dtype=np.int64)
    local_id[li, lj] = np.arange(li.size)
    npts = li.size
    W = np.zeros((20 * npts, 12), dtype=np.int32)
    for f, (a, b, c) in enumerate(ifc):
        rows = slice(f * npts, (f + 1) * npts)
        W[rows, a] += (D - li - lj)
        W[rows, b] += li
        W[rows, c] += lj
    uniq, inverse = np.unique(W, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    pos = uniq.astype(np.float64) @ iv
    pos /= np.linalg.norm(pos, axis=1)[:, None]
    # small triangles of one face
    up_i, up_j = np.nonzero((ii + jj) <= D - 1)
    up = np.stack([local_id[up_i, up_j], local_id[up_i + 1, up_j], local_id[up_i, up_j + 1]], axis=1)
    dn_i, dn_j = np.nonzero((ii + jj) <= D - 2)
    dn = np.stack([local_id[dn_i + 1, dn_j], local_id[dn_i + 1, dn_j + 1], local_id[dn_i, dn_j + 1]], axis=1)
    tri_local = np.concatenate([up, dn], axis=0)
    faces = np.concatenate([inverse[f * npts + tri_local] for f in range(20)], axis=0)
    # outward CCW
    a, b, c = pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - b), a + b + c) < 0
    faces[flip] = faces[flip][:, ::-1]
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    e = np.unique(np.sort(e, axis=1), axis=0)
    assert pos.shape[0] == 10 * D * D + 2 and e.shape[0] == 30 * D * D and faces.shape[0] == 20 * D * D
    return Mesh(np.ascontiguousarray(pos), e.astype(np.int32), faces.astype(np.int32))


# --------------------------------------------------------------------------- initial geometry


def edge_lengths(mesh: Mesh, pos: np.ndarray | None = None) -> np.ndarray:
    """``EDGE::length`` for every edge (src/edge.cpp:24-31)."""
    p = mesh.pos if pos is None else pos
    d = p[mesh.edges[:, 1]] - p[mesh.edges[:, 0]]
    return np.sqrt(np.einsum("ij,ij->i", d, d))


def face_areas(mesh: Mesh, pos: np.ndarray | None = None) -> np.ndarray:
    """``FACE::computearea`` (src/face.cpp:28-36)."""
    p = mesh.pos if pos is None else pos
    v0, v1, v2 = p[mesh.faces[:, 0]], p[mesh.faces[:, 1]], p[mesh.faces[:, 2]]
    return 0.5 * np.linalg.norm(np.cross(v1 - v0, v2 - v1), axis=1)


def face_volumes(mesh: Mesh, pos: np.ndarray | None = None) -> np.ndarray:
    """``FACE::computevolume`` (src/face.cpp:38-45)."""
    p = mesh.pos if pos is None else pos
    v0, v1, v2 = p[mesh.faces[:, 0]], p[mesh.faces[:, 1]], p[mesh.faces[:, 2]]
    return np.einsum("ij,ij->i", v0, np.cross(v1, v2)) / 6.0


def vertex_areas(mesh: Mesh) -> np.ndarray:
    """``VERTEX::computearea``: one third of the incident face areas (src/vertex.cpp:21-27)."""
    fa = face_areas(mesh)
    va = np.zeros(mesh.n_vertices)
    for k in range(3):
        np.add.at(va, mesh.faces[:, k], fa)
    return va / 3.0


class GlibcRand:
    """glibc's ``srand`` / ``rand`` (TYPE_3 additive feedback generator, r[i] = r[i-3] + r[i-31]): the stream behind the
    reference's ``srand(123456); std::random_shuffle(...)`` (src/interface.cpp:367-371) on the platforms it is built on."""

    def __init__(self, seed: int):
        r = [0] * 34
        r[0] = seed & 0xFFFFFFFF if seed else 1
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1] if r[i - 1] < 2 ** 31 else r[i - 1] - 2 ** 32, 127773)
            word = 16807 * lo - 2836 * hi
            if word < 0:
                word += 2147483647
            r[i] = word
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = r
        for _ in range(310):
            self._next()

    def _next(self) -> int:
        v = (self.r[-31] + self.r[-3]) & 0xFFFFFFFF
        self.r.append(v)
        self.r.pop(0)
        return v

    def rand(self) -> int:
        return self._next() >> 1


def random_shuffle(seq: list, rng: GlibcRand) -> None:
    """libstdc++'s two-iterator ``std::random_shuffle``: swap element i with element rand() % (i + 1), i = 1..n-1."""
    for i in range(1, len(seq)):
        j = rng.rand() % (i + 1)
        if i != j:
            seq[i], seq[j] = seq[j], seq[i]


def assign_charges(mesh: Mesh, q_strength: float, num_divisions: int = 1, frac_patch: float = 0.5, alpha: float = 1.0,
                   random_flag: str = "n", function_flag: str = "n") -> np.ndarray:
    """``INTERFACE::assign_random_q_values`` (src/interface.cpp:339-798): uniform (``-N 1``), Janus (``-N 2``, with the
    yin-yang variant ``-H y``), the cube pattern (``-N 1 -H c``), z-stripes (``-N 3..17``), fractional occupancy
    ``alpha`` and the shuffled variants (``-F y``: ``srand(123456)`` + ``std::random_shuffle`` of the occupancy and area
    lists, reproduced for glibc / libstdc++).

    Vertices are ranked by (z, index); the vertex of rank i receives the area share of list entry i (vertex i unless
    shuffled -- the reference's deliberate index mismatch, src/interface.cpp:361-364,392); the total is then rescaled
    to an integer number of charges (src/interface.cpp:776-798)."""
    n = mesh.n_vertices
    if not 1 <= num_divisions <= 17:
        raise ValueError("num_divisions must be in [1, 17] (src/interface.cpp:343)")
    if q_strength == 0:
        return np.zeros(n)
    perm = np.lexsort((np.arange(n), mesh.pos[:, 2]))
    total_area = float(np.sum(face_areas(mesh)))
    state = [1] * int(alpha * n)
    state += [0] * (n - len(state))
    area = [float(a) for a in vertex_areas(mesh)]
    if random_flag == "y":
        rng = GlibcRand(123456)
        random_shuffle(state, rng)
        random_shuffle(area, rng)
    share = q_strength * np.array(area) / total_area
    pos = mesh.pos[perm]
    q_ranked = np.zeros(n)
    N = num_divisions
    if N == 1:
        if function_flag == "c":
            x2, y2, z2 = pos[:, 0] * pos[:, 0], pos[:, 1] * pos[:, 1], pos[:, 2] * pos[:, 2]
            on = (x2 + y2 <= 0.25) | (y2 + z2 <= 0.25) | (x2 + z2 <= 0.25)
        else:
            on = np.array(state) == 1
        q_ranked = np.where(on, share, 0.0)
    elif N == 2:
        n_patch = int(frac_patch * n)
        q_ranked[:n_patch] = share[:n_patch]
        if function_flag == "y":
            x, z = pos[:, 0], pos[:, 2]
            off = (z <= 0) & ((x - 0.5) * (x - 0.5) + z * z <= 0.25)
            on = (z >= 0) & ((x + 0.5) * (x + 0.5) + z * z <= 0.25)
            q_ranked = np.where(off, 0.0, q_ranked)
            q_ranked = np.where(on, share, q_ranked)
    else:  # stripes about the z axis: bands k = 0..N-1 of the ranking, even bands charged
        step = alpha / N
        i = 0
        for k in range(1, N):
            bound = (n * step) if k == 1 else float(n * k) * step  # "i < number_of_vertices * k*(alpha / num_divisions)"
            while i < bound and i < n:
                q_ranked[i] = share[i] if k % 2 == 1 else 0.0
                i += 1
        while i < n:
            q_ranked[i] = share[i] if N % 2 == 1 else 0.0
            i += 1
    q = np.zeros(n)
    q[perm] = q_ranked
    q_target = c_round(q_target_raw(q_strength, N, frac_patch))
    q_actual = 0.0
    for v in q:  # same left-to-right double accumulation as the reference
        q_actual += v
    return q * (q_target / q_actual)


def c_round(x: float) -> float:
    """C's round(): half away from zero (Python's round() goes to even)."""
    return math.copysign(math.floor(abs(x) + 0.5), x)


def q_target_raw(q_strength: float, num_divisions: int, frac_patch: float) -> float:
    """Argument of the reference's ``round`` (C round: half away from zero), src/interface.cpp:788-792."""
    if num_divisions == 1:
        return q_strength
    if num_divisions == 2:
        return frac_patch * q_strength
    return math.ceil(0.5 * num_divisions) * frac_patch * q_strength


def assign_external_charges(mesh: Mesh, q_strength: float, pattern_file: str) -> np.ndarray:
    """``INTERFACE::assign_external_q_values`` (src/interface.cpp:818-862): ``index value`` rows of
    ``infiles/Charge_Patterns/<name>.dat`` scaled by q_strength / total_area, then rescaled to an integer total."""
    n = mesh.n_vertices
    if q_strength == 0 or not os.path.isfile(pattern_file):  # "Mesh will be uncharged."
        return np.zeros(n)
    # vertices the file does not list keep the charge column of the mesh file (src/interface.cpp:94)
    q = np.ones(n) if mesh.file_q is None else np.array(mesh.file_q, dtype=np.float64)
    total_area = float(np.sum(face_areas(mesh)))
    tok = open(pattern_file).read().split()
    for k in range(0, len(tok) - 1, 2):
        q[int(tok[k])] = (q_strength / total_area) * float(tok[k + 1])
    q_actual = 0.0
    for v in q:
        q_actual += v
    return q * (c_round(q_actual) / q_actual)


# --------------------------------------------------------------------------- derived parameters


@dataclasses.dataclass
class PhysicalInputs:
    """The CLI inputs of the reference that reach the path (src/main.cpp:109-185)."""

    R: float = 10.0          # -R unit radius (nm)
    q: float = 600.0         # -q net charge
    c: float = 0.005         # -c salt concentration (M)
    t: float = 1.0           # -t surface tension (dyn/cm)
    v: float = 1.0           # -v volume tension (atm/nm^3)
    b: float = 40.0          # -b bending modulus
    s: float = 40.0          # -s stretching modulus
    B: str = "n"             # -B buckling form of the stretching term
    N: int = 1               # -N number of patches
    p: float = 0.5           # -p patch fraction
    F: str = "n"             # -F shuffle the occupancy / area lists (the reference's CLI default is 'y')
    H: str = "n"             # -H pattern variant: 'y' yin-yang (with -N 2), 'c' cube (with -N 1)
    dt: float = 0.001        # -d
    T: float = 0.001         # -T
    Q_mesh: float = 0.01     # -Q
    Q_ions: float = 0.01     # -Y
    chain: int = 5           # -C


def derive_params(inp: PhysicalInputs, mesh: Mesh) -> dict:
    """Reduced-unit parameters exactly as ``main`` derives them (src/main.cpp:196-273) plus the
    initial-geometry references (ref_volume, avg_edge_length, l0). Returns a plain dict whose
    keys match ``npsl_params`` in include/npsl.h."""
    R = inp.R
    sigma_a = (math.pow(10, -14) * (math.pow(R, 2) / UNITENERGY)) * inp.t
    sigma_v = ((1.01325 * math.pow(10, 5)) * math.pow(10, -27) * (math.pow(R, 6) / (UNITENERGY * math.pow(10, -7)))) * inp.v
    scalefactor = EPSILON_WATER * LB_WATER / R
    ein = eout = EPSILON_WATER
    em = 0.5 * (ein + eout)
    lB_out = (LB_WATER * EPSILON_WATER / eout) / R
    if inp.c == 0:
        inv_kappa_out = 100000000.0
    else:
        inv_kappa_out = (0.257 / math.sqrt(1.0 * lB_out * R * inp.c)) / R
    l0 = edge_lengths(mesh)
    avg = 0.0
    for v in l0:
        avg += v
    avg /= mesh.n_edges
    fa, fv = face_areas(mesh), face_volumes(mesh)
    ref_area = 0.0
    for v in fa:
        ref_area += v
    ref_volume = 0.0
    for v in fv:
        ref_volume += v
    return dict(
        scalefactor=scalefactor, em=em, inv_kappa_out=inv_kappa_out, elj=1.0,
        lj_length_mesh_mesh=0.1 * avg, bkappa=inp.b, sconstant=inp.s, sigma_a=sigma_a, sigma_v=sigma_v,
        ref_area=ref_area, ref_volume=ref_volume, avg_edge_length=avg, buckling=1 if inp.B == "y" else 0,
        dt=inp.dt, T=inp.T, l0=l0,
    )


def make_thermostats(inp: PhysicalInputs, n_vertices: int, n_ions: int = 0):
    """Nosé-Hoover chain set-up of src/main.cpp:276-290. Returns (mesh_bath, ions_bath), each an
    array [C,5] of (Q, T, dof, xi, eta); the last link is the Q=0 dummy."""

    def chain(Q, dof0):
        C = inp.chain
        if C == 1:
            return np.array([[0.0, inp.T, dof0, 0.0, 0.0]])
        rows = [[Q, inp.T, dof0, 0.0, 0.0]]
        while len(rows) != C - 1:
            rows.append([Q, inp.T, 1.0, 0.0, 0.0])
        rows.append([0.0, inp.T, dof0, 0.0, 0.0])
        return np.array(rows, dtype=np.float64)

    return chain(inp.Q_mesh, 3.0 * n_vertices), chain(inp.Q_ions, 3.0 * n_ions)


def row_partition(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """The reference's MPI row decomposition (src/main.cpp:448-455): ``range = ceil(N/P)``,
    rank r owns [r*range, min((r+1)*range, N)). Returns the half-open interval."""
    rng = (n_rows + world - 1) // world
    lo = min(rank * rng, n_rows)
    hi = min((rank + 1) * rng, n_rows)
    return lo, hi


def shard_range(n_rows: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Rows / vertices owned by one GPU rank of a sharded context (csrc/npsl.cu, create_impl): the
    reference's ``ceil(N/P)`` rounded up to a whole warp of vertices, so that the per-warp
    kinetic-energy partials -- and with them the trajectory -- are bit-identical to the single-GPU
    run. Half-open interval; the last rank takes the remainder."""
    rng = (n_rows + world - 1) // world
    rng = (rng + align - 1) // align * align
    return min(rank * rng, n_rows), min((rank + 1) * rng, n_rows)
