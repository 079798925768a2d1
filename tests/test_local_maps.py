"""Data feeds of the reference's writers: vertex areas / normals
(VERTEX::compute_area_normal, src/vertex.cpp:7-18) and the per-face local energy maps of
INTERFACE::compute_local_energies(_by_component) (src/interface.cpp:1140-1337).

tests/golden/local_maps.npz (oracle/gen_golden.py) holds, from the compiled reference: the state after a few steps,
the maps in full precision, the g channel of the reference's own local_*_E.off files of the same run (6 digits), and
the local_*_E.off files the reference ships under examples/ (final configurations after 25 000 steps)."""
import os

import numpy as np
import pytest

from conftest import GOLD, golden, relerr

KEYS = ("electrostatic", "elastic", "bending", "stretching")
CASES = ["disc_D4", "janus_D8", "buckled_D4"]
EXAMPLES = ["disc_D4", "rod_D4", "janus_D4", "buckled_D4"]


@pytest.fixture(scope="module")
def maps():
    return np.load(os.path.join(GOLD, "local_maps.npz"))


def oracle_at(g, pos):
    from oracle import oracle as O
    o = O.Oracle(g["edges"], g["faces"], g["l0"], g.params)
    o.set_state(pos, np.zeros_like(pos), g["q"])
    o.set_baths(g["k0_mesh_bath"], g["k0_ions_bath"])
    o.dressup()
    o.force()
    return o


def face_es(pos, q, face, p):
    """sum over the face's vertices of their screened-Coulomb energy with every other vertex (src/newenergies.cpp:21-25)."""
    tot = 0.0
    for v in face:
        r = np.linalg.norm(pos - pos[v], axis=1)
        r[v] = np.inf
        tot += float(np.sum(p["scalefactor"] * q[v] * q * np.exp(-r / p["inv_kappa_out"]) / (p["em"] * r)))
    return tot


def floor_of(ref_all):
    """Absolute floor for maps that are identically ~0 (stretching at step 1, electrostatics of uncharged meshes)."""
    return 1e-6 * max(float(np.max(np.abs(ref_all))), 1e-30)


@pytest.mark.parametrize("name", CASES)
def test_oracle_local_maps_match_reference(name, maps):
    g = golden(name)
    o = oracle_at(g, maps[name + "_pos"])
    geo = o.vertex_geometry()
    assert relerr(geo["area"], maps[name + "_vertex_area"]) < 1e-13
    assert relerr(geo["normal"], maps[name + "_vertex_normal"]) < 1e-13
    loc = o.local_energies()
    ref = maps[name + "_local_faces"]
    for k, key in enumerate(KEYS):
        # stretching is quadratic in (l - l0) with (l - l0) / l ~ 1e-5 after one or two steps: the double-rounded
        # positions of the fixture alone move it by ~1e-11 relative
        assert relerr(loc[key], ref[:, k], floor=floor_of(ref)) < (1e-9 if key == "stretching" else 1e-12), (name, key)
        off = maps["%s_off_%s" % (name, key)]  # the reference's own file of the same run, 6 significant digits
        assert np.max(np.abs(loc[key] - off)) <= 1e-5 * max(float(np.max(np.abs(off))), 1e-300), (name, key)


@pytest.mark.parametrize("name", EXAMPLES)
def test_oracle_reproduces_shipped_local_off_files(name, maps):
    """examples/*/local_*_E.off: positions (6 digits) and per-face energies of the final configuration."""
    g = golden(name)
    pos = maps["example_%s_pos" % name]
    o = oracle_at(g, pos)
    loc = o.local_energies()
    for key in KEYS:
        ref = maps["example_%s_%s" % (name, key)]
        scale = max(float(np.max(np.abs(ref))), 1e-300)
        # positions are rounded to 6 digits in the file: edge lengths move by ~1e-5 relative, the stretching map,
        # quadratic in (l - l0) with l - l0 ~ 1e-2 l, by a few 1e-3 of its scale, and 1 - cos(dihedral) of the bending
        # map by up to ~5e-4 of its scale
        tol = {"stretching": 2e-2, "elastic": 2e-2, "bending": 2e-3, "electrostatic": 2e-4}[key]
        assert np.max(np.abs(loc[key] - ref)) <= tol * scale, (name, key, np.max(np.abs(loc[key] - ref)) / scale)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_local_maps_match_reference(name, maps):
    from np_shape_lab_b200 import capi
    g = golden(name)
    P = capi.make_params(g.params, g["k0_mesh_bath"], g["k0_ions_bath"])
    pos = maps[name + "_pos"]
    with capi.Context(g["edges"], g["faces"], g["l0"], P, n_vertices=g.nv()) as ctx:
        ctx.upload_state(pos, np.zeros_like(pos), g["q"])
        ctx.force_init()
        geo = ctx.vertex_geometry()
        assert relerr(geo["area"], maps[name + "_vertex_area"]) < 1e-12
        assert relerr(geo["normal"], maps[name + "_vertex_normal"]) < 1e-12
        loc = ctx.local_energies()
        ref = maps[name + "_local_faces"]
        for k, key in enumerate(KEYS):
            assert relerr(loc[key], ref[:, k], floor=floor_of(ref)) < (1e-9 if key == "stretching" else 1e-10), (name, key)
        # the maps follow the positions: after a step they are recomputed, and they add up to the global energies
        ctx.step(3)
        loc = ctx.local_energies()
        e = ctx.energy()
        assert abs(loc["bending"].sum() / 2.0 - e["bending"]) <= 1e-10 * abs(e["bending"])  # every edge feeds two faces
        st = ctx.download_state()
        for f in (0, g["faces"].shape[0] - 1):
            assert abs(loc["electrostatic"][f] - face_es(st["pos"], g["q"], g["faces"][f], g.params)) <= 1e-10 * max(
                float(np.max(np.abs(loc["electrostatic"]))), 1e-9)


@pytest.mark.gpu
def test_gpu_local_maps_follow_the_symmetric_kernel_path():
    """Meshes above 4096 vertices take forces from k_pair_sym; the maps still come from the ordered energy pass."""
    from np_shape_lab_b200 import capi, workloads as W
    w = W.build("S32_disc")
    P = capi.make_params(w.params, w.mesh_bath, w.ions_bath)
    with capi.Context(w.mesh.edges, w.mesh.faces, w.params["l0"], P, n_vertices=w.n_vertices) as ctx:
        ctx.upload_state(w.mesh.pos, np.zeros_like(w.mesh.pos), w.q)
        ctx.force_init()
        assert ctx.pair_mode() == 2
        ctx.step(2)
        loc = ctx.local_energies()
        e = ctx.energy()
        assert abs(loc["bending"].sum() / 2.0 - e["bending"]) <= 1e-11 * abs(e["bending"])
        st = ctx.download_state()
        for f in (0, 5000, w.mesh.faces.shape[0] - 1):
            ref = face_es(st["pos"], w.q, w.mesh.faces[f], w.params)
            assert abs(loc["electrostatic"][f] - ref) <= 1e-11 * abs(ref)
        geo = ctx.vertex_geometry()
        assert abs(geo["area"].sum() - e["total_area"]) <= 1e-12 * e["total_area"]
        assert np.allclose(np.linalg.norm(geo["normal"], axis=1), 1.0, atol=1e-14)
