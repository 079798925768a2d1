"""The oracle (oracle/npsl_oracle.c, a CPU restatement) is pinned here against outputs of the
reference's own code: tests/golden/*.npz were produced by oracle/_ref/npsl_ref, i.e. the
unmodified reference sources (oracle/gen_golden.py), and tests/golden/reference_examples.json
holds the series files the reference ships under examples/ plus the README's final values."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, golden, make_oracle, relerr

VEC_KEYS = [("force", "force"), ("pforce", "pforce"), ("bforce", "bforce"), ("sforce", "sforce"),
            ("tforce", "tforce"), ("vforce", "vforce")]
E_MAP = {"KE": "netMeshKE", "bE": "bE", "sE": "sE", "tE": "tE", "ljE": "ljE", "esE": "esE", "volE": "volE",
         "penergy": "penergy", "energy": "energy"}


def check_step(o, g, k, tol=1e-12):
    fscale = float(np.max(np.abs(g["k%d_force" % k])))
    for ok, gk in VEC_KEYS:
        ref = g["k%d_%s" % (k, gk)]
        assert relerr(o.vec(ok), ref, floor=1e-6 * fscale) < tol, (g.name, k, ok)
    assert relerr(o.vec("pos"), g["k%d_pos" % k]) < tol
    assert relerr(o.vec("vel"), g["k%d_vel" % k], floor=1e-12) < 1e-10
    assert relerr(o.edge_S(), g["k%d_itsS" % k]) < tol
    e = o.energy()
    s = g.scalars(k)
    escale = abs(s["energy"])
    for ok, gk in E_MAP.items():
        assert abs(e[ok] - s[gk]) <= tol * max(escale, 1.0), (g.name, k, ok, e[ok], s[gk])
    sc = o.scalars()
    assert abs(sc["total_area"] - s["total_area"]) < tol * s["total_area"]
    assert abs(sc["total_volume"] - s["total_volume"]) < tol * s["total_volume"]
    be = o.bath_energies()
    for j, n in enumerate(["mesh_bath_ke", "mesh_bath_pe", "ions_bath_ke", "ions_bath_pe"]):
        assert abs(be[j] - s[n]) <= 1e-10 * max(abs(s[n]), 1e-12), (g.name, k, n, be[j], s[n])
    mb, ib = o.get_baths()
    assert relerr(mb, g["k%d_mesh_bath" % k]) < 1e-10
    assert relerr(ib, g["k%d_ions_bath" % k]) < 1e-10


def test_oracle_matches_reference_first_steps(gold):
    """Per-vertex forces (total and per term), EDGE::itsS, every energy component and the
    thermostat state at the steps the fixture holds in full (0, 1[, 2])."""
    o = make_oracle(gold)
    o.md_init()
    done = 0
    for k in gold.full_steps:
        o.md_steps(k - done)
        done = k
        check_step(o, gold, k)


@pytest.mark.parametrize("name", ["disc_D4", "buckled_D4"])
def test_oracle_trajectory_matches_reference(name):
    """A/U/E (+ every component) along 1000 steps, rel <= 1e-9 (the oracle is a different
    compilation of the same arithmetic; drift from re-association stays far below 1e-6)."""
    g = golden(name)
    o = make_oracle(g)
    o.md_init()
    done = 0
    for k in g.traj_steps:
        o.md_steps(k - done)
        done = k
        e, s = o.energy(), g.scalars(k)
        sc = o.scalars()
        assert abs(sc["total_area"] - s["total_area"]) < 1e-9 * s["total_area"]
        assert abs(e["penergy"] - s["penergy"]) < 1e-9 * abs(s["penergy"])
        assert abs(e["energy"] - s["energy"]) < 1e-9 * abs(s["energy"])
        assert abs(e["KE"] - s["netMeshKE"]) < 1e-7 * abs(s["energy"])
    assert relerr(o.vec("pos"), g["k%d_pos" % g.traj_steps[-1]]) < 1e-8


@pytest.mark.parametrize("name", ["disc_D4_GV", "janus_D8_GV"])
def test_oracle_volume_constraint_matches_reference(name):
    """SHAKE / RATTLE (-G V, src/functions.h:33-171) and the restated cubic against the reference's own
    constrained runs: first steps in full, then A/U/E and the pinned volume along the trajectory."""
    g = golden(name)
    o = make_oracle(g)
    o.set_constraint(True)
    o.md_init()
    o.constraint_init()
    done = 0
    for k in g.full_steps:
        o.md_steps(k - done)
        done = k
        check_step(o, g, k)
    for k in g.traj_steps[:1]:
        o.md_steps(k - done)
        done = k
        e, s = o.energy(), g.scalars(k)
        sc = o.scalars()
        assert abs(sc["total_area"] - s["total_area"]) < 1e-9 * s["total_area"]
        assert abs(e["penergy"] - s["penergy"]) < 1e-9 * abs(s["penergy"])
        assert abs(sc["total_volume"] - g.params["ref_volume"]) < 1e-12 * g.params["ref_volume"]


def test_golden_fixtures_reproduce_shipped_example_series():
    """The fixtures (compiled reference, %.17g) agree with the series files the reference ships in
    examples/*_Control at their printed 6 significant digits, rows 0,1,2 (and 500/1000 where the
    fixture holds them)."""
    ex = json.load(open(os.path.join(GOLD, "reference_examples.json")))
    for name, d in ex.items():
        g = golden(name)
        rows_e = {int(r[0]): r for r in d["energy_nanomembrane"]}
        rows_p = {int(r[0]): r for r in d["energy_in_parts"]}
        rows_a = {int(r[0]): r for r in d["area"]}
        for k in g.full_steps + g.traj_steps:
            if k not in rows_e:
                continue
            s = g.scalars(k)
            tol = 2e-5 if k <= 2 else 5e-5  # 6 printed digits; chaotic growth by step 500+ stays below this
            escale = abs(rows_e[k][3])

            def close(a, b, scale=None):
                return abs(a - b) <= tol * max(abs(b), scale if scale else 0.0)

            assert close(s["penergy"], rows_e[k][2]), (name, k)
            assert close(s["energy"], rows_e[k][3]), (name, k)
            cons = s["energy"] + s["mesh_bath_ke"] + s["mesh_bath_pe"] + s["ions_bath_ke"] + s["ions_bath_pe"]
            assert close(cons, rows_e[k][4]), (name, k)
            assert close(s["total_area"], rows_a[k][1]), (name, k)
            p = rows_p[k]
            for val, col in [(s["bE"], 2), (s["sE"], 3), (s["tE"], 4), (s["ljE"], 5), (s["esE"], 6), (s["volE"], 7)]:
                assert close(val, p[col], 1e-3 * escale), (name, k, col, val, p[col])
            if k >= 1 and abs(rows_e[k][1]) > 1e-300:
                assert close(s["netMeshKE"], rows_e[k][1], 1e-6 * escale), (name, k)


def test_oracle_row_slices_add_up():
    """The reference's MPI decomposition (src/main.cpp:448-455): pair forces of disjoint row slices
    concatenate to the full result and slice energies add up (all_gather / all_reduce semantics)."""
    from np_shape_lab_b200.mesh import row_partition
    g = golden("disc_D4")
    full = make_oracle(g)
    full.dressup()
    full.force()
    ef = full.energy()
    n = g.nv()
    for world in (2, 3, 8):
        es = lj = 0.0
        pf = np.zeros((n, 3))
        for r in range(world):
            lo, hi = row_partition(n, r, world)
            o = make_oracle(g)
            o.dressup()
            o.force(lo, hi - 1)
            pf[lo:hi] = o.vec("pforce")[lo:hi]
            e = o.energy(lo, hi - 1)
            es += e["esE"]
            lj += e["ljE"]
        assert relerr(pf, full.vec("pforce")) < 1e-15
        assert abs(es - ef["esE"]) < 1e-12 * abs(ef["esE"])
        assert lj == ef["ljE"] == 0.0


def make_ion_oracle(g):
    o = make_oracle(g)
    o.set_ions(g["ion_pos0"], None, g["ion_q"], float(g["ion_params"][0]), float(g["ion_params"][1]))
    return o


def test_oracle_counterions_match_reference():
    """Explicit counterions (mesh-ion and ion-ion loops of src/newforces.cpp:44-65,179-250, their energies
    src/interface.cpp:1059-1089, the ion half-kicks / drift and the ion Nose-Hoover chain driven by the ions' kinetic
    energy, src/newmd.cpp:99-170) against the compiled reference run with 64 hand-placed ions, some inside the
    mesh-ion WCA cutoff and two overlapping."""
    g = golden("disc_D4_ions")
    o = make_ion_oracle(g)
    o.md_init()
    done = 0
    for k in g.full_steps:
        o.md_steps(k - done)
        done = k
        check_step(o, g, k)
        fi = g["k%d_ion_force" % k]
        assert relerr(o.ions("force"), fi) < 1e-12
        assert relerr(o.ions("pos"), g["k%d_ion_pos" % k]) < 1e-13
        assert relerr(o.ions("vel"), g["k%d_ion_vel" % k], floor=1e-12) < 1e-12
        assert abs(o.ions_ke() - float(g["k%d_netIonKE" % k])) <= 1e-12 * max(float(g["k%d_netIonKE" % k]), 1e-9)
    for k in g.traj_steps:
        o.md_steps(k - done)
        done = k
        e, s = o.energy(), g.scalars(k)
        assert abs(e["penergy"] - s["penergy"]) < 1e-9 * abs(s["penergy"])
        assert abs(e["energy"] - s["energy"]) < 1e-9 * abs(s["energy"])
        assert abs(o.ions_ke() - float(g["k%d_netIonKE" % k])) <= 1e-8 * float(g["k%d_netIonKE" % k])
    assert relerr(o.ions("pos"), g["k%d_ion_pos" % g.traj_steps[-1]]) < 1e-8


# ---- the large synthetic meshes (BASELINE.json configs[3], configs[4]) ------------------------------------------
def _window_case(name):
    import os
    from conftest import GOLD
    from np_shape_lab_b200 import workloads as W
    z = np.load(os.path.join(GOLD, name + "_windows.npz"))
    w = W.build(str(z["workload"]))
    prm = dict(w.params)
    prm.update(zip([str(k) for k in z["param_names"]], [float(v) for v in z["params"]]))
    return z, w, prm


def _lcg(n, amp):
    st, mask = 0x9E3779B97F4A7C15, (1 << 64) - 1
    out = np.empty(3 * n)
    for k in range(3 * n):
        st = (st * 6364136223846793005 + 1442695040888963407) & mask
        out[k] = amp * (2.0 * ((st >> 11) * (1.0 / 9007199254740992.0)) - 1.0)
    return out.reshape(n, 3)


@pytest.mark.parametrize("name", ["S64_disc_perturbed", "S128_rod"])
def test_restatement_matches_reference_row_windows_on_large_meshes(name):
    """The C restatement against the compiled reference's force_calculation on the bench meshes: pair force on 32 rows
    of each of the three windows (the fixture holds 512 per window; the CUDA path is checked on all of them), mesh
    terms on the window rows, in the displaced state."""
    from oracle import oracle as O
    z, w, prm = _window_case(name)
    n = w.n_vertices
    pos = w.mesh.pos + _lcg(n, float(z["perturb"]))
    o = O.Oracle(w.mesh.edges, w.mesh.faces, w.params["l0"], {**prm, "T": w.params["T"], "dt": w.params["dt"], "buckling": 0})
    o.set_state(pos, np.zeros_like(pos), z["q"])
    o.dressup()
    rows = z["pert_rows"]
    pscale = float(np.max(np.abs(z["pert_pforce"])))
    for lo, hi in z["windows"]:
        lo, hi = int(lo), int(lo) + 31
        o.force(lo, hi)
        sel = np.where((rows >= lo) & (rows <= hi))[0]
        assert relerr(o.vec("pforce")[lo:hi + 1], z["pert_pforce"][sel]) < 1e-12
    for ok, gk in (("bforce", "bforce"), ("sforce", "sforce"), ("tforce", "tforce"), ("vforce", "vforce")):
        assert relerr(o.vec(ok)[rows], z["pert_" + gk], floor=1e-3 * pscale) < 1e-11, gk


def test_restatement_matches_reference_on_s64_disc_step_1():
    """S64_disc (the bench configuration) after one whole reference step: mesh terms, EDGE::itsS, face fields in full,
    pair force on three windows of 32 rows."""
    from conftest import golden
    from oracle import oracle as O
    g = golden("S64_disc")
    o = O.Oracle(g["edges"], g["faces"], g["l0"], g.params)
    o.set_state(g["k1_pos"], g["k1_vel"], g["q"])
    o.dressup()
    n = g.nv()
    fscale = float(np.max(np.abs(g["k1_force"])))
    for lo in (0, n // 2, n - 32):
        o.force(lo, lo + 31)
        assert relerr(o.vec("pforce")[lo:lo + 32], g["k1_pforce"][lo:lo + 32]) < 1e-12
    for k in ("bforce", "sforce", "tforce"):
        assert relerr(o.vec(k), g["k1_" + k], floor=1e-3 * fscale) < 1e-11, k
    assert relerr(o.vec("vforce"), g["k1_vforce"], floor=fscale) < 1e-11
    assert relerr(o.edge_S(), g["k1_itsS"]) < 1e-12
