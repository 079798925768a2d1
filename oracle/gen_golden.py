"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz from the compiled reference.

Run in the development container (needs /root/reference and ``make -C oracle ref``):

    python oracle/gen_golden.py [--only NAME] [--threads T]

Every fixture is produced by the reference's *own* md_interface / force_calculation /
compute_energy (oracle/_ref/npsl_ref, unmodified sources); this script only chooses the flags
and stores what the driver prints, rounded once from long double to float64.

A state "at step K" is obtained by running the reference from scratch for K steps (its loop is
monolithic, src/newmd.cpp:96), with -W K so that compute_energy has run on the last step
(src/newmd.cpp:174).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refdump  # noqa: E402
from np_shape_lab_b200 import mesh as M  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

DISC = ["-q", "600", "-c", "0.005", "-t", "1", "-v", "1", "-b", "40", "-s", "40"]
ROD = ["-q", "600", "-c", "0.01", "-t", "1", "-v", "1", "-b", "40", "-s", "40"]
JANUS = DISC + ["-N", "2", "-p", "0.5"]
BUCKLED_README = ["-q", "0", "-c", "0.005", "-t", "0", "-v", "0", "-b", "1", "-s", "1000", "-B", "y"]
BUCKLING_CTRL = ["-q", "0", "-B", "y", "-b", "1", "-s", "1000"]  # BASELINE.json configs[2] literal: -t 1 -v 1 stay

# name: (mesh source, D, flags, full-state steps, scalar-only trajectory steps)
CONFIGS = {
    "disc_D4": ("ref:3_4", 4, DISC, [0, 1, 2], [250, 500, 750, 1000]),
    "rod_D4": ("ref:3_4", 4, ROD, [0, 1], []),
    "janus_D4": ("ref:3_4", 4, JANUS, [0, 1], []),
    "buckled_D4": ("ref:3_4", 4, BUCKLED_README, [0, 1], [500, 1000]),
    "janus_D8": ("ref:3_8", 8, JANUS, [0, 1], [250, 500, 750, 1000]),
    "buckling_D8": ("ref:3_8", 8, BUCKLING_CTRL, [0, 1], [500, 1000]),
    "disc_ico6": ("ico:6", 6, DISC, [0, 1], []),  # synthetic class-I icosphere, 362 vertices
    # BASELINE.json configs[3], the configuration bench.py times: class-I icosphere D=64, 40 962 vertices, every field at
    # steps 0 and 1 (whole reference steps: ~25 s per force_calculation on 8 threads). -d 2e-4: see workloads.py.
    "S64_disc": ("ico:64", 64, DISC + ["-d", "0.0002"], [0, 1], []),
    # rigid volume constraint (SHAKE / RATTLE, src/functions.h:33-171); the cubic comes from oracle/shim/gsl/gsl_poly.h
    "disc_D4_GV": ("ref:3_4", 4, DISC + ["-G", "V"], [0, 1, 2], [250, 500, 1000]),
    "janus_D8_GV": ("ref:3_8", 8, JANUS + ["-G", "V"], [0, 1], [500, 1000]),
    # explicit counterions: 64 monovalent ions placed by the driver (the reference's own placement
    # loop cannot terminate with its hard-wired box, see oracle/ref_driver.cpp:place_ions)
    "disc_D4_ions": ("ref:3_4", 4, DISC + ["--ions", "64"], [0, 1, 2], [100, 250, 500]),
    # gradual annealing of the mesh thermostat (src/newmd.cpp:258-299) with CONTROL::annealfreq = 400 and annealDuration =
    # 100 instead of 50 000 / 5 000: windows at steps 400..499 (T / 1.25) and 800..899 (T / 2) inside a 1000-step run
    "disc_D4_anneal": ("ref:3_4", 4, DISC + ["--anneal-freq", "400", "--anneal-duration", "100"], [0], [399, 400, 450, 500, 799, 850, 900, 1000]),
}

FULL_KEYS = ["pos", "vel", "force", "bforce", "sforce", "tforce", "vforce", "pforce", "itsS", "face_area",
             "face_volume"]


def build(name: str, threads: int | None) -> None:
    src, D, flags, full_steps, traj_steps = CONFIGS[name]
    scratch = os.path.join("/tmp", "npsl_golden")
    os.makedirs(scratch, exist_ok=True)
    if src.startswith("ref:"):
        mesh_file = os.path.join(REF, "bin", "infiles", src[4:] + ".dat")
        mesh = M.load_dat(mesh_file)
    else:
        mesh = M.icosphere(int(src[4:]))
        mesh_file = os.path.join(scratch, "ico_%s.dat" % src[4:])
        M.write_dat(mesh_file, mesh)
    wd = refdump.make_workdir(mesh_file, D, root=scratch)
    out = {}
    first = None
    for k in full_steps:
        d = refdump.dump_at_step(wd, D, flags, k, threads=threads, writedata=max(k, 1) if k >= 3 else 500)
        if first is None:
            first = d
            assert np.array_equal(d["edges"], mesh.edges), "edge order differs from file order"
            assert np.array_equal(d["faces"], mesh.faces), "face winding differs from loader"
            if src.startswith("ico:") and mesh.n_vertices > 10000:
                # large synthetic meshes: the fixture keeps the generator's name instead of 3 MB of topology
                out["mesh_source"] = np.array(src)
                out["l0"] = d["l0"]
                assert np.array_equal(d["pos"], mesh.pos) or k != 0, "positions differ from the generator's"
            else:
                out["pos0"] = mesh.pos
                out["edges"] = d["edges"]
                out["faces"] = d["faces"]
                out["l0"] = d["l0"]
            out["q"] = d["q"]
            out["params"] = d["params"]
        for key in FULL_KEYS:
            if "mesh_source" in out and k == 0 and key in ("pos", "vel"):
                continue  # = the generator's positions (checked above) and zero
            out["k%d_%s" % (k, key)] = d[key]
        out["k%d_scalars" % k] = d["scalars"]
        out["k%d_mesh_bath" % k] = d["mesh_bath"]
        out["k%d_ions_bath" % k] = d["ions_bath"]
        if "ion_q" in d:
            if k == full_steps[0]:
                out["ion_q"] = d["ion_q"]
                out["ion_pos0"] = d["ion_pos"]
                out["ion_params"] = np.array([d["ion_diameter"], d["lj_length_mesh_ions"]])
            for key in ("ion_pos", "ion_vel", "ion_force"):
                out["k%d_%s" % (k, key)] = d[key]
            out["k%d_netIonKE" % k] = np.array(d["netIonKE"])
        print(name, "step", k, "E=%.12g" % d["energy"], flush=True)
    for k in traj_steps:
        d = refdump.dump_at_step(wd, D, flags, k, threads=threads, writedata=k)
        out["k%d_scalars" % k] = d["scalars"]
        out["k%d_mesh_bath" % k] = d["mesh_bath"]
        out["k%d_ions_bath" % k] = d["ions_bath"]
        if "ion_q" in d:
            out["k%d_netIonKE" % k] = np.array(d["netIonKE"])
        if k == traj_steps[-1]:
            out["k%d_pos" % k] = d["pos"]
            if "ion_q" in d:
                out["k%d_ion_pos" % k] = d["ion_pos"]
        print(name, "step", k, "A=%.10g U=%.10g E=%.10g" % (d["total_area"], d["penergy"], d["energy"]), flush=True)
    out["full_steps"] = np.array(full_steps, dtype=np.int64)
    out["traj_steps"] = np.array(traj_steps, dtype=np.int64)
    out["scalar_names"] = np.array(refdump.SCALARS)
    out["param_names"] = np.array(refdump.PARAMS)
    out["flags"] = np.array(" ".join(["-D", str(D)] + flags))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    shutil.rmtree(wd, ignore_errors=True)


def readme_series() -> None:
    """The reference's shipped golden series (examples/*_Control, 6 significant digits) and the
    README's final A/U/E (README.md:33,41,47,53) as one small JSON fixture."""
    ex = {
        "disc_D4": "HomogeneouslyCharged_Disc_Control",
        "rod_D4": "HomogeneouslyCharged_Rod_Control",
        "janus_D4": "InhomogeneouslyCharged_Hemisphere_Control",
        "buckled_D4": "Uncharged_Buckled_Control",
    }
    readme = {
        "disc_D4": {"A": 15.675, "U": 2620.98, "E": 2699.97},
        "rod_D4": {"A": 13.259, "U": 1919.14, "E": 1939.96},
        "janus_D4": {"A": 12.53, "U": 1224.0, "E": 1437.0},
        "buckled_D4": {"A": 12.3795, "U": 26.86, "E": 127.69},
    }
    out = {}
    for name, d in ex.items():
        base = os.path.join(REF, "examples", d)

        def table(fn):
            rows = []
            for line in open(os.path.join(base, fn)):
                t = line.split()
                if t:
                    rows.append([float(x) for x in t])
            return rows

        out[name] = {
            "flags": " ".join(CONFIGS[name][2]),
            "readme_final": readme[name],
            "energy_nanomembrane": table("energy_nanomembrane.dat"),
            "energy_in_parts": table("energy_in_parts_kE_bE_sE_tE_ljE_esE.dat"),
            "area": table("area.dat"),
            "volume": table("volume.dat"),
        }
    with open(os.path.join(GOLD, "reference_examples.json"), "w") as fh:
        json.dump(out, fh)


LOCAL_CASES = {"disc_D4": 2, "janus_D8": 1, "buckled_D4": 1}  # config name -> step


def local_maps() -> None:
    """tests/golden/local_maps.npz: vertex areas / normals and the per-face local energy maps
    (INTERFACE::compute_local_energies(_by_component), src/interface.cpp:1140-1337) of the compiled reference in full
    precision (the driver repeats the reference's sums with the reference's own pair / edge functions) together with the
    g channel of the reference's own local_*_E.off files (6 digits), which pin the repeated sums; plus the shipped
    examples/*/local_*_E.off (positions and per-face energies after 25 000 steps, 6 digits)."""
    out = {}
    scratch = os.path.join("/tmp", "npsl_golden")
    os.makedirs(scratch, exist_ok=True)
    for name, step in LOCAL_CASES.items():
        src, D, flags, _, _ = CONFIGS[name]
        mesh_file = os.path.join(REF, "bin", "infiles", src[4:] + ".dat")
        wd = refdump.make_workdir(mesh_file, D, root=scratch)
        d = refdump.dump_at_step(wd, D, flags, step, local=True)
        lf = d["local_faces"]
        for k, key in enumerate(("electrostatic", "elastic", "bending", "stretching")):
            off = d["off_" + key]
            tol = 1e-5 * max(np.max(np.abs(off)), 1e-300)
            assert np.max(np.abs(lf[:, k] - off)) <= tol, (name, key)
            out["%s_off_%s" % (name, key)] = off
        out[name + "_step"] = np.array(step)
        out[name + "_pos"] = d["pos"]
        out[name + "_vertex_area"] = d["vertex_area"]
        out[name + "_vertex_normal"] = d["vertex_normal"]
        out[name + "_local_faces"] = lf
        print("local maps", name, "step", step, "es range %.6g..%.6g" % (lf[:, 0].min(), lf[:, 0].max()), flush=True)
        shutil.rmtree(wd, ignore_errors=True)
    ex = {"disc_D4": "HomogeneouslyCharged_Disc_Control", "rod_D4": "HomogeneouslyCharged_Rod_Control",
          "janus_D4": "InhomogeneouslyCharged_Hemisphere_Control", "buckled_D4": "Uncharged_Buckled_Control"}
    for name, d in ex.items():
        pos = None
        for key, fn in refdump.LOCAL_OFF.items():
            p, e = refdump.parse_local_off(os.path.join(REF, "examples", d, fn))
            assert pos is None or np.array_equal(pos, p)
            pos = p
            out["example_%s_%s" % (name, key)] = e
        out["example_%s_pos" % name] = pos
    np.savez_compressed(os.path.join(GOLD, "local_maps.npz"), **out)


def writers() -> None:
    """tests/golden/writers_disc_D4.npz: what the reference's snapshot writers leave behind in a 5000-step README disc
    run -- p.lammpstrj frames (interface_movie, every 1000 steps; frames 1000 and 5000 are kept), 5000.off
    (interface_off) and the four local_*_E.off maps of the final state."""
    name, K = "disc_D4", 5000
    src, D, flags, _, _ = CONFIGS[name]
    scratch = os.path.join("/tmp", "npsl_golden")
    os.makedirs(scratch, exist_ok=True)
    wd = refdump.make_workdir(os.path.join(REF, "bin", "infiles", src[4:] + ".dat"), D, root=scratch)
    d = refdump.dump_at_step(wd, D, flags, K, writedata=K, local=True)
    steps, frames = refdump.parse_lammpstrj(os.path.join(wd, "outfiles", "p.lammpstrj"))
    keep = [int(np.where(steps == s - 1)[0][0]) for s in (1000, K)]  # the writer labels step num as num - 1
    pos, faces, colours = refdump.parse_shape_off(os.path.join(wd, "outfiles", "%d.off" % K))
    out = {"steps": np.array(K), "movie_timesteps": steps[keep], "movie_frames": frames[keep], "movie_all_timesteps": steps,
           "off_pos": pos, "off_faces": faces, "off_colours": colours}
    for key in refdump.LOCAL_OFF:
        out["local_" + key] = d["off_" + key]
    out["local_full"] = d["local_faces"]
    np.savez_compressed(os.path.join(GOLD, "writers_disc_D4.npz"), **out)
    print("writers", name, "frames", steps.tolist(), flush=True)
    shutil.rmtree(wd, ignore_errors=True)


WINDOW_CASES = {
    # BASELINE.json configs[4]: class-I icosphere D=128, 163 842 vertices, rod recipe. A whole reference step is ~7 min on
    # 8 threads, so the pair force is pinned on three windows of 512 rows through the reference's own row bounds
    # (src/main.cpp:448-455, src/newforces.cpp:153-177) and the mesh terms in full; "perturbed" repeats it away from
    # the symmetric sphere (every coordinate moved by up to 2e-4, about 2 % of an edge), mesh terms on the window rows.
    "S128_rod": ("S128_rod", [(0, 511), (81664, 82175), (163330, 163841)], 2e-4),
    "S64_disc_perturbed": ("S64_disc", [(0, 511), (20224, 20735), (40450, 40961)], 4e-4),
}


WINDOW_PARAMS = ["scalefactor", "em", "inv_kappa_out", "elj", "lj_length_mesh_mesh", "bkappa", "sconstant", "sigma_a",
                 "sigma_v", "ref_area", "ref_volume", "avg_edge_length"]


def lcg_displacements(n: int, amp: float) -> np.ndarray:
    """The displacement --perturb applies (oracle/ref_driver.cpp, mode window): amp * (2u - 1) per coordinate, u the top
    53 bits of a fixed 64-bit LCG."""
    st = 0x9E3779B97F4A7C15
    out = np.empty(3 * n)
    mask = (1 << 64) - 1
    for k in range(3 * n):
        st = (st * 6364136223846793005 + 1442695040888963407) & mask
        out[k] = amp * (2.0 * ((st >> 11) * (1.0 / 9007199254740992.0)) - 1.0)
    return out.reshape(n, 3)


def windows(name: str, threads: int | None) -> None:
    """tests/golden/<name>_windows.npz from ``npsl_ref --mode window`` (the reference's own force_calculation per row
    window)."""
    from np_shape_lab_b200 import workloads as W
    wl, wins, amp = WINDOW_CASES[name]
    w = W.build(wl)
    n = w.n_vertices
    scratch = os.path.join("/tmp", "npsl_golden")
    os.makedirs(scratch, exist_ok=True)
    mesh_file = os.path.join(scratch, "ico_%d.dat" % w.D)
    M.write_dat(mesh_file, w.mesh)
    wd = refdump.make_workdir(mesh_file, w.D, root=scratch)
    out = {"workload": np.array(wl), "windows": np.array(wins, dtype=np.int64), "perturb": np.array(amp)}
    for tag, a in (("k0", 0.0), ("pert", amp)):
        od = os.path.join(wd, "out_" + tag)
        os.makedirs(od)
        js = refdump.last_json_line(refdump.run_ref(
            wd, ["--mode", "window", "--out", od, "--windows", ",".join("%d:%d" % t for t in wins), "--perturb", a] +
            W.reference_flags(w), threads=threads))
        rd = lambda k, shape=None: np.fromfile(os.path.join(od, k + ".f64")).reshape(shape or -1)
        rows = rd("rows").astype(np.int64)
        pos = rd("pos", (n, 3))
        expect = w.mesh.pos if a == 0 else w.mesh.pos + lcg_displacements(n, a)
        assert np.array_equal(pos, expect), "positions differ from the host's"
        # the host's set-up (np_shape_lab_b200/mesh.py) agrees with the reference's to rounding; the fixture keeps the
        # reference's own charges and derived parameters so that both sides of the comparison see identical inputs
        assert np.max(np.abs(rd("q") - w.q)) <= 1e-14 * np.max(np.abs(w.q)), "charges differ from the host's assignment"
        for k in WINDOW_PARAMS:
            assert abs(js[k] - w.params[k]) <= 1e-13 * abs(js[k]), (k, js[k], w.params[k])
        if a == 0:
            out["q"] = rd("q")
            out["params"] = np.array([js[k] for k in WINDOW_PARAMS])
            out["param_names"] = np.array(WINDOW_PARAMS)
        out[tag + "_rows"] = rows
        out[tag + "_pforce"] = rd("pforce", (len(rows), 3))
        for k in ("bforce", "sforce", "tforce", "vforce"):
            full = rd(k, (n, 3))
            out[tag + "_" + k] = full if a == 0 else full[rows]
        out[tag + "_total_area"] = np.array(js["total_area"])
        out[tag + "_total_volume"] = np.array(js["total_volume"])
        print(name, tag, "rows", len(rows), "force_s", js["force_s"], "A=%.12g V=%.12g" % (js["total_area"], js["total_volume"]), flush=True)
    np.savez_compressed(os.path.join(GOLD, name + "_windows.npz"), **out)
    shutil.rmtree(wd, ignore_errors=True)


CHARGE_CASES = dict(
    [("N%d" % n, ["-N", str(n), "-p", "0.5"]) for n in range(3, 18)] +
    [("N2_Hy", ["-N", "2", "-p", "0.5", "-H", "y"]), ("N1_Hc", ["-N", "1", "-H", "c"]),
     ("N1_Fy", ["-F", "y"]), ("N2_Fy_p0.3", ["-F", "y", "-N", "2", "-p", "0.3"]), ("N5_Fy", ["-F", "y", "-N", "5", "-p", "0.5"]),
     ("N4_p0.3", ["-N", "4", "-p", "0.3"]), ("alpha0.5_Fn", ["--alpha", "0.5"]), ("alpha0.6_Fy", ["--alpha", "0.6", "-F", "y"]),
     ("external", ["-E", "test_pattern"])])


def charge_patterns() -> None:
    """tests/golden/charge_patterns.npz: VERTEX::q after INTERFACE::assign_random_q_values / assign_external_q_values of
    the compiled reference on the 372-vertex mesh (q = 600) for the stripe patterns -N 3..17, yin-yang, cube, shuffled
    (-F y), fractional-occupancy and external-file variants (src/interface.cpp:339-862)."""
    scratch = os.path.join("/tmp", "npsl_golden")
    os.makedirs(scratch, exist_ok=True)
    mesh_file = os.path.join(REF, "bin", "infiles", "3_4.dat")
    mesh = M.load_dat(mesh_file)
    wd = refdump.make_workdir(mesh_file, 4, root=scratch)
    os.makedirs(os.path.join(wd, "infiles", "Charge_Patterns"))
    rng = np.random.default_rng(11)
    idx = rng.permutation(mesh.n_vertices)[:300]
    val = rng.uniform(-0.02, 0.05, size=300)
    with open(os.path.join(wd, "infiles", "Charge_Patterns", "test_pattern.dat"), "w") as fh:
        for i, v in zip(idx, val):
            fh.write("%d %.12g\n" % (i, v))
    out = {"file_q": mesh.file_q, "external_index": idx.astype(np.int32), "external_value": np.array([float("%.12g" % v) for v in val])}
    for name, flags in CHARGE_CASES.items():
        d = refdump.dump_at_step(wd, 4, ["-q", "600", "-c", "0.005"] + flags, 0)
        out[name] = d["q"]
        out[name + "_flags"] = np.array(" ".join(flags))
        print("charges", name, "sum %.12g nonzero %d" % (d["q"].sum(), int(np.count_nonzero(d["q"]))), flush=True)
    np.savez_compressed(os.path.join(GOLD, "charge_patterns.npz"), **out)
    shutil.rmtree(wd, ignore_errors=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--threads", type=int, default=None)
    a = ap.parse_args()
    if not refdump.have_ref():
        raise SystemExit("build oracle/_ref first: make -C oracle ref")
    os.makedirs(GOLD, exist_ok=True)
    readme_series()
    if a.only in (None, "local_maps"):
        local_maps()
    if a.only in (None, "writers"):
        writers()
    if a.only in (None, "charge_patterns"):
        charge_patterns()
    for name in CONFIGS:
        if a.only and a.only != name:
            continue
        build(name, a.threads)
    for name in WINDOW_CASES:
        if a.only in (None, name + "_windows"):
            windows(name, a.threads)


if __name__ == "__main__":
    main()
