"""Host logic that needs no GPU: mesh formats, the synthetic icosphere, the deterministic charge
assignment and derived parameters (checked against what the reference itself derived, stored in the
golden fixtures), the row partition, and the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLD, ROOT, golden, relerr
from np_shape_lab_b200 import capi
from np_shape_lab_b200 import mesh as M


@pytest.mark.parametrize("D", [1, 2, 6, 16])
def test_icosphere_counts_and_topology(D):
    m = M.icosphere(D)
    assert (m.n_vertices, m.n_edges, m.n_faces) == (10 * D * D + 2, 30 * D * D, 20 * D * D)
    assert np.allclose(np.linalg.norm(m.pos, axis=1), 1.0, atol=1e-15)
    deg = np.bincount(m.edges.reshape(-1), minlength=m.n_vertices)
    assert sorted(set(deg.tolist())) in ([5], [5, 6]) and int((deg == 5).sum()) == 12
    # outward CCW: positive volume contributions, and every directed edge used exactly once
    assert np.all(M.face_volumes(m) > 0)
    d = np.concatenate([m.faces[:, [0, 1]], m.faces[:, [1, 2]], m.faces[:, [2, 0]]])
    assert len({(int(a), int(b)) for a, b in d}) == 3 * m.n_faces
    assert abs(M.face_volumes(m).sum() - 4 * np.pi / 3) < 4.5 / (D * D)


def test_icosphere_is_bit_reproducible():
    a, b = M.icosphere(7), M.icosphere(7)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.faces, b.faces) and np.array_equal(a.edges, b.edges)


def test_dat_round_trip(tmp_path):
    m = M.icosphere(3)
    p = str(tmp_path / "3_3.dat")
    M.write_dat(p, m)
    r = M.load_dat(p)
    assert np.array_equal(r.pos, m.pos) and np.array_equal(r.edges, m.edges) and np.array_equal(r.faces, m.faces)


def test_synthetic_mesh_matches_what_the_reference_loaded():
    """disc_ico6 was produced by feeding icosphere(6) to the reference through its own .dat parser."""
    g = golden("disc_ico6")
    m = M.icosphere(6)
    assert np.array_equal(g["pos0"], m.pos) and np.array_equal(g["edges"], m.edges) and np.array_equal(g["faces"], m.faces)


CASES = {
    "disc_D4": dict(q=600.0, c=0.005), "rod_D4": dict(q=600.0, c=0.01),
    "janus_D4": dict(q=600.0, c=0.005, N=2, p=0.5), "janus_D8": dict(q=600.0, c=0.005, N=2, p=0.5),
    "buckled_D4": dict(q=0.0, c=0.005, t=0.0, v=0.0, b=1.0, s=1000.0, B="y"),
    "buckling_D8": dict(q=0.0, b=1.0, s=1000.0, B="y"), "disc_ico6": dict(q=600.0, c=0.005),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_derived_params_and_charges_match_reference(name):
    """main()'s unit scalings, Debye length, l0 / avg edge / reference volume, thermostat chains and
    assign_random_q_values (-F n) -- vs the values the compiled reference printed (src/main.cpp:196-294)."""
    g = golden(name)
    m = M.Mesh(g["pos0"], g["edges"], g["faces"])
    inp = M.PhysicalInputs(**CASES[name])
    p = M.derive_params(inp, m)
    for k, ref in g.params.items():
        assert abs(p[k] - ref) <= 1e-13 * max(abs(ref), 1e-300), (k, p[k], ref)
    assert relerr(p["l0"], g["l0"]) < 1e-15
    q = M.assign_charges(m, inp.q, inp.N, inp.p)
    assert relerr(q, g["q"], floor=1e-300) < 1e-13
    mb, ib = M.make_thermostats(inp, m.n_vertices)
    assert np.array_equal(mb, g["k0_mesh_bath"]) and np.array_equal(ib, g["k0_ions_bath"])


def test_row_partition_is_the_reference_decomposition():
    for n in (372, 972, 40962, 163842, 7):
        for world in (1, 2, 3, 4, 8):
            rng = -(-n // world)
            seen = []
            for r in range(world):
                lo, hi = M.row_partition(n, r, world)
                assert lo == min(r * rng, n) and hi == min((r + 1) * rng, n)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_library_exports_every_symbol_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "npsl.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(npsl_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(capi.SYMBOLS)
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert capi.load().npsl_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header():
    assert ctypes.sizeof(capi.BathLink) == 40
    assert ctypes.sizeof(capi.Params) == 12 * 8 + 8 + 2 * 16 * 40
    assert ctypes.sizeof(capi.Energies) == 16 * 8  # ABI 2: + ions_kinetic


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (no CPU path anywhere)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = golden("disc_D4")
    P = capi.make_params(g.params, g["k0_mesh_bath"], g["k0_ions_bath"])
    with pytest.raises(capi.NpslError) as ei:
        capi.Context(g["edges"], g["faces"], g["l0"], P)
    assert ei.value.status == -2 and "no CPU path" in str(ei.value)


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the shipped package may import, link or run it."""
    pkg = os.path.join(ROOT, "np-shape-lab_b200")
    banned = ["import oracle", "from oracle", "libnpsl_oracle", "npo_", "npsl_ref", "oracle."]
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, fn)).read()
                for b in banned:
                    assert b not in src, (fn, b)


# ---- charge patterns: stripes, yin-yang, cube, shuffled, external file -------------------------

def _pattern_cases():
    z = np.load(os.path.join(GOLD, "charge_patterns.npz"))
    return z, sorted(k for k in z.files if k + "_flags" in z.files)


def _pattern_kwargs(flags):
    kw = dict(num_divisions=1, frac_patch=0.5, alpha=1.0, random_flag="n", function_flag="n")
    names = {"-N": ("num_divisions", int), "-p": ("frac_patch", float), "-F": ("random_flag", str), "-H": ("function_flag", str),
             "--alpha": ("alpha", float)}
    ext = None
    for f, v in zip(flags[::2], flags[1::2]):
        if f == "-E":
            ext = v
        else:
            kw[names[f][0]] = names[f][1](v)
    return kw, ext


def test_charge_patterns_match_reference(tmp_path):
    """Every pattern of INTERFACE::assign_random_q_values / assign_external_q_values (src/interface.cpp:339-862) against
    VERTEX::q of the compiled reference: stripes -N 3..17, yin-yang, cube, fractional occupancy, external file and the
    shuffled (-F y) variants, whose srand(123456) + std::random_shuffle stream is reproduced (glibc / libstdc++)."""
    z, cases = _pattern_cases()
    g = golden("disc_D4")
    mesh = M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"], file_q=z["file_q"])
    assert len(cases) >= 24
    for name in cases:
        kw, ext = _pattern_kwargs(str(z[name + "_flags"]).split())
        if ext:
            pat = tmp_path / "pattern.dat"
            pat.write_text("".join("%d %.12g\n" % (i, v) for i, v in zip(z["external_index"], z["external_value"])))
            q = M.assign_external_charges(mesh, 600.0, str(pat))
        else:
            q = M.assign_charges(mesh, 600.0, **kw)
        ref = z[name]
        assert np.array_equal(q == 0, ref == 0), name
        assert np.max(np.abs(q - ref)) <= 1e-13 * np.max(np.abs(ref)), name


def test_glibc_rand_stream():
    """First outputs of srand(123456); rand() on glibc (known answers, checked against libc when it is glibc)."""
    r = M.GlibcRand(123456)
    mine = [r.rand() for _ in range(8)]
    assert mine == [1303402199, 1406886723, 391888296, 1607856518, 940034812, 1299653596, 1439853304, 6265467]


def test_host_cpp_charge_patterns_match_reference(tmp_path):
    """The C++ host mirror's assign_random_q_values / assign_external_q_values through its driver (--dump-q prints the
    charges and exits before any device work)."""
    import subprocess
    from np_shape_lab_b200 import build as B
    B.build()
    host = os.path.join(ROOT, "np-shape-lab_b200", "host_cpp")
    subprocess.check_call(["make", "-s", "-C", host])
    z, cases = _pattern_cases()
    g = golden("disc_D4")
    os.makedirs(tmp_path / "infiles" / "Charge_Patterns")
    M.write_dat(str(tmp_path / "infiles" / "3_4.dat"), M.Mesh(pos=g["pos0"], edges=g["edges"], faces=g["faces"], file_q=z["file_q"]))
    (tmp_path / "infiles" / "Charge_Patterns" / "test_pattern.dat").write_text(
        "".join("%d %.12g\n" % (i, v) for i, v in zip(z["external_index"], z["external_value"])))
    for name in cases:
        flags = [f for f in str(z[name + "_flags"]).split()]
        if "--alpha" in flags:
            continue  # alpha is hard-wired to 1 in the reference's main() (src/main.cpp:213) and in the driver
        r = subprocess.run([os.path.join(host, "np_shape_lab_b200"), "-D", "4", "-q", "600", "--dump-q", "1", "--out", ""] + flags,
                           cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, (name, r.stderr)
        q = np.array([float(x) for x in r.stdout.split("\n") if x and x[0] in "-0123456789"])
        ref = z[name]
        assert q.shape == ref.shape, (name, r.stdout[:200])
        assert np.max(np.abs(q - ref)) <= 1e-13 * np.max(np.abs(ref)), name
