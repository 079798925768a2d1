"""Host logic that needs no GPU: mesh formats, the synthetic icosphere, the deterministic charge
assignment and derived parameters (checked against what the reference itself derived, stored in the
golden fixtures), the row partition, and the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden, relerr
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
    assert ctypes.sizeof(capi.Energies) == 15 * 8


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
