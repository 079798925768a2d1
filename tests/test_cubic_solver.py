"""gsl_poly_solve_cubic, the one third-party routine on the -G V path (src/functions.h:102), is restated in
oracle/shim/gsl/gsl_poly.h (compiled into the reference build), in oracle/npsl_oracle.c and on the device
(csrc/mesh_kernels.cuh cubic_x0). GSL itself is not in the image, so the three restatements agreeing with each other
proves nothing about GSL. This test pins the restatement to what GSL's manual promises for the function -- the real
roots of x^3 + a x^2 + b x + c in ascending order, return value 1 or 3 -- by an independent method: numpy's
companion-matrix eigenvalues (LAPACK), on random cubics, on cubics with prescribed roots (including the scales SHAKE
produces: one root ~ -1e-8 beside two of order +-1e4) and on the degenerate branches."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cstdio>
#include "gsl/gsl_poly.h"
int main() {
    double a, b, c;
    while (std::scanf("%lf %lf %lf", &a, &b, &c) == 3) {
        double x0 = 0, x1 = 0, x2 = 0;
        const int n = gsl_poly_solve_cubic(a, b, c, &x0, &x1, &x2);
        std::printf("%d %.17g %.17g %.17g\n", n, x0, x1, x2);
    }
    return 0;
}
"""


@pytest.fixture(scope="module")
def solver(tmp_path_factory):
    d = tmp_path_factory.mktemp("cubic")
    src, exe = d / "cubic.cpp", d / "cubic"
    src.write_text(HARNESS)
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-I", os.path.join(ROOT, "oracle", "shim"), str(src), "-o", str(exe)])

    def run(coeffs):
        txt = "\n".join("%.17g %.17g %.17g" % tuple(c) for c in coeffs) + "\n"
        out = subprocess.run([str(exe)], input=txt, stdout=subprocess.PIPE, text=True, check=True).stdout
        rows = [ln.split() for ln in out.strip().splitlines()]
        return [(int(r[0]), [float(x) for x in r[1:]]) for r in rows]
    return run


def real_roots(a, b, c):
    r = np.roots([1.0, a, b, c])
    scale = max(1.0, np.max(np.abs(r)))
    return np.sort(r[np.abs(r.imag) < 1e-7 * scale].real)


def test_random_cubics_against_companion_matrix(solver):
    rng = np.random.default_rng(11)
    coeffs = rng.standard_normal((400, 3)) * 10.0 ** rng.integers(-2, 3, size=(400, 1))
    for (a, b, c), (n, x) in zip(coeffs, solver(coeffs)):
        want = real_roots(a, b, c)
        assert n in (1, 3)
        if len(want) != n:  # numerically a near-double root: the residual decides
            want = None
        got = x[:n]
        assert got == sorted(got)
        for r in got:
            p = ((r + a) * r + b) * r + c
            dp = abs((3 * r + 2 * a) * r + b)
            # |p| small against the size of its terms, or the root is within rounding of a (near-)multiple root
            assert abs(p) <= 1e-9 * (abs(r) ** 3 + abs(a) * r * r + abs(b * r) + abs(c)) or dp < 1e-6, (a, b, c, r, p)
        if want is not None:
            assert np.allclose(got, want, rtol=1e-7, atol=1e-7 * max(1.0, np.max(np.abs(want)))), (a, b, c, got, want)


def test_prescribed_roots_including_shake_scales(solver):
    cases = [(-3.0, 1.0, 2.5), (-1e4, -2.5e-8, 3e3), (-7.1e3, -1.3e-9, 6.4e3), (0.5, 0.5, 4.0), (-2.0, -2.0, 5.0),
             (-1e-3, 2e-3, 7.0)]
    coeffs = []
    for r0, r1, r2 in cases:  # (x - r0)(x - r1)(x - r2)
        coeffs.append((-(r0 + r1 + r2), r0 * r1 + r0 * r2 + r1 * r2, -r0 * r1 * r2))
    for roots, (n, x) in zip(cases, solver(coeffs)):
        want = sorted(roots)
        assert n == 3
        # x0 is what SHAKE uses (src/functions.h:102-110): the smallest real root
        assert abs(x[0] - want[0]) <= 1e-9 * max(abs(want[0]), 1e-6 * max(abs(w) for w in want)) or want[0] == want[1]
        assert np.allclose(sorted(x), want, rtol=1e-7, atol=1e-9 * max(abs(w) for w in want))


def test_single_real_root_and_triple_root(solver):
    (n1, x1), (n3, x3) = solver([(0.0, 1.0, 0.0), (-6.0, 12.0, -8.0)])  # x (x^2 + 1);  (x - 2)^3
    assert n1 == 1 and abs(x1[0]) < 1e-15
    assert n3 == 3 and np.allclose(x3, [2.0, 2.0, 2.0])
