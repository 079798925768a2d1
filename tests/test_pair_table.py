"""Accuracy of the pair kernel's piecewise-polynomial radial factor (csrc/pair_sym_kernel.cuh, npsl.cu build_pair_table),
checked on the host against an extended-precision evaluation of
    G(s) = e^{-r/lambda} (1/r^3 + 1/(lambda r^2)),  r = sqrt(s)
-- the scalar the reference multiplies r_vec with (src/newforces.cpp:173-174). No GPU needed: npsl_pair_poly_eval rebuilds
the table and runs the device's Horner scheme in FP64 on the host."""
import numpy as np
import pytest

from np_shape_lab_b200 import capi

LAMBDAS = [0.4301292881004416, 0.3041473364, 1.0, 100000000.0]  # disc, rod, weak screening, c = 0 (src/main.cpp:235-239)


def samples():
    rng = np.random.default_rng(7)
    # log-uniform over the whole covered range plus a dense sweep of every interval's edges and interior
    s = np.exp(rng.uniform(np.log(2.0 ** -21), np.log(2.0 ** 5), size=400000))
    k = np.arange(26 * 32)
    lo = np.ldexp(1.0 + (k % 32) / 32.0, -21 + k // 32)
    w = np.ldexp(1.0, -21 + k // 32 - 5)
    edge = np.concatenate([lo, np.nextafter(lo + w, 0.0), lo + 0.5 * w, lo + 0.146 * w, lo + 0.854 * w])
    return np.concatenate([s, edge])


@pytest.mark.parametrize("lam", LAMBDAS)
def test_table_reproduces_the_radial_factor(lam):
    s = samples()
    s = s[(s >= 2.0 ** -21) & (s < 2.0 ** 5)]
    tv, ev = capi.pair_poly_eval(lam, s)
    rel = np.abs(tv - ev) / ev
    u = np.sqrt(s) / lam  # r / lambda
    # The error of a degree-5 fit on an interval of relative width 1/32 grows with the logarithmic slope of G, which is
    # 3/2 near contact and ~r / (2 lambda) in the tail: 1.3e-12 up to r = lambda, < 1e-11 up to 4 lambda, ~1e-10 at
    # 10 lambda -- where a pair contributes e^-10 of a near pair. Contract of the path: 1e-10 on per-vertex forces.
    assert rel[u <= 1].max() < 2e-12, rel[u <= 1].max()
    assert rel[u <= 4].max() < 1e-11, rel[u <= 4].max()
    assert rel.max() < 1e-9, rel.max()  # r = 18 lambda at the top of the range for the rod recipe: weight e^-18
    # weighted by what a pair contributes to a force (G r against its value at r = lambda): 1e-12 everywhere
    g1 = capi.pair_poly_eval(lam, np.array([min(lam * lam, 1.0)]))[1][0] * min(lam, 1.0)
    far = u > 1
    assert not far.any() or (np.abs(tv - ev) * np.sqrt(s) / g1)[far].max() < 2e-12


def test_table_is_continuous_across_interval_edges():
    lam = LAMBDAS[0]
    k = np.arange(1, 26 * 32)
    edge = np.ldexp(1.0 + (k % 32) / 32.0, -21 + k // 32)
    below = np.nextafter(edge, 0.0)
    tv_hi, _ = capi.pair_poly_eval(lam, edge)
    tv_lo, _ = capi.pair_poly_eval(lam, below)
    u = np.sqrt(edge) / lam
    jump = np.abs(tv_hi - tv_lo) / tv_hi
    assert jump[u <= 4].max() < 2e-11 and jump.max() < 3e-10


def test_reduced_accuracy_scheme_error_budget():
    """The DEFAULT evaluation of k_pair_sym (second-order rsqrt correction, 2048-entry index-biased 2^(j/2048) table,
    quadratic, exponent added by one integer multiply-add), restated operation for operation on the host
    (npsl_pair_fast_eval): relative error of e^{-r}(1/r^3 + 1/r^2) against extended precision, for a seed no better than
    20 bits (the device's MUFU.RSQ64H seed is better). The error of e^{-r} grows like r x (error of r), so the bound is
    stated as rel / (1 + r); what matters for a force is the error weighted with the size of the term, far below 1e-12
    of the largest term everywhere. The contract of the path is 1e-10."""
    rng = np.random.default_rng(0)
    r = np.concatenate([10.0 ** rng.uniform(-2, 1, 200000), np.linspace(0.01, 700.0, 200000)])
    for bits, bound in ((20, 5e-12), (22, 8e-13)):
        v, e = capi.pair_fast_eval(r * r, bits)
        assert np.isfinite(v).all() and (v > 0).all()
        rel = np.abs(v - e) / e
        assert (rel / (1.0 + r)).max() < bound, (bits, (rel / (1.0 + r)).max())
        far = r >= 1.0
        assert np.abs(v - e)[far].max() < 1e-12 * 2.0 * np.exp(-1.0)  # weighted with the term's size, against G(1)
    # the exponent path: k = n >> 11 down to about -1000 (r = 700 lambda) through the biased table, still a normal number
    v, e = capi.pair_fast_eval(np.array([690.0 ** 2, 700.0 ** 2]), 24)
    assert np.all(np.abs(v - e) / e < 1e-9) and np.all(v > 0)
