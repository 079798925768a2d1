"""Generates the minimax polynomial for 2^f on [-1/2, 1/2] used by the pair kernel's
range-reduced exponential (np-shape-lab_b200/csrc/pair_kernel.cuh). Remez exchange in
50-digit arithmetic (mpmath); prints C hex-float literals and the achieved relative error
both in exact arithmetic and when evaluated in double-precision Horner form."""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def remez(deg, a, b, iters=30):
    # relative-error minimax: minimise max |p(x)/2^x - 1|
    n = deg + 2
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (n - 1 - i) / (n - 1)) for i in range(n)]
    f = lambda x: mp.power(2, x)
    for _ in range(iters):
        A = mp.matrix(n, n)
        rhs = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for k in range(deg + 1):
                A[i, k] = x ** k / f(x)
            A[i, deg + 1] = (-1) ** i
            rhs[i] = 1
        sol = mp.lu_solve(A, rhs)
        c = [sol[k] for k in range(deg + 1)]
        err = lambda x: mp.polyval(c[::-1], x) / f(x) - 1
        # find extrema of err on a fine grid, then refine
        grid = [a + (b - a) * mp.mpf(i) / 4000 for i in range(4001)]
        vals = [err(x) for x in grid]
        ext = [grid[0]]
        for i in range(1, 4000):
            if (vals[i] - vals[i - 1]) * (vals[i + 1] - vals[i]) <= 0:
                try:
                    xr = mp.findroot(lambda x: mp.diff(err, x), grid[i])
                    if a <= xr <= b:
                        ext.append(xr)
                        continue
                except Exception:
                    pass
                ext.append(grid[i])
        ext.append(grid[-1])
        # dedupe & keep n alternating extrema
        ext = sorted(set([mp.nstr(e, 40) for e in ext]), key=lambda s: mp.mpf(s))
        ext = [mp.mpf(s) for s in ext]
        if len(ext) < n:
            break
        if len(ext) > n:
            ext = sorted(ext, key=lambda x: -abs(err(x)))[:n]
            ext = sorted(ext)
        xs = ext
    E = max(abs(err(x)) for x in grid)
    return c, E


deg = int(sys.argv[1]) if len(sys.argv) > 1 else 10
c, E = remez(deg, mp.mpf(-0.5), mp.mpf(0.5))
print("degree", deg, "exact-arithmetic max rel err %.3e" % float(E))
cd = [float(x) for x in c]
# double Horner evaluation error
xs = np.linspace(-0.5, 0.5, 200001)
p = np.full_like(xs, cd[-1])
for k in range(deg - 1, -1, -1):
    p = p * xs + cd[k]   # not fused; fma on device is slightly better
ref = np.array([float(mp.power(2, mp.mpf(float(x)))) for x in xs[::50]])
print("double Horner max rel err %.3e" % np.max(np.abs(p[::50] / ref - 1)))
for k, v in enumerate(cd):
    print("    %s,  // c%d = %.17g" % (float.hex(v), k, v))
