// Screened-Coulomb (Debye-Hueckel) all-pairs kernel, FP64, sm_100a.
//
// Replaces the mesh-mesh loop of force_calculation (reference src/newforces.cpp:153-177, init
// twin :18-42) and the pair part of INTERFACE::compute_energy (src/interface.cpp:1046-1065,
// src/newenergies.cpp:21-25), including their MPI row decomposition (src/mpi_utility.h,
// src/main.cpp:448-455): a launch covers the owned rows [row_lo,row_hi) against all N columns.
//
// Per ordered pair (i,j), i != j, with r = |x_i - x_j| and lambda = inv_kappa_out:
//   F_i += (sf/em) q_i q_j exp(-r/lambda) (1/r + 1/lambda)/r^2 (x_i - x_j)
//   U   += 1/2 (sf/em) q_i q_j exp(-r/lambda)/r
// The common factor (sf/em) q_i is applied once per row by the consumer (mesh_kernels.cuh);
// this kernel accumulates  sum_j q_j e (1/r^2 + 1/(lambda r))/r (x_i-x_j)  and  sum_j q_j e/r.
//
// Layout: positions and charge are one 32-byte record per vertex (x,y,z,q) so that a column tile
// is staged into shared memory with fully coalesced 16-byte loads and read back with two
// broadcast LDS.128 per column. Each thread keeps ROWS rows (positions + accumulators) in
// registers. The grid is (row tiles) x (column splits); every (split,row) partial sum goes to its
// own slot and is combined later in split order, so results are bit-reproducible run to run and
// independent of how rows are sharded across GPUs. No atomics, no tensor cores (this is an
// N-body with a transcendental per pair, not a contraction).
//
// FP64-pipe work per ordered pair: 27 instructions force-only, 28.5 with energy:
//   3 sub, 3 r^2, 5 rsqrt (MUFU.RSQ64H seed + one third-order step), 1 r, 3 range reduction,
//   4 exp (1024-entry 2^(j/1024) table in shared memory + cubic polynomial), 5 force factor, 3 FMA.
// exp: x = -r/lambda <= 0, so 2^x = 2^k T[j] (1 + f g(f)) needs no overflow / NaN handling; k is
// clamped at -1000 on the integer pipe. exp and rsqrt hold <= 3e-16 relative error.
// On B200 the FP64 pipe is limited by register-file operand bandwidth (tools/ubench/fp64_pipe.cu): an
// instruction with one register source issues every 2.0 cycles, two 2.3, three 3.1. The arithmetic
// below is arranged so that constants sit in uniform registers / immediates wherever possible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "exp2_table.inc"

namespace npsl {

constexpr int kPairThreads = 128;  // threads per CTA
constexpr int kPairTileJ = 256;    // columns staged per shared-memory tile (8 KB)
constexpr int kExpTab = 1 << NPSL_EXP2_TB;

// 2^(j / 2^TB), j < 2^TB (8 KB at TB = 10): lives in global memory (L2-resident) and is staged into shared memory by every
// CTA with coalesced loads; the per-lane lookups in the inner loop then hit shared memory.
__device__ const double c_exp2_tab[kExpTab] = {NPSL_EXP2_TABLE};

struct PairParams {
    double c2s;         // -log2(e)/lambda * 2^TB
    double inv_lambda;  // 1/lambda
    int n;              // number of vertices (columns)
    int row_lo, row_hi; // owned rows, half-open
    int cols_per_cta;   // columns handled by one CTA (column chunk width, multiple of kPairTileJ)
    int n_splits;       // number of column chunks
    int row_stride;     // rows in a partial slab (>= row_hi-row_lo, padded)
    int lj_cut_hi;      // high word of the LJ cutoff r^2 (conservative detection on the integer pipe)
};

__device__ __forceinline__ double rsqrt_seed(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));  // MUFU.RSQ64H
    return y;
}

// 1/sqrt(a) for normal a > 0: seed (~2^-22) + one third-order correction -> ~1 ulp.
__device__ __forceinline__ double rsqrt_fast(double a) {
    double y0 = rsqrt_seed(a);
    double t = a * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    double ye = y0 * e;
    return fma(ye, p, y0);
}

// 2^(r*c2s/2^TB) for r*c2s <= 0 (no overflow / NaN paths). Magic-number rounding leaves
// n = round(r*c2s) in the low word of tt; f = r*c2s - n in one fused step; T[n mod 2^TB] comes from
// shared memory and takes the exponent k = n >> TB on the integer pipe.
__device__ __forceinline__ double exp2_neg(double r, double c2s, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double tt = fma(r, c2s, MAGIC);
    double kf = tt - MAGIC;
    double f = fma(r, c2s, -kf);
    const int n = __double2loint(tt);
    const double T = tab[n & (kExpTab - 1)];
    const int k = max(n >> NPSL_EXP2_TB, -1000);
    const double Tk = __hiloint2double(__double2hiint(T) + (k << 20), __double2loint(T));
    double g = fma(f, NPSL_EXP2_G3, NPSL_EXP2_G2);
    g = fma(g, f, NPSL_EXP2_G1);
    double q = f * g;
    return fma(Tk, q, Tk);
}

struct PairAcc {
    double fx, fy, fz, ue;
};

template <bool ENERGY>
__device__ __forceinline__ void pair_accumulate(double xi, double yi, double zi, double xj, double yj, double zj,
                                                double qj, const double c2s, const double inv_lambda,
                                                const double *__restrict__ tab, PairAcc &a, int &min_r2_hi) {
    double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    min_r2_hi = min(min_r2_hi, __double2hiint(r2));
    double rinv = rsqrt_fast(r2);
    double r = r2 * rinv;
    double ex = exp2_neg(r, c2s, tab);
    double qe = qj * ex;                  // q_j e^{-r/lambda}
    double u = rinv * rinv;
    double g = rinv + inv_lambda;         // one register source
    double w = u * g;                     // 1/r^3 + 1/(lambda r^2)
    double s = qe * w;
    a.fx = fma(s, dx, a.fx);
    a.fy = fma(s, dy, a.fy);
    a.fz = fma(s, dz, a.fz);
    if (ENERGY) a.ue = fma(qe, rinv, a.ue);
}

template <int ROWS>
struct PairUnroll {
    static constexpr int value = ROWS >= 4 ? 1 : 2;
};

// grid = (row tiles, column chunks); block = kPairThreads; ROWS rows per thread.
// partial: [n_splits][row_stride] records (fx,fy,fz,ue); row index is relative to row_lo.
// lj_hit: set to 1 when any evaluated pair may lie inside the LJ cutoff (exact test in k_lj).
template <int ROWS, bool ENERGY>
__global__ void __launch_bounds__(kPairThreads) k_pair(const double4 *__restrict__ xyzq, double4 *__restrict__ partial,
                                                      int *__restrict__ lj_hit, const PairParams P) {
    __shared__ double4 sj[kPairTileJ];
    __shared__ double tab[kExpTab];
    const int t = threadIdx.x;
    for (int k = t; k < kExpTab; k += kPairThreads) tab[k] = c_exp2_tab[k];
    const int row0 = P.row_lo + blockIdx.x * (kPairThreads * ROWS);
    const int col0 = blockIdx.y * P.cols_per_cta;
    const int col1 = min(col0 + P.cols_per_cta, P.n);

    int irow[ROWS];
    double xi[ROWS], yi[ROWS], zi[ROWS];
    PairAcc acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        irow[r] = row0 + r * kPairThreads + t;
        int ic = min(irow[r], P.row_hi - 1);  // rows past the end recompute the last row, never stored
        double4 v = xyzq[ic];
        xi[r] = v.x; yi[r] = v.y; zi[r] = v.z;
        acc[r].fx = acc[r].fy = acc[r].fz = acc[r].ue = 0.0;
        irow[r] = ic;
    }
    const double c2s = P.c2s, inv_lambda = P.inv_lambda;
    int min_r2_hi = 0x7fffffff;
    const int row_end = min(row0 + kPairThreads * ROWS, P.row_hi);

    for (int j0 = col0; j0 < col1; j0 += kPairTileJ) {
        const int cnt = min(kPairTileJ, col1 - j0);
        __syncthreads();
        for (int k = t; k < cnt; k += kPairThreads) sj[k] = xyzq[j0 + k];
        __syncthreads();
        const bool diag = (j0 < row_end) && (j0 + cnt > row0);  // tile contains some i == j
        if (!diag) {
#pragma unroll(PairUnroll<ROWS>::value)
            for (int k = 0; k < cnt; k++) {
                const double4 pj = sj[k];
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    pair_accumulate<ENERGY>(xi[r], yi[r], zi[r], pj.x, pj.y, pj.z, pj.w, c2s, inv_lambda, tab, acc[r],
                                            min_r2_hi);
            }
        } else {
            for (int k = 0; k < cnt; k++) {
                const double4 pj = sj[k];
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    // self pair: dx=dy=dz=0 exactly -> force terms vanish for any finite s; make s finite
                    // by moving the column to unit distance along x and zeroing its charge.
                    const bool self = (j0 + k == irow[r]);
                    pair_accumulate<ENERGY>(xi[r], yi[r], zi[r], self ? xi[r] - 1.0 : pj.x, pj.y, pj.z,
                                            self ? 0.0 : pj.w, c2s, inv_lambda, tab, acc[r], min_r2_hi);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        int i = row0 + r * kPairThreads + t;
        if (i < P.row_hi)
            partial[(size_t) blockIdx.y * P.row_stride + (i - P.row_lo)] =
                make_double4(acc[r].fx, acc[r].fy, acc[r].fz, acc[r].ue);
    }
    if (min_r2_hi <= P.lj_cut_hi) *lj_hit = 1;  // benign race: every writer stores the same value
}

// ---------------------------------------------------------------------------------------------
// WCA / Lennard-Jones term of the same loops (src/newforces.cpp:165-171; energy_lj_pair,
// src/newenergies.cpp:4-17). The reference tests every pair but the cutoff (2^(1/6) * 0.1 * mean edge)
// is only reached on near-collision, so this exact all-pairs pass runs only when k_pair's
// conservative detector fired (or when all charges are zero and k_pair is skipped).
// One thread per owned row, columns in fixed order -> deterministic. out: (fx,fy,fz,energy) per row,
// energy already carries the reference's 1/2 double-counting factor.
struct LjParams {
    double d2;    // lj_length^2
    double cut2;  // dcut2 * d2
    double elj;
    int n, row_lo, row_hi;
};

__global__ void __launch_bounds__(128) k_lj(const double4 *__restrict__ xyzq, double4 *__restrict__ out,
                                            const int *__restrict__ lj_hit, int force_run, const LjParams P) {
    __shared__ double4 sj[256];
    const int i = P.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const bool run = force_run || (*lj_hit != 0);
    if (!run) {
        if (i < P.row_hi) out[i - P.row_lo] = make_double4(0, 0, 0, 0);
        return;
    }
    const int ic = min(i, P.row_hi - 1);
    const double4 pi = xyzq[ic];
    double fx = 0, fy = 0, fz = 0, e = 0;
    const double d6 = P.d2 * P.d2 * P.d2, d12 = d6 * d6;
    for (int j0 = 0; j0 < P.n; j0 += 256) {
        const int cnt = min(256, P.n - j0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) sj[k] = xyzq[j0 + k];
        __syncthreads();
        for (int k = 0; k < cnt; k++) {
            const double4 pj = sj[k];
            double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < P.cut2 && (j0 + k) != ic) {
                double r6 = r2 * r2 * r2, r12 = r6 * r6;
                double s = 48 * P.elj * ((d12 / r12) - 0.5 * (d6 / r6)) * (1 / r2);
                fx += s * dx; fy += s * dy; fz += s * dz;
                e += 4 * P.elj * (d6 / r6) * ((d6 / r6) - 1) + P.elj;
            }
        }
    }
    if (i < P.row_hi) out[i - P.row_lo] = make_double4(fx, fy, fz, 0.5 * e);
}

// Pure-DFMA throughput probe: the roofline denominator for k_pair (MEASURED_PEAKS.json has no
// FP64 entry). 16 independent chains per thread, 2 flops per instruction.
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = a + k + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = fma(v[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += v[k];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace npsl
