// Screened-Coulomb all-pairs kernel that evaluates every unordered pair once (Newton's third law),
// FP64, sm_100a. Force-only twin of k_pair (pair_kernel.cuh) for the per-step path; it replaces the
// same reference loop (src/newforces.cpp:153-177), which walks all N(N-1) ordered pairs.
//
// Work unit = (row block I of 512 rows) x (column chunk c of chunk_tiles x 128 columns); a unit exists
// when the chunk reaches past the start of the row block, and it covers the pairs (i,j), i in I, j in c,
// j > i. Each thread keeps 4 rows in registers (position, charge, force accumulator). A 128-column tile is
// staged in shared memory (SoA) together with one force accumulator per column. Within a tile the four
// warps work on the four 32-column groups, rotating groups between four phases; inside a phase lane l
// visits column (l + s) mod 32 at step s, so at any time every column is touched by exactly one thread:
// the column accumulators need no atomics and receive their contributions in a fixed order.
//   row partials    rowpart[c][k][i]   sum over the columns of chunk c        (k = x,y,z)
//   column partials colpart[I][k][j]   minus the sum over the rows of block I
// k_pair_reduce then adds, per vertex, the partials of each of 8 "virtual ranks" (unit id mod 8) in a
// fixed order, and k_vertex adds the 8 results: the value does not depend on which GPU computed which
// unit, so trajectories are bit-identical on 1, 2, 4 and 8 GPUs.
//
// FP64-pipe work per unordered pair: 32 instructions (20 for r, 1/r, exp; 4 for the geometry factor;
// 2 charge products; 6 FMA into the two accumulators) = 16 per ordered pair, against 28 in k_pair.
#pragma once
#include "pair_kernel.cuh"

namespace npsl {

constexpr int kSymRows = 4;                          // rows per thread
constexpr int kSymRB = kPairThreads * kSymRows;      // rows per row block (512)
constexpr int kSymCT = 128;                          // columns per tile: 4 warps x 32
constexpr int kSymV = 8;                             // virtual ranks (fixed grouping of the units)

struct SymParams {
    double c2s, inv_lambda;
    int n;            // vertices
    int npad;         // plane length of the partial buffers (>= n, multiple of kSymRB)
    int chunk_tiles;  // tiles per column chunk
    int n_chunks;
    int lj_cut_hi;
};

// e^{-r/lambda} (1/r^3 + 1/(lambda r^2)) of one pair; d = x_i - x_j
__device__ __forceinline__ double pair_geom(double dx, double dy, double dz, const double c2s, const double inv_lambda,
                                            const double *__restrict__ tab, int &min_r2_hi) {
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    min_r2_hi = min(min_r2_hi, __double2hiint(r2));
    double rinv = rsqrt_fast(r2);
    double r = r2 * rinv;
    double ex = exp2_neg(r, c2s, tab);
    double u = rinv * rinv;
    double g = rinv + inv_lambda;
    double w = u * g;
    return ex * w;
}

struct SymRow {
    double x, y, z, q;
    double fx, fy, fz;
};

template <bool DIAG>
__device__ __forceinline__ void sym_phase(SymRow (&row)[kSymRows], const int (&gi)[kSymRows], const int gj0,
                                          const double *__restrict__ sx, const double *__restrict__ sy,
                                          const double *__restrict__ sz, const double *__restrict__ sq,
                                          double *__restrict__ sax, double *__restrict__ say, double *__restrict__ saz,
                                          const int lane, const double c2s, const double inv_lambda,
                                          const double *__restrict__ tab, int &min_r2_hi) {
#pragma unroll 1
    for (int s = 0; s < 32; s++) {
        const int col = (lane + s) & 31;
        const double xj = sx[col], yj = sy[col], zj = sz[col], qj = sq[col];
        double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
        for (int r = 0; r < kSymRows; r++) {
            double cx = xj, qc = qj, qr = row[r].q;
            if (DIAG) {  // inside the row block: count (i,j) once, j > i; everything else contributes exactly zero
                const bool on = (gj0 + col) > gi[r];
                cx = on ? xj : row[r].x - 1.0;  // keeps r finite for the self pair
                qc = on ? qj : 0.0;
                qr = on ? qr : 0.0;
            }
            const double dx = row[r].x - cx, dy = row[r].y - yj, dz = row[r].z - zj;
            const double t = pair_geom(dx, dy, dz, c2s, inv_lambda, tab, min_r2_hi);
            const double sj = qc * t, si = qr * t;
            row[r].fx = fma(sj, dx, row[r].fx);
            row[r].fy = fma(sj, dy, row[r].fy);
            row[r].fz = fma(sj, dz, row[r].fz);
            ax = fma(si, dx, ax);
            ay = fma(si, dy, ay);
            az = fma(si, dz, az);
        }
        sax[col] -= ax;
        say[col] -= ay;
        saz[col] -= az;
        __syncwarp();
    }
}

// grid = units of this rank (heaviest first); block = kPairThreads.
__global__ void __launch_bounds__(kPairThreads)
k_pair_sym(const double4 *__restrict__ xyzq, const int2 *__restrict__ units, double *__restrict__ rowpart,
           double *__restrict__ colpart, int *__restrict__ lj_hit, const SymParams P) {
    __shared__ double sx[kSymCT], sy[kSymCT], sz[kSymCT], sq[kSymCT];
    __shared__ double sax[kSymCT], say[kSymCT], saz[kSymCT];
    __shared__ double tab[kExpTab];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int k = t; k < kExpTab; k += kPairThreads) tab[k] = c_exp2_tab[k];
    const int2 unit = units[blockIdx.x];
    const int I = unit.x, c = unit.y;
    const int row0 = I * kSymRB;

    SymRow row[kSymRows];
    int gi[kSymRows];
#pragma unroll
    for (int r = 0; r < kSymRows; r++) {
        gi[r] = row0 + r * kPairThreads + t;
        const double4 v = xyzq[min(gi[r], P.n - 1)];
        row[r].x = v.x; row[r].y = v.y; row[r].z = v.z;
        row[r].q = gi[r] < P.n ? v.w : 0.0;  // rows past the end exert nothing and are never stored
        row[r].fx = row[r].fy = row[r].fz = 0.0;
    }
    const double c2s = P.c2s, inv_lambda = P.inv_lambda;
    int min_r2_hi = 0x7fffffff;
    const int n_tiles = (P.n + kSymCT - 1) / kSymCT;
    const int tile_end = min((c + 1) * P.chunk_tiles, n_tiles);
    const int tile_begin = max(c * P.chunk_tiles, row0 / kSymCT);  // tiles below the row block belong to other units

    for (int tile = tile_begin; tile < tile_end; tile++) {
        const int j0 = tile * kSymCT;
        const int j = j0 + t;
        __syncthreads();
        if (j < P.n) {
            const double4 v = xyzq[j];
            sx[t] = v.x; sy[t] = v.y; sz[t] = v.z; sq[t] = v.w;
        } else {  // padding columns: far away, no charge
            sx[t] = 1.0e3 + t; sy[t] = 0.0; sz[t] = 0.0; sq[t] = 0.0;
        }
        sax[t] = 0.0; say[t] = 0.0; saz[t] = 0.0;
        __syncthreads();
        const bool diag = j0 < row0 + kSymRB;
#pragma unroll 1
        for (int p = 0; p < 4; p++) {
            const int o = ((warp + p) & 3) * 32;
            if (!diag)
                sym_phase<false>(row, gi, j0 + o, sx + o, sy + o, sz + o, sq + o, sax + o, say + o, saz + o, lane, c2s,
                                 inv_lambda, tab, min_r2_hi);
            else
                sym_phase<true>(row, gi, j0 + o, sx + o, sy + o, sz + o, sq + o, sax + o, say + o, saz + o, lane, c2s,
                                inv_lambda, tab, min_r2_hi);
            __syncthreads();
        }
        if (j < P.n) {
            double *cp = colpart + (size_t) I * 3 * P.npad + j;
            cp[0] = sax[t];
            cp[P.npad] = say[t];
            cp[2 * (size_t) P.npad] = saz[t];
        }
    }
#pragma unroll
    for (int r = 0; r < kSymRows; r++) {
        if (gi[r] < P.n) {
            double *rp = rowpart + (size_t) c * 3 * P.npad + gi[r];
            rp[0] = row[r].fx;
            rp[P.npad] = row[r].fy;
            rp[2 * (size_t) P.npad] = row[r].fz;
        }
    }
    if (min_r2_hi <= P.lj_cut_hi) *lj_hit = 1;  // benign race: every writer stores the same value
}

// Per vertex x and virtual rank v: the partials of v's units that contain x, in a fixed order (row partials
// by ascending chunk, then column partials by ascending row block).  grid = (vertex blocks, local virtual
// ranks); unit_id[I * n_chunks + c] is the unit's index in the global enumeration (-1: no such unit).
__global__ void __launch_bounds__(128)
k_pair_reduce(const double *__restrict__ rowpart, const double *__restrict__ colpart, const int *__restrict__ unit_id,
              double *__restrict__ fv, int v_first, const SymParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = v_first + blockIdx.y;
    if (x >= P.n) return;
    const int Ix = x / kSymRB, cx = x / (P.chunk_tiles * kSymCT);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int c = cx; c < P.n_chunks; c++) {
        const int id = __ldg(unit_id + Ix * P.n_chunks + c);
        if (id >= 0 && (id % kSymV) == v) {
            const double *rp = rowpart + (size_t) c * 3 * P.npad + x;
            fx += __ldcg(rp); fy += __ldcg(rp + P.npad); fz += __ldcg(rp + 2 * (size_t) P.npad);
        }
    }
    for (int I = 0; I <= Ix; I++) {
        const int id = __ldg(unit_id + I * P.n_chunks + cx);
        if (id >= 0 && (id % kSymV) == v) {
            const double *cp = colpart + (size_t) I * 3 * P.npad + x;
            fx += __ldcg(cp); fy += __ldcg(cp + P.npad); fz += __ldcg(cp + 2 * (size_t) P.npad);
        }
    }
    double *o = fv + (size_t) v * 3 * P.npad + x;
    o[0] = fx;
    o[P.npad] = fy;
    o[2 * (size_t) P.npad] = fz;
}

}  // namespace npsl
