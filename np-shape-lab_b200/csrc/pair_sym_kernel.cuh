// Screened-Coulomb all-pairs kernel that evaluates every unordered pair once (Newton's third law),
// FP64, sm_100a. Force-only twin of k_pair (pair_kernel.cuh) for the per-step path; it replaces the
// same reference loop (src/newforces.cpp:153-177), which walks all N(N-1) ordered pairs.
//
// Work unit = (row block I of ROWS x THREADS rows) x (column chunk c of chunk_tiles x 32 columns); a unit exists
// when the chunk reaches past the start of the row block, and it covers the pairs (i,j), i in I, j in c,
// j > i. Each thread keeps ROWS rows in registers (position, charge, force accumulator). A THREADS-column tile is
// staged in shared memory as 16-byte pairs (x,y), (z,q), every group of 32 columns twice in a row. Within a tile the
// warps work on the 32-column groups, rotating groups between phases; inside a phase lane l visits column
// (l + s) mod 32 at step s, and the column's force accumulator travels with the column through registers (lane l hands
// it to lane l - 1 after its step): at any time every column is touched by exactly one thread, so the column sums
// need no atomics and receive their contributions in a fixed order.
//   row partials    rowpart[c][k][i]   sum over the columns of chunk c        (k = x,y,z)
//   column partials colpart[I][k][j]   minus the sum over the rows of block I
// k_pair_reduce then adds, per vertex, the partials of each of 8 "virtual ranks" (a fixed function of the
// row block) in a fixed order, and k_vertex adds the 8 results: the value does not depend on which GPU computed which
// unit, so trajectories are bit-identical on 1, 2, 4 and 8 GPUs.
//
// FP64-pipe work per unordered pair in the default form (SYM_FAST, 6 rows x 64 threads): 28 instructions (3 sub, 3 r^2,
// 14 for 1/r, r, exp and the radial factor, 2 charge products, 6 FMA into the two accumulators) = 14 per ordered pair,
// against 27 in k_pair; the loop is 221 instructions per 6 pairs (tools/sass_loops.py). What bounds it is the dispatch
// port of an SM sub-partition: every instruction of the loop that is not FP64 costs FP64 throughput,
// and the loop is sensitive to every register (243 of 255 in use).
#pragma once
#include "pair_kernel.cuh"
#include "xchg.cuh"
#include "exp2_table_fast.inc"

namespace npsl {

// How k_pair_sym evaluates the radial factor e^{-r/lambda} (1/r^3 + 1/(lambda r^2)) of a pair:
//   SYM_EXACT  rsqrt and exp to ~3e-16 (31.75 FP64-pipe instructions per pair);
//   SYM_POLY   piecewise degree-5 polynomial in r^2 from a 39 KB table (20.75; bound by shared-memory bandwidth);
//   SYM_FAST   the exact scheme with the accuracy the contract of the path does not need taken out: coordinates
//              pre-scaled by 1/lambda (the factor becomes e^{-r}(1/r^3 + 1/r^2), one FMA less), second-order instead of
//              third-order correction of the rsqrt seed (3/8 e^2 <= 4e-13), 2048-entry 2^(j/2048) table with a quadratic
//              (3.2e-13): 28 per pair, relative error of the factor < 1e-12 against a contract of 1e-10.
enum { SYM_EXACT = 0, SYM_POLY = 1, SYM_FAST = 2 };
constexpr int kExpTabFast = 1 << NPSL_EXP2FAST_TB;
constexpr int kSymPad = 512;  // partial-buffer planes are padded to a multiple of this
constexpr int kSymV = 8;      // virtual ranks (fixed grouping of the units)

struct SymParams {
    double c2s, inv_lambda;
    int n;            // vertices
    int npad;         // plane length of the partial buffers (>= n, multiple of kSymPad)
    int rb;           // rows per row block = threads per CTA x rows per thread
    int chunk_tiles;  // kSymCT-column tiles per column chunk (the first n_coarse chunks)
    int n_coarse;     // chunks [n_coarse, n_coarse + n_fine) further up the column range are fine_tiles wide, the chunks
    int fine_tiles;   //   [n_coarse + n_fine, n_chunks) at its high end finest_tiles wide: smaller units that are
    int n_fine;       //   scheduled last and even out the tail of the launch
    int finest_tiles;
    int n_chunks;
    int lj_cut_hi;
    double fast_c2s, fast_g1, fast_g2;  // SYM_FAST: +log2(e) 2^TB (applied to -r) and the quadratic of exp2_table_fast.inc, as kernel
                                        //   parameters so that they are constant-bank operands, not registers
    const double *etab;   // SYM_FAST: 2^(j/2^TB) with the high words biased by the index (k_pair_sym), 16-byte aligned
    const double2 *ptab;  // piecewise-polynomial table of the radial factor (kPoly* below), null: exact evaluation only
    unsigned long long *trace;  // measurement only (NPSL_SYM_TRACE): per CTA {SM id, start, rows and table there, pairs done, end} in ns, else null
};

// Piecewise polynomial of the whole radial factor  G(s) = e^{-sqrt(s)/lambda} (s^{-3/2} + s^{-1}/lambda),  s = r^2.
// The contract of the path is 1e-10 (north_star); evaluating rsqrt and exp to 3e-16 costs 17 of the 31.75 FP64
// instructions per pair. Here the interval comes from the exponent and the top kPolyBits mantissa bits of s (integer
// pipe), the argument is t = s - (midpoint of the interval) (exact, one DADD), and G is a degree-5 polynomial in t
// whose six coefficients were fitted per interval at context creation (Chebyshev nodes, long double; npsl.cu
// build_pair_table): 6 FP64 instructions instead of 17, no MUFU, relative error <= ~3e-12 in the near field
// (tests/test_pair_table.py). Covered range s in [2^kPolyEmin, 2^(kPolyEmin + kPolyOct)), i.e. r from 4.9e-4 to 5.66
// reduced units; a unit that meets a pair outside it (or any non-finite distance) is recomputed exactly by the same
// CTA, so the result stays a function of the unit's data alone (bit-identical on 1..8 GPUs).
constexpr int kPolyBits = 5;                        // 32 intervals per octave
constexpr int kPolyEmin = -21, kPolyOct = 26;
constexpr int kPolyCount = kPolyOct << kPolyBits;   // 832 intervals, 48 B each = 39 936 B of shared memory
constexpr int kPolyShift = 20 - kPolyBits;          // interval index = (high word of s >> kPolyShift) - kPolyBase
constexpr int kPolyBase = (1023 + kPolyEmin) << kPolyBits;
constexpr int kPolyHiMin = kPolyBase << kPolyShift; // high words of the covered range [min, max)
constexpr int kPolyHiMax = (kPolyBase + kPolyCount) << kPolyShift;
constexpr int kSymCT = 32;    // granularity of the column chunks (the kernels' column tile is THREADS wide)
// first tile of chunk c (c == n_chunks: one past the last tile, before clamping to the vertex count)
__host__ __device__ __forceinline__ int sym_chunk_tile0(const SymParams &P, const int c) {
    if (c <= P.n_coarse) return c * P.chunk_tiles;
    const int t1 = P.n_coarse * P.chunk_tiles, c1 = c - P.n_coarse;
    return c1 <= P.n_fine ? t1 + c1 * P.fine_tiles : t1 + P.n_fine * P.fine_tiles + (c1 - P.n_fine) * P.finest_tiles;
}
// chunk that holds tile t
__host__ __device__ __forceinline__ int sym_chunk_of_tile(const SymParams &P, const int t) {
    const int t1 = P.n_coarse * P.chunk_tiles;
    if (t < t1) return t / P.chunk_tiles;
    const int t2 = t1 + P.n_fine * P.fine_tiles;
    return t < t2 ? P.n_coarse + (t - t1) / P.fine_tiles : P.n_coarse + P.n_fine + (t - t2) / P.finest_tiles;
}

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

// One phase: this warp against one 32-column group. Lane l visits column (l + s) mod 32 at step s. The
// ROWS pair evaluations of a step are written stage by stage across the rows, so that the instruction
// stream offers ROWS independent FP64 chains (the dependent-issue latency of the FP64 pipe is ~13 cycles).
template <int ROWS, bool DIAG, int MATH, int UNROLL>
__device__ __forceinline__ void sym_phase(SymRow (&row)[ROWS], const int (&gi)[ROWS], const int gj0,
                                          const double2 *__restrict__ sxy, const double2 *__restrict__ szq,
                                          double2 *__restrict__ saxy, double *__restrict__ saz, const int lane, const double c2s, const double inv_lambda,
                                          const double *__restrict__ tab, const double2 *__restrict__ ptab, int &min_r2_hi,
                                          int &max_r2_hi, const double g1 = 0.0, const double g2 = 0.0) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    const unsigned tab32 = (unsigned) __cvta_generic_to_shared(tab);
    // The accumulator of the column a lane visits lives in registers and travels with the column: after its step the
    // lane hands it to lane - 1, which visits that column next (3 x 2 shuffles per step instead of two loads, two stores,
    // their addresses, three subtractions and a warp barrier). After the 32 steps lane l holds the sum of column l.
    double cax = 0.0, cay = 0.0, caz = 0.0;
    const int from = (lane + 1) & 31;
    // the group's 32 columns are staged twice in a row (entries [0, 63)), so that column (lane + s) mod 32 is entry
    // lane + s: one running address, no wrap-around arithmetic (every non-FP64 instruction of this loop costs a dispatch
    // slot of the FP64-bound sub-partition)
    const double2 *pxy = sxy + lane, *pzq = szq + lane;
#pragma unroll UNROLL
    for (int s = 0; s < 32; s++) {
        const double2 xy = pxy[s], zq = pzq[s];
        const double xj = xy.x, yj = xy.y, zj = zq.x, qj = zq.y;
        double dx[ROWS], dy[ROWS], dz[ROWS], qc[ROWS], qr[ROWS], a[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            double cx = xj;
            qc[r] = qj;
            qr[r] = row[r].q;
            if (DIAG) {  // inside the row block: count (i,j) once, j > i; everything else contributes exactly zero
                const bool on = (gj0 + ((lane + s) & 31)) > gi[r];
                cx = on ? xj : row[r].x - 1.0;  // keeps r finite for the self pair
                qc[r] = on ? qj : 0.0;
                qr[r] = on ? qr[r] : 0.0;
            }
            dx[r] = row[r].x - cx; dy[r] = row[r].y - yj; dz[r] = row[r].z - zj;
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) a[r] = fma(dz[r], dz[r], fma(dy[r], dy[r], dx[r] * dx[r]));  // r^2
        if (MATH == SYM_POLY) {
            double2 c01[ROWS], c23[ROWS], c45[ROWS];
            double t[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const int hi = __double2hiint(a[r]);
                min_r2_hi = min(min_r2_hi, hi);
                max_r2_hi = max(max_r2_hi, hi);
                const int k = min(max((hi >> kPolyShift) - kPolyBase, 0), kPolyCount - 1);
                const double2 *e = ptab + 3 * k;
                c01[r] = e[0]; c23[r] = e[1]; c45[r] = e[2];
                // midpoint of the interval: the kept mantissa bits, then 1000...0
                t[r] = a[r] - __hiloint2double((hi & (int) (0xffffffffu << kPolyShift)) | (1 << (kPolyShift - 1)), 0);
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = fma(c45[r].y, t[r], c45[r].x);
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = fma(a[r], t[r], c23[r].y);
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = fma(a[r], t[r], c23[r].x);
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = fma(a[r], t[r], c01[r].y);
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = fma(a[r], t[r], c01[r].x);  // e^{-r/lambda} (1/r^3 + 1/(lambda r^2))
        } else if (MATH == SYM_FAST) {  // coordinates are in units of lambda here
            double b[ROWS], y[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                min_r2_hi = min(min_r2_hi, __double2hiint(a[r]));
                y[r] = rsqrt_seed(a[r]);
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = a[r] * y[r];
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = fma(-b[r], y[r], 1.0);  // e = 1 - a y0^2
#pragma unroll
            for (int r = 0; r < ROWS; r++) y[r] = fma(y[r] * b[r], 0.5, y[r]);  // 1/r (1 - 3/8 e^2 ...)
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = -(a[r] * y[r]);  // -r: the scale c2p = log2(e) 2^TB then is a positive
            double tt[ROWS], f[ROWS], T[ROWS];                     // constant (an un-negated uniform operand of both FMAs)
#pragma unroll
            for (int r = 0; r < ROWS; r++) tt[r] = fma(a[r], c2s, MAGIC);
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                f[r] = fma(a[r], c2s, -(tt[r] - MAGIC));
                const int n = __double2loint(tt[r]);
                // entry n mod 2^TB: one and, one multiply-add for the shared-memory address (left to itself the compiler
                // shifts, masks and then adds the table's base: every integer instruction here costs a dispatch slot)
                double Tj;
                unsigned addr;
                asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(n & (kExpTabFast - 1)), "r"(tab32));
                asm("ld.shared.f64 %0, [%1];" : "=d"(Tj) : "r"(addr));
                // exponent k = n >> TB added on the integer pipe in ONE multiply-add: the staged table holds the high word of
                // 2^(j/2^TB) minus j << (20 - TB) (k_pair_sym), so  hi + (n << (20 - TB)) = hi(2^(j/2^TB)) + (k << 20).
                // No underflow guard: r < 700 lambda (npsl.cu checks the extent)
                T[r] = __hiloint2double(__double2hiint(Tj) + n * (1 << (20 - NPSL_EXP2FAST_TB)), __double2loint(Tj));
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = fma(f[r], g2, g1);
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const double ex = fma(T[r], f[r] * b[r], T[r]);
                const double y2 = y[r] * y[r];
                a[r] = ex * fma(y2, y[r], y2);  // e^{-r} (1/r^3 + 1/r^2)
            }
        } else {
            double b[ROWS], y[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                min_r2_hi = min(min_r2_hi, __double2hiint(a[r]));
                y[r] = rsqrt_seed(a[r]);
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = a[r] * y[r];
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = fma(-b[r], y[r], 1.0);  // e
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const double p = fma(b[r], 0.375, 0.5);
                y[r] = fma(y[r] * b[r], p, y[r]);  // 1/r
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) a[r] = a[r] * y[r];  // r
            double tt[ROWS], f[ROWS], T[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; r++) tt[r] = fma(a[r], c2s, MAGIC);
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                f[r] = fma(a[r], c2s, -(tt[r] - MAGIC));
                const int n = __double2loint(tt[r]);
                const double Tj = tab[n & (kExpTab - 1)];
                const int k = max(n >> NPSL_EXP2_TB, -1000);
                T[r] = __hiloint2double(__double2hiint(Tj) + (k << 20), __double2loint(Tj));
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = fma(f[r], NPSL_EXP2_G3, NPSL_EXP2_G2);
#pragma unroll
            for (int r = 0; r < ROWS; r++) b[r] = fma(b[r], f[r], NPSL_EXP2_G1);
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const double ex = fma(T[r], f[r] * b[r], T[r]);
                const double w = (y[r] * y[r]) * (y[r] + inv_lambda);
                a[r] = ex * w;  // e^{-r/lambda} (1/r^3 + 1/(lambda r^2))
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const double sj = qc[r] * a[r], si = qr[r] * a[r];
            row[r].fx = fma(sj, dx[r], row[r].fx);
            row[r].fy = fma(sj, dy[r], row[r].fy);
            row[r].fz = fma(sj, dz[r], row[r].fz);
            cax = fma(si, dx[r], cax);
            cay = fma(si, dy[r], cay);
            caz = fma(si, dz[r], caz);
        }
        cax = __shfl_sync(0xffffffffu, cax, from);
        cay = __shfl_sync(0xffffffffu, cay, from);
        caz = __shfl_sync(0xffffffffu, caz, from);
    }
    // the force on column j is minus the sum over the rows; the group's accumulators in shared memory collect the
    // phases of the tile (one warp per group and phase, __syncthreads between the phases)
    double2 a2 = saxy[lane];
    a2.x -= cax;
    a2.y -= cay;
    saxy[lane] = a2;
    saz[lane] -= caz;
}

// One pass over the column tiles of a unit. POLY: groups of 32 columns that are entirely real take the polynomial
// evaluation; the (single) partial group at the end of the column range, whose padding columns sit far away, takes the
// exact one. Returns through min / max of the high words of r^2 what the caller needs for the WCA detector and for
// the range check.
template <int ROWS, int THREADS, int MATH, int UNROLL>
__device__ __forceinline__ void sym_unit(SymRow (&row)[ROWS], const int (&gi)[ROWS], const double4 *__restrict__ xyzq,
                                         double *__restrict__ colpart, const int I, const int row0, const int col_begin,
                                         const int col_end, double2 *sxy, double2 *szq, double2 *saxy, double *saz,
                                         const double *__restrict__ tab,
                                         const double2 *__restrict__ ptab, const SymParams &P, int &min_r2_hi,
                                         int &max_r2_hi) {
    constexpr int RB = ROWS * THREADS;
    constexpr int GROUPS = THREADS / 32;
    // the exact evaluation used by SYM_POLY for partial groups / out-of-range units
    constexpr int EXACT = MATH == SYM_POLY ? SYM_EXACT : MATH;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const double inv_lambda = P.inv_lambda;
    // SYM_FAST works in units of lambda: the exponent scale no longer carries 1/lambda, the partial sums come out
    // lambda^2 too large (d and the radial factor scale by 1/lambda and lambda^3)
    const double c2s = MATH == SYM_FAST ? P.fast_c2s : P.c2s;
    const double cs = MATH == SYM_FAST ? inv_lambda : 1.0, out_scale = MATH == SYM_FAST ? inv_lambda * inv_lambda : 1.0;
#pragma unroll
    for (int r = 0; r < ROWS; r++) row[r].fx = row[r].fy = row[r].fz = 0.0;
    for (int j0 = col_begin; j0 < col_end; j0 += THREADS) {
        const int j = j0 + t;
        __syncthreads();
        {
            double2 cxy, czq;
            if (j < col_end) {
                const double4 v = xyzq[j];
                cxy = make_double2(v.x * cs, v.y * cs); czq = make_double2(v.z * cs, v.w);
            } else {  // padding columns: no charge, away from every vertex (SYM_FAST: 64 + t Debye lengths, within exp's range)
                cxy = make_double2(MATH == SYM_FAST ? 64.0 + t : 1.0e3 + t, 0.0); czq = make_double2(0.0, 0.0);
            }
            const int e = (t >> 5) * 64 + (t & 31);  // every group of 32 columns twice in a row (sym_phase)
            sxy[e] = cxy; szq[e] = czq;
            sxy[e + 32] = cxy; szq[e + 32] = czq;
        }
        saxy[t] = make_double2(0.0, 0.0); saz[t] = 0.0;
        __syncthreads();
        const bool diag = j0 < row0 + RB;
        // A tile with a single group of real columns (the 32-column chunks at the end of the column range, the smallest
        // units of the launch): both warps take that group in ONE phase instead of taking turns; warp 1's column sums go
        // to the slots of the empty second group and are added below in a fixed order.
        const bool single = GROUPS == 2 && col_end - j0 <= 32;
#pragma unroll 1
        for (int p = 0; p < (single ? 1 : GROUPS); p++) {
            const int o = single ? 0 : ((warp + p) % GROUPS) * 32;
            const int oa = single ? warp * 32 : o;
            if (j0 + o < col_end) {  // whole group of padding columns: nothing to do
                int dummy = 0;
                if (MATH == SYM_POLY && j0 + o + 32 <= col_end) {
                    if (!diag)
                        sym_phase<ROWS, false, SYM_POLY, UNROLL>(row, gi, j0 + o, sxy + 2 * o, szq + 2 * o, saxy + oa, saz + oa, lane, c2s, inv_lambda, tab, ptab,
                                                                 min_r2_hi, max_r2_hi);
                    else
                        sym_phase<ROWS, true, SYM_POLY, UNROLL>(row, gi, j0 + o, sxy + 2 * o, szq + 2 * o, saxy + oa, saz + oa, lane, c2s, inv_lambda, tab, ptab,
                                                                min_r2_hi, max_r2_hi);
                } else {
                    if (!diag)
                        sym_phase<ROWS, false, EXACT, UNROLL>(row, gi, j0 + o, sxy + 2 * o, szq + 2 * o, saxy + oa, saz + oa, lane, c2s, inv_lambda, tab, ptab, min_r2_hi, dummy, P.fast_g1, P.fast_g2);
                    else
                        sym_phase<ROWS, true, EXACT, UNROLL>(row, gi, j0 + o, sxy + 2 * o, szq + 2 * o, saxy + oa, saz + oa, lane, c2s, inv_lambda, tab, ptab, min_r2_hi, dummy, P.fast_g1, P.fast_g2);
                }
            }
            __syncthreads();
        }
        if (j < col_end) {
            double2 axy = saxy[t];
            double az = saz[t];
            if (single) { axy.x += saxy[t + 32].x; axy.y += saxy[t + 32].y; az += saz[t + 32]; }
            double *cp = colpart + (size_t) I * 3 * P.npad + j;
            cp[0] = axy.x * out_scale;
            cp[P.npad] = axy.y * out_scale;
            cp[2 * (size_t) P.npad] = az * out_scale;
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        row[r].fx *= out_scale; row[r].fy *= out_scale; row[r].fz *= out_scale;
    }
}

// grid = units of this rank (heaviest first); block = THREADS; row block = ROWS * THREADS rows (== P.rb).
// POLY: dynamic shared memory holds the polynomial table (kPolyCount x 48 B); the exact evaluation, needed for partial
// column groups and for the rare unit that leaves the table's range, then reads its 2^(j/1024) table from global memory.
template <int ROWS, int THREADS, int MINB, int MATH, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB)
k_pair_sym(const double4 *__restrict__ xyzq, const int2 *__restrict__ units, double *__restrict__ rowpart,
           double *__restrict__ colpart, int *__restrict__ lj_hit, const SymParams P) {
    constexpr int RB = ROWS * THREADS;
    extern __shared__ double2 s_ptab[];
    // one block of static shared memory with the exponential's table first: its entries sit at compile-time offsets
    struct __align__(16) Smem {
        double tab[MATH == SYM_POLY ? 2 : (MATH == SYM_FAST ? kExpTabFast : kExpTab)];
        double2 xy[2 * THREADS], zq[2 * THREADS], axy[THREADS];
        double az[THREADS];
        unsigned long long mbar;  // arrival of the table (SYM_FAST: one bulk copy)
    };
    __shared__ Smem sm;
    double *const s_tab = sm.tab;
    double2 *const sxy = sm.xy, *const szq = sm.zq, *const saxy = sm.axy;
    double *const saz = sm.az;
    const int t = threadIdx.x;
    unsigned long long t_begin = 0;
    if (P.trace && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    const double *tab = s_tab;
    if (MATH == SYM_POLY) {
        for (int k = t; k < 3 * kPolyCount; k += THREADS) s_ptab[k] = __ldg(P.ptab + k);
        tab = c_exp2_tab;
    } else if (MATH == SYM_FAST) {
        // the 16 KB table (high words biased by the index, see sym_phase; prepared once per context, P.etab) arrives by ONE
        // bulk asynchronous copy issued by thread 0 while all threads load their rows; a short unit (one 64-column tile,
        // ~30 us) otherwise spends several microseconds of its life in 32 dependent load / store rounds
        if (t == 0) {
            const unsigned mb = (unsigned) __cvta_generic_to_shared(&sm.mbar), dst = (unsigned) __cvta_generic_to_shared(sm.tab);
            constexpr unsigned bytes = kExpTabFast * sizeof(double);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(P.etab), "r"(bytes), "r"(mb) : "memory");
        }
    } else {
        for (int k = t; k < kExpTab; k += THREADS) s_tab[k] = c_exp2_tab[k];
    }
    const int2 unit = units[blockIdx.x];
    const int I = unit.x, c = unit.y;
    const int row0 = I * RB;
    const double cs = MATH == SYM_FAST ? P.inv_lambda : 1.0;

    SymRow row[ROWS];
    int gi[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        gi[r] = row0 + r * THREADS + t;
        const double4 v = xyzq[min(gi[r], P.n - 1)];
        row[r].x = v.x * cs; row[r].y = v.y * cs; row[r].z = v.z * cs;
        row[r].q = gi[r] < P.n ? v.w : 0.0;  // rows past the end exert nothing and are never stored
    }
    int min_r2_hi = 0x7fffffff, max_r2_hi = 0;
    // columns of this unit: the chunk, from the start of the row block on (what lies below belongs to other units)
    const int col_begin = max(sym_chunk_tile0(P, c) * kSymCT, row0);
    const int col_end = min(sym_chunk_tile0(P, c + 1) * kSymCT, P.n);

    unsigned long long t_ready = 0, t_done = 0;
    if (MATH == SYM_FAST) {  // the table has to be there (thread 0's mbarrier.init is ordered before by the barrier)
        __syncthreads();
        const unsigned mb = (unsigned) __cvta_generic_to_shared(&sm.mbar);
        asm volatile("{\n.reg .pred p;\nNPSL_TAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra NPSL_TAB_DONE;\n"
                     "bra NPSL_TAB_WAIT;\nNPSL_TAB_DONE:\n}" ::"r"(mb) : "memory");
    }
    if (P.trace && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ready));
    if (MATH == SYM_POLY) {
        sym_unit<ROWS, THREADS, SYM_POLY, UNROLL>(row, gi, xyzq, colpart, I, row0, col_begin, col_end, sxy, szq, saxy, saz, tab, s_ptab, P, min_r2_hi, max_r2_hi);
        // a pair outside the table (or a non-finite distance): the whole unit is redone exactly
        if (__syncthreads_or(min_r2_hi < kPolyHiMin || max_r2_hi >= kPolyHiMax))
            sym_unit<ROWS, THREADS, SYM_EXACT, UNROLL>(row, gi, xyzq, colpart, I, row0, col_begin, col_end, sxy, szq, saxy, saz, tab, s_ptab, P, min_r2_hi, max_r2_hi);
    } else {
        sym_unit<ROWS, THREADS, MATH, UNROLL>(row, gi, xyzq, colpart, I, row0, col_begin, col_end, sxy, szq, saxy, saz, tab, s_ptab, P, min_r2_hi, max_r2_hi);
    }
    if (P.trace && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_done));
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        if (gi[r] < P.n) {
            double *rp = rowpart + (size_t) c * 3 * P.npad + gi[r];
            rp[0] = row[r].fx;
            rp[P.npad] = row[r].fy;
            rp[2 * (size_t) P.npad] = row[r].fz;
        }
    }
    if (min_r2_hi <= P.lj_cut_hi) *lj_hit = 1;  // benign race: every writer stores the same value
    if (P.trace) {
        __syncthreads();
        if (t == 0) {
            unsigned long long t_end;
            unsigned int smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.trace[5 * blockIdx.x] = smid;
            P.trace[5 * blockIdx.x + 1] = t_begin;
            P.trace[5 * blockIdx.x + 2] = t_end;
            P.trace[5 * blockIdx.x + 3] = t_ready;
            P.trace[5 * blockIdx.x + 4] = t_done;
        }
    }
}

// Virtual rank of a row block: 0..7, 7..0, 0..7, ... so that the triangular work (row block I meets
// n_chunks - chunk(I) chunks) spreads evenly. All units of a row block belong to its virtual rank.
__host__ __device__ __forceinline__ int sym_vrank(int I) {
    const int m = I & 15;
    return m < 8 ? m : 15 - m;
}

// Sums the partials per vertex x and virtual rank v in a fixed order:
//   fv[v][x] = (v == vrank(block of x) ? sum of the row partials of x over its chunks : 0)
//              + sum over the row blocks I <= block(x) with vrank(I) == v of colpart[I][x]   (ascending I).
// Only the virtual ranks [v_first, v_first + v_count) are computed (the others live on other GPUs).
// A CTA has 8 slots of 32 lanes (lane = vertex of a group of 32 consecutive vertices):
//   CTAs [0, n_a): one group whose row block belongs to this rank -- slot k adds the row partials of the chunks
//     congruent to k mod 8 (the eight segments are then added in order 0..7) and, if virtual rank k is computed
//     here, the column partials of virtual rank k;
//   the others: 8 / v_count groups that only receive column partials -- slot s works on group s / v_count for
//     virtual rank v_first + s % v_count.
// groups[] lists the group of every (CTA, slot / v_count); the loads of a thread are issued in batches of 8 chunks /
// all row blocks so that a CTA needs only a few memory round trips (the kernel is latency-bound on a shard).
__global__ void __launch_bounds__(256)
k_pair_reduce(const double *__restrict__ rowpart, const double *__restrict__ colpart, double *__restrict__ fv,
              const int *__restrict__ groups, int n_a, int n_b_groups, int v_first, int v_count,
              double *const *__restrict__ peer_fv, int rows_per_rank, const Xchg X, const SymParams P) {
    __shared__ double seg[3][8][32];
    const int tx = threadIdx.x, slot = threadIdx.y;
    const bool remote = xchg_has(X, XK_FV);
    // receive buffers are double-buffered on the parity of the exchange epoch (read before this CTA's ticket,
    // i.e. before the last CTA can bump it)
    const size_t fv_off = remote ? (size_t) (X.state[XK_FV] & 1ull) * (kSymV * 3) * P.npad : 0;
    const bool part_a = (int) blockIdx.x < n_a;
    int g = -1, k = slot;
    if (part_a) {
        g = __ldg(groups + blockIdx.x);
    } else {
        const int per_cta = 8 / v_count;
        const int idx = ((int) blockIdx.x - n_a) * per_cta + slot / v_count;
        if (idx < n_b_groups) g = __ldg(groups + n_a + idx);
        k = v_first + slot % v_count;
    }
    const int x = g * 32 + tx;
    const bool live = g >= 0 && x < P.n;
    const int Ix = live ? x / P.rb : 0;
    const int vx = sym_vrank(Ix);
    const bool cols_here = live && k >= v_first && k < v_first + v_count;
    const size_t plane = P.npad;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (cols_here) {  // row blocks of virtual rank k: 16j + k and 16j + 15 - k, ascending
        constexpr int UB = 4;
        for (int b0 = 0; b0 <= Ix; b0 += 16 * UB) {
            double v[UB][2][3];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int I1 = b0 + 16 * u + k, I2 = b0 + 16 * u + 15 - k;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int I = h ? I2 : I1;
                    const bool on = I <= Ix;
                    const double *cp = colpart + (size_t) (on ? I : 0) * 3 * plane + x;
                    v[u][h][0] = on ? __ldcg(cp) : 0.0;
                    v[u][h][1] = on ? __ldcg(cp + plane) : 0.0;
                    v[u][h][2] = on ? __ldcg(cp + 2 * plane) : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < UB; u++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int I = h ? b0 + 16 * u + 15 - k : b0 + 16 * u + k;
                    if (I <= Ix) { fx += v[u][h][0]; fy += v[u][h][1]; fz += v[u][h][2]; }
                }
        }
    }
    if (part_a) {
        double rx = 0.0, ry = 0.0, rz = 0.0;
        if (live) {  // chunks congruent to k mod 8, ascending
            constexpr int UR = 8;
            const int cmin = sym_chunk_of_tile(P, (Ix * P.rb) / kSymCT);
            for (int c0 = cmin + ((k - cmin) & 7); c0 < P.n_chunks; c0 += 8 * UR) {
                double v[UR][3];
#pragma unroll
                for (int u = 0; u < UR; u++) {
                    const int c = c0 + 8 * u;
                    const bool on = c < P.n_chunks;
                    const double *rp = rowpart + (size_t) (on ? c : 0) * 3 * plane + x;
                    v[u][0] = on ? __ldcg(rp) : 0.0;
                    v[u][1] = on ? __ldcg(rp + plane) : 0.0;
                    v[u][2] = on ? __ldcg(rp + 2 * plane) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UR; u++)
                    if (c0 + 8 * u < P.n_chunks) { rx += v[u][0]; ry += v[u][1]; rz += v[u][2]; }
            }
        }
        seg[0][k][tx] = rx; seg[1][k][tx] = ry; seg[2][k][tx] = rz;
        __syncthreads();
        if (cols_here && k == vx) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) { sx += seg[0][j][tx]; sy += seg[1][j][tx]; sz += seg[2][j][tx]; }
            fx = sx + fx; fy = sy + fy; fz = sz + fz;
        }
    }
    if (cols_here) {
        const size_t off = fv_off + (size_t) k * 3 * plane + x;
        if (!remote) {
            double *o = fv + off;
            o[0] = fx; o[plane] = fy; o[2 * plane] = fz;
        } else if (!X.fv_bcast) {  // row-sharded plan: the sum goes to the rank that owns vertex x, nobody else needs it
            double *o = peer_fv[x / rows_per_rank] + off;
            o[0] = fx; o[plane] = fy; o[2 * plane] = fz;
        } else {  // replicated integrator: every rank adds up all 8 sums of every vertex
            for (int r = 0; r < X.world; r++) {
                double *o = peer_fv[r] + off;
                o[0] = fx; o[plane] = fy; o[2 * plane] = fz;
            }
        }
    }
    if (remote) {  // xchg_signal expects a 1-D block: fold the 2-D index
        __syncthreads();
        if (tx == 0 && slot == 0) xchg_signal_cta(X, XK_FV, gridDim.x);
    }
}

}  // namespace npsl
