// Screened-Coulomb all-pairs kernel that evaluates every unordered pair once (Newton's third law),
// FP64, sm_100a. Force-only twin of k_pair (pair_kernel.cuh) for the per-step path; it replaces the
// same reference loop (src/newforces.cpp:153-177), which walks all N(N-1) ordered pairs.
//
// Work unit = (row block I of ROWS x THREADS rows) x (column chunk c of chunk_tiles x 128 columns); a unit exists
// when the chunk reaches past the start of the row block, and it covers the pairs (i,j), i in I, j in c,
// j > i. Each thread keeps ROWS rows in registers (position, charge, force accumulator). A THREADS-column tile is
// staged in shared memory (SoA) together with one force accumulator per column. Within a tile the
// warps work on the 32-column groups, rotating groups between phases; inside a phase lane l
// visits column (l + s) mod 32 at step s, so at any time every column is touched by exactly one thread:
// the column accumulators need no atomics and receive their contributions in a fixed order.
//   row partials    rowpart[c][k][i]   sum over the columns of chunk c        (k = x,y,z)
//   column partials colpart[I][k][j]   minus the sum over the rows of block I
// k_pair_reduce then adds, per vertex, the partials of each of 8 "virtual ranks" (a fixed function of the
// row block) in a fixed order, and k_vertex adds the 8 results: the value does not depend on which GPU computed which
// unit, so trajectories are bit-identical on 1, 2, 4 and 8 GPUs.
//
// FP64-pipe work per unordered pair: 32.75 instructions (20 for r, 1/r, exp; 4 for the geometry factor;
// 2 charge products; 6 FMA into the two accumulators; 3 adds per 4 pairs for the column update)
// = 16.4 per ordered pair, against 28 in k_pair (tools/sass_count.py).
#pragma once
#include "pair_kernel.cuh"
#include "xchg.cuh"

namespace npsl {

constexpr int kSymPad = 512;  // partial-buffer planes are padded to a multiple of this
constexpr int kSymV = 8;      // virtual ranks (fixed grouping of the units)

struct SymParams {
    double c2s, inv_lambda;
    int n;            // vertices
    int npad;         // plane length of the partial buffers (>= n, multiple of kSymPad)
    int rb;           // rows per row block = threads per CTA x rows per thread
    int chunk_tiles;  // 128-column tiles per column chunk
    int n_chunks;
    int lj_cut_hi;
};
constexpr int kSymCT = 128;   // granularity of the column chunks (the kernels' column tile is THREADS wide)

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
template <int ROWS, bool DIAG>
__device__ __forceinline__ void sym_phase(SymRow (&row)[ROWS], const int (&gi)[ROWS], const int gj0,
                                          const double *__restrict__ sx, const double *__restrict__ sy,
                                          const double *__restrict__ sz, const double *__restrict__ sq,
                                          double *__restrict__ sax, double *__restrict__ say, double *__restrict__ saz,
                                          const int lane, const double c2s, const double inv_lambda,
                                          const double *__restrict__ tab, int &min_r2_hi) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll 1
    for (int s = 0; s < 32; s++) {
        const int col = (lane + s) & 31;
        const double xj = sx[col], yj = sy[col], zj = sz[col], qj = sq[col];
        double dx[ROWS], dy[ROWS], dz[ROWS], qc[ROWS], qr[ROWS], a[ROWS], b[ROWS], y[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            double cx = xj;
            qc[r] = qj;
            qr[r] = row[r].q;
            if (DIAG) {  // inside the row block: count (i,j) once, j > i; everything else contributes exactly zero
                const bool on = (gj0 + col) > gi[r];
                cx = on ? xj : row[r].x - 1.0;  // keeps r finite for the self pair
                qc[r] = on ? qj : 0.0;
                qr[r] = on ? qr[r] : 0.0;
            }
            dx[r] = row[r].x - cx; dy[r] = row[r].y - yj; dz[r] = row[r].z - zj;
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) a[r] = fma(dz[r], dz[r], fma(dy[r], dy[r], dx[r] * dx[r]));  // r^2
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
        for (int r = 0; r < ROWS; r++) b[r] = fma(f[r], NPSL_EXP2_G4, NPSL_EXP2_G3);
#pragma unroll
        for (int r = 0; r < ROWS; r++) b[r] = fma(b[r], f[r], NPSL_EXP2_G2);
#pragma unroll
        for (int r = 0; r < ROWS; r++) b[r] = fma(b[r], f[r], NPSL_EXP2_G1);
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const double ex = fma(T[r], f[r] * b[r], T[r]);
            const double w = (y[r] * y[r]) * (y[r] + inv_lambda);
            a[r] = ex * w;  // e^{-r/lambda} (1/r^3 + 1/(lambda r^2))
        }
        double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const double sj = qc[r] * a[r], si = qr[r] * a[r];
            row[r].fx = fma(sj, dx[r], row[r].fx);
            row[r].fy = fma(sj, dy[r], row[r].fy);
            row[r].fz = fma(sj, dz[r], row[r].fz);
            ax = fma(si, dx[r], ax);
            ay = fma(si, dy[r], ay);
            az = fma(si, dz[r], az);
        }
        sax[col] -= ax;
        say[col] -= ay;
        saz[col] -= az;
        __syncwarp();
    }
}

// grid = units of this rank (heaviest first); block = THREADS; row block = ROWS * THREADS rows (== P.rb).
template <int ROWS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_pair_sym(const double4 *__restrict__ xyzq, const int2 *__restrict__ units, double *__restrict__ rowpart,
           double *__restrict__ colpart, int *__restrict__ lj_hit, const SymParams P) {
    constexpr int RB = ROWS * THREADS;
    constexpr int GROUPS = THREADS / 32;
    __shared__ double sx[THREADS], sy[THREADS], sz[THREADS], sq[THREADS];
    __shared__ double sax[THREADS], say[THREADS], saz[THREADS];
    __shared__ double tab[kExpTab];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int k = t; k < kExpTab; k += THREADS) tab[k] = c_exp2_tab[k];
    const int2 unit = units[blockIdx.x];
    const int I = unit.x, c = unit.y;
    const int row0 = I * RB;

    SymRow row[ROWS];
    int gi[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        gi[r] = row0 + r * THREADS + t;
        const double4 v = xyzq[min(gi[r], P.n - 1)];
        row[r].x = v.x; row[r].y = v.y; row[r].z = v.z;
        row[r].q = gi[r] < P.n ? v.w : 0.0;  // rows past the end exert nothing and are never stored
        row[r].fx = row[r].fy = row[r].fz = 0.0;
    }
    const double c2s = P.c2s, inv_lambda = P.inv_lambda;
    int min_r2_hi = 0x7fffffff;
    // columns of this unit: the chunk, from the start of the row block on (what lies below belongs to other units)
    const int col_begin = max(c * P.chunk_tiles * kSymCT, row0);
    const int col_end = min((c + 1) * P.chunk_tiles * kSymCT, P.n);

    for (int j0 = col_begin; j0 < col_end; j0 += THREADS) {
        const int j = j0 + t;
        __syncthreads();
        if (j < col_end) {
            const double4 v = xyzq[j];
            sx[t] = v.x; sy[t] = v.y; sz[t] = v.z; sq[t] = v.w;
        } else {  // padding columns: far away, no charge
            sx[t] = 1.0e3 + t; sy[t] = 0.0; sz[t] = 0.0; sq[t] = 0.0;
        }
        sax[t] = 0.0; say[t] = 0.0; saz[t] = 0.0;
        __syncthreads();
        const bool diag = j0 < row0 + RB;
#pragma unroll 1
        for (int p = 0; p < GROUPS; p++) {
            const int o = ((warp + p) % GROUPS) * 32;
            if (j0 + o < col_end) {  // whole group of padding columns: nothing to do
                if (!diag)
                    sym_phase<ROWS, false>(row, gi, j0 + o, sx + o, sy + o, sz + o, sq + o, sax + o, say + o, saz + o,
                                           lane, c2s, inv_lambda, tab, min_r2_hi);
                else
                    sym_phase<ROWS, true>(row, gi, j0 + o, sx + o, sy + o, sz + o, sq + o, sax + o, say + o, saz + o,
                                          lane, c2s, inv_lambda, tab, min_r2_hi);
            }
            __syncthreads();
        }
        if (j < col_end) {
            double *cp = colpart + (size_t) I * 3 * P.npad + j;
            cp[0] = sax[t];
            cp[P.npad] = say[t];
            cp[2 * (size_t) P.npad] = saz[t];
        }
    }
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
// block = (32 vertices, 8): thread (x, k) adds the column partials of virtual rank k and the row partials of
// the chunks congruent to k mod 8; the eight row segments are then added in order 0..7.
// Only the virtual ranks [v_first, v_first + v_count) are computed (the others live on other GPUs).
__global__ void __launch_bounds__(256)
k_pair_reduce(const double *__restrict__ rowpart, const double *__restrict__ colpart, double *__restrict__ fv,
              int v_first, int v_count, double *const *__restrict__ peer_fv, int rows_per_rank, const Xchg X,
              const SymParams P) {
    __shared__ double seg[3][8][32];
    const int tx = threadIdx.x, k = threadIdx.y;
    const int x = blockIdx.x * 32 + tx;
    const bool live = x < P.n;
    const int Ix = live ? x / P.rb : 0;
    const int vx = sym_vrank(Ix);
    const bool rows_here = live && vx >= v_first && vx < v_first + v_count;
    const bool cols_here = live && k >= v_first && k < v_first + v_count;
    double rx = 0.0, ry = 0.0, rz = 0.0;
    if (rows_here) {
        const int cmin = (Ix * P.rb) / (P.chunk_tiles * kSymCT);
#pragma unroll 4
        for (int c = cmin + ((k - cmin) & 7); c < P.n_chunks; c += 8) {  // chunks congruent to k mod 8, ascending
            const double *rp = rowpart + (size_t) c * 3 * P.npad + x;
            rx += __ldcg(rp); ry += __ldcg(rp + P.npad); rz += __ldcg(rp + 2 * (size_t) P.npad);
        }
    }
    seg[0][k][tx] = rx; seg[1][k][tx] = ry; seg[2][k][tx] = rz;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (cols_here) {
#pragma unroll 2
        for (int b = 0; b <= Ix; b += 16) {  // row blocks of virtual rank k: 16j + k and 16j + 15 - k
            const int I1 = b + k, I2 = b + 15 - k;
            if (I1 <= Ix) {
                const double *cp = colpart + (size_t) I1 * 3 * P.npad + x;
                fx += __ldcg(cp); fy += __ldcg(cp + P.npad); fz += __ldcg(cp + 2 * (size_t) P.npad);
            }
            if (I2 <= Ix) {
                const double *cp = colpart + (size_t) I2 * 3 * P.npad + x;
                fx += __ldcg(cp); fy += __ldcg(cp + P.npad); fz += __ldcg(cp + 2 * (size_t) P.npad);
            }
        }
    }
    __syncthreads();
    if (cols_here) {
        if (k == vx) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) { sx += seg[0][j][tx]; sy += seg[1][j][tx]; sz += seg[2][j][tx]; }
            fx = sx + fx; fy = sy + fy; fz = sz + fz;
        }
        // sharded over peer memory: the sum goes to the rank that owns vertex x, nobody else needs it
        double *o = (X.on ? peer_fv[x / rows_per_rank] : fv) + (size_t) k * 3 * P.npad + x;
        o[0] = fx;
        o[P.npad] = fy;
        o[2 * (size_t) P.npad] = fz;
    }
    if (X.on) {  // xchg_signal expects a 1-D block: fold the 2-D index
        __syncthreads();
        if (tx == 0 && k == 0) {
            __threadfence_system();
            const unsigned int t = atomicAdd(X.tickets + XK_FV, 1u);
            if (t == gridDim.x - 1) {
                X.tickets[XK_FV] = 0;
                __threadfence();
                const unsigned long long e = X.state[XK_FV] + 1;
                X.state[XK_FV] = e;
                __threadfence_system();
                for (int r = 0; r < X.world; r++) st_release_sys(X.peer_flags[r] + XK_FV * kMaxRanks + X.rank, e);
            }
        }
    }
}

}  // namespace npsl
