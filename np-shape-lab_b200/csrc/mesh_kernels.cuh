// Mesh-elastic, integrator and thermostat kernels, FP64, sm_100a.
//
// These replace the serial O(N) loops around the pair term in the reference's step
// (src/newmd.cpp:96-170): the Nose-Hoover half-chains (src/functions.h:176-189), the two
// velocity half-kicks and the drift (src/vertex.h:105-120), INTERFACE::dressup
// (src/interface.cpp:33-67), the bending / stretching / surface-tension / volume-tension forces
// (src/edge.cpp:38-152, src/vertex.cpp:34-106), the force sum (src/newforces.cpp:279-280), the
// kinetic energy (src/newenergies.h:8-16) and the mesh part of INTERFACE::compute_energy
// (src/interface.cpp:990-1033).
//
// The reference walks pointer-linked VERTEX/EDGE/FACE objects and scatters bending forces into
// four vertices per edge through a std::map. Here the static topology is flattened once into
//   * one int4 per edge  (v0, v1, a0, a1): faces F0 = (v0,v1,a0) and F1 = (v1,v0,a1);
//   * per vertex, k-major tables of width NPSL_MAX_DEGREE: the counter-clockwise one-ring
//     ring[k][N] (face k of the vertex is (v, ring[k], ring[k+1])), the rest lengths l0 of the ring
//     edges, and the 2*degree bending-slot indices the vertex gathers;
// so every kernel is a coalesced sweep plus L2-resident gathers, with no atomics on data and a
// fixed summation order (run-to-run deterministic, independent of the row sharding).
//
// Step = k_kick_drift -> { pair kernels || k_mesh -> k_vertex_mesh } -> k_vertex<STEP>, captured in a CUDA Graph.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xchg.cuh"

namespace npsl {

constexpr int kMaxDeg = 8;     // == NPSL_MAX_DEGREE
constexpr int kMaxChain = 16;  // == NPSL_MAX_CHAIN
constexpr int kMeshThreads = 128;

struct BathLink {
    double Q, T, dof, xi, eta;
};
struct Thermo {
    BathLink mesh[kMaxChain];
    BathLink ions[kMaxChain];
};

// device scalar block
enum {
    SC_AREA = 0,  // total_area
    SC_VOL,       // total_volume
    SC_BE,        // bending energy
    SC_SE,        // stretching energy (before the 1/2 sconstant (avg^2) prefactor is applied: final value)
    SC_KE,        // kinetic energy of all vertices (all ranks)
    SC_ES,        // electrostatic energy of the owned rows
    SC_LJ,        // LJ energy of the owned rows
    SC_EXPFAC,    // exp(-dt/2 xi_0) of the current step
    SC_KICK,      // dt/2 sqrt(expfac)
    SC_LAMBDA,    // SHAKE multiplier of the last step (volume constraint)
    SC_MU,        // RATTLE multiplier of the last step
    SC_KE_IONS,      // kinetic energy of the explicit counterions (0 without ions)
    SC_EXPFAC_IONS,  // exp(-dt/2 xi_0) of the ion chain
    SC_KICK_IONS,    // dt/2 sqrt(.)
    SC_ES_IONS,      // ion-ion electrostatic energy
    SC_LJ_IONS,      // ion-ion LJ energy
    SC_COUNT = 16
};

struct MeshParams {
    int nv, ne, nf;
    int row_lo, row_hi;  // owned vertex range
    int n_ranks;
    int chain;
    int buckling;
    int n_splits, row_stride;  // layout of the pair partial sums
    int use_pair;              // 0 when every charge is zero (pair partials are not read)
    int pair_sym;              // 1: pair forces come from the 8 virtual-rank sums of k_pair_reduce
    int npad;                  // plane length of those sums
    int constraint;            // 1: rigid volume constraint (SHAKE / RATTLE, -G V)
    int n_ions;                // explicit counterions (ion_kernels.cuh); 0: none
    double dt;
    double pair_coef;  // scalefactor / em
    double bkappa, sconstant, sigma_a, sigma_v, ref_volume, avg_edge;
};

// ------------------------------------------------------------------------------------------------
// small vector helpers
struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 r = {x, y, z}; return r; }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
// The library is compiled with --fmad=false: every fused multiply-add is written out, so the
// rounding of each expression is fixed by the source (not by per-instantiation contraction
// choices of the compiler) and a force evaluated by k_vertex<FORCE> is bit-identical to the one
// k_vertex<STEP> used.
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ double dot(V3 a, V3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
// 32-byte records are moved as two 16-byte halves (there is no 32-byte __ldg/__ldcg overload)
__device__ __forceinline__ double4 ldg4(const double4 *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double4 ldcg4(const double4 *p) {
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ V3 ld3(const double4 *p, int i) {
    const double4 v = ldg4(p + i);
    return mk(v.x, v.y, v.z);
}

// Deterministic block sum of NV values per thread: warp shuffles in a fixed tree, then warp 0
// sums the per-warp results in warp order. Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* [NV][32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0;
            for (int w = 0; w < nwarp; w++) s += smem[k * 32 + w];
            v[k] = s;
        }
    }
}

// Last-block-done: every block publishes its partials, takes a ticket; the block that draws the
// last ticket sums all partials in block order (so the result does not depend on which block
// that is) and resets the ticket. Returns true in thread 0 of the last block.
__device__ __forceinline__ bool take_last_ticket(unsigned int *ticket, unsigned int nblocks) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    if (t == nblocks - 1) {
        *ticket = 0;
        __threadfence();
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Nose-Hoover chain, src/functions.h:176-189 (update_chain_xi), src/thermostat.h:53-58 (update_eta)
__device__ __forceinline__ void chain_xi(BathLink *b, int j, int chain, double dt, double ke) {
    if (b[j].Q == 0) return;
    const double xn = (j + 1 < chain) ? b[j + 1].xi : 0.0;
    const double drive = (j == 0) ? (2 * ke - b[j].dof * b[j].T) : (b[j - 1].Q * b[j - 1].xi * b[j - 1].xi - b[j].dof * b[j].T);
    b[j].xi = b[j].xi * exp(-0.5 * dt * xn) + 0.5 * dt * (1.0 / b[j].Q) * drive * exp(-0.25 * dt * xn);
}
__device__ __forceinline__ void chain_eta(BathLink *b, int chain, double dt) {
    for (int j = 0; j < chain; j++)
        if (b[j].Q != 0) b[j].eta = b[j].eta + 0.5 * dt * b[j].xi;
}
// Kinetic energy, summed the same way however the vertices are sharded: every warp of k_vertex
// leaves the sum of its 32 vertices (fixed shuffle tree) in ke_warp[global vertex index / 32]; the
// whole array (all ranks' warps, all-gathered when sharded) is then summed by one block with this
// fixed pattern. Owned ranges are multiples of 32, so the warp partials, and with them the
// thermostat and the trajectory, are bit-identical on 1 and on P GPUs. Result valid in thread 0.
__device__ __forceinline__ double canonical_sum(const double *__restrict__ p, int n, double *smem /* [32] */) {
    double v[1] = {0.0};
    for (int b = threadIdx.x; b < n; b += blockDim.x) v[0] += __ldcg(p + b);
    __syncthreads();
    block_sum<1>(v, smem);
    return v[0];
}

// One sweep of update_chain_xi over a chain held one link per lane (lanes [0,16) mesh chain, [16,32) ion chain;
// link j in lane j of its half). The arithmetic is the reference's, operation for operation:
//   xi_j <- xi_j exp(-dt/2 xi_{j+1}) + (dt/2 1/Q_j) drive_j exp(-dt/4 xi_{j+1}),
//   drive_0 = 2 KE - dof_0 T_0,   drive_j = Q_{j-1} xi_{j-1}^2 - dof_j T_j          (src/functions.h:176-189)
// FORWARD (j ascending, src/newmd.cpp:161-170): xi_{j+1} is still the old value, so both exponentials of every
// link are evaluated in parallel before the sweep and only the drive is serial. Reverse (j descending,
// src/newmd.cpp:99-110): xi_{j+1} has just been updated, so the exponentials are part of the serial chain.
template <bool FORWARD>
__device__ __forceinline__ void chain_sweep(BathLink &b, const int j, const int n, const double dt, const double ke) {
    const unsigned int full = 0xffffffffu;
    const bool real = j < n && b.Q != 0;
    const double c = real ? 0.5 * dt * (1.0 / b.Q) : 0.0;
    const double dofT = b.dof * b.T;
    double e1 = 1.0, e2 = 1.0;
    if (FORWARD) {
        double xn = __shfl_down_sync(full, b.xi, 1, 16);
        if (j + 1 >= n) xn = 0.0;
        if (real) { e1 = exp(-0.5 * dt * xn); e2 = exp(-0.25 * dt * xn); }
    }
    for (int it = 0; it < n; it++) {
        const int cur = FORWARD ? it : n - 1 - it;
        const double xp = __shfl_up_sync(full, b.xi, 1, 16), Qp = __shfl_up_sync(full, b.Q, 1, 16);
        double xn = __shfl_down_sync(full, b.xi, 1, 16);
        if (j == cur && real) {
            if (!FORWARD) {
                if (j + 1 >= n) xn = 0.0;
                e1 = exp(-0.5 * dt * xn);
                e2 = exp(-0.25 * dt * xn);
            }
            const double drive = (j == 0) ? (2 * ke - dofT) : (Qp * xp * xp - dofT);
            b.xi = b.xi * e1 + c * drive * e2;
        }
    }
}

// End of a force evaluation, called by warp 0 of the block that holds the total kinetic energy; lanes [0,16) carry
// the mesh chain, lanes [16,32) the ion chain (they are independent, src/newmd.cpp:99-110), one link per lane.
// Inside a step: forward half-chain, th_mid -> th_cur (src/newmd.cpp:161-170). Always: the reverse half-chain
// of the step that may follow from th_cur into th_mid, with exp(-dt/2 xi_0) and dt/2 sqrt(.) left in the
// scalar block, so that k_kick_drift is a pure streaming kernel.
__device__ __forceinline__ void thermo_finish(Thermo *th_mid, Thermo *th_cur, double *scal, const double ke_mesh,
                                              const bool step, const MeshParams &P) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, who = lane >> 4, j = lane & 15;  // who 0: mesh chain, 1: ion chain
    BathLink *mid = who == 0 ? th_mid->mesh : th_mid->ions;
    BathLink *cur = who == 0 ? th_cur->mesh : th_cur->ions;
    // the ion chain is driven by the ions' kinetic energy (k_ion_finish left it in the scalar block; 0 without ions)
    const double ke = who == 0 ? ke_mesh : __ldcg(scal + SC_KE_IONS);
    const int n = P.chain;
    const double dt = P.dt;
    BathLink b = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (j < n) b = step ? mid[j] : cur[j];
    if (step) {  // th_mid -> (eta, then xi forward) -> th_cur
        if (j < n && b.Q != 0) b.eta = b.eta + 0.5 * dt * b.xi;
        chain_sweep<true>(b, j, n, dt, ke);
        if (j < n) cur[j] = b;
    }
    chain_sweep<false>(b, j, n, dt, ke);
    if (j < n && b.Q != 0) b.eta = b.eta + 0.5 * dt * b.xi;
    if (j < n) mid[j] = b;
    if (j == 0) {  // lane 0: mesh chain, lane 16: ion chain (expfac_mesh / expfac_ions, src/newmd.cpp:109-110)
        const double ef = exp(-0.5 * dt * b.xi);
        scal[who == 0 ? SC_EXPFAC : SC_EXPFAC_IONS] = ef;
        scal[who == 0 ? SC_KICK : SC_KICK_IONS] = 0.5 * dt * sqrt(ef);
    }
}

// After npsl_set_thermostat (annealing): th_cur changed, prepare th_mid and the kick scalars again.
__global__ void k_thermo_prepare(Thermo *th_cur, Thermo *th_mid, double *scal, const MeshParams P) {
    if (blockIdx.x == 0) thermo_finish(th_mid, th_cur, scal, scal[SC_KE], false, P);
}

// ------------------------------------------------------------------------------------------------
// k_kick_drift: first half-kick and drift of the owned vertices, src/newmd.cpp:114-123. The reverse
// half-chain of this step was already done by thermo_finish at the end of the previous force evaluation.
// Packed host-state step (npsl_md_step_packed): pk_in = [pos | vel | force] as packed [n][3] doubles that were just
// copied from the caller's pinned block; pk_out = [pos | vel | force | KE], the block that goes back. Null: the state
// is resident in the records.
struct Packed {
    const double *in;
    double *out;
};

__global__ void __launch_bounds__(kMeshThreads)
k_kick_drift(double4 *__restrict__ xyzq, double4 *__restrict__ vel, const double4 *__restrict__ force,
             const double *__restrict__ scal, int *__restrict__ lj_hit, double4 *const *__restrict__ peer_xyzq,
             const Xchg X, const MeshParams P, const Packed K) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *lj_hit = 0;
    if (K.in && xchg_has(X, XK_STATE)) {  // the packed image arrives in slices, one from every rank (k_state_share)
        if (threadIdx.x < 32) xchg_wait_own_epoch(X, XK_STATE);
        __syncthreads();
    }
    const int i = P.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.row_hi) {
        const double ef = __ldcg(scal + SC_EXPFAC), kick = __ldcg(scal + SC_KICK), dt = P.dt;
        double4 v, f, x;
        if (K.in) {  // (P.nv is the plane length of the packed block)
            const double *px = K.in + 3 * (size_t) i, *pv = px + 3 * (size_t) P.nv, *pf = pv + 3 * (size_t) P.nv;
            x = make_double4(__ldcg(px), __ldcg(px + 1), __ldcg(px + 2), xyzq[i].w);
            v = make_double4(__ldcg(pv), __ldcg(pv + 1), __ldcg(pv + 2), 0.0);
            f = make_double4(__ldcg(pf), __ldcg(pf + 1), __ldcg(pf + 2), 0.0);
        } else {
            v = vel[i];
            f = force[i];
            x = xyzq[i];
        }
        v.x = fma(v.x, ef, f.x * kick);
        v.y = fma(v.y, ef, f.y * kick);
        v.z = fma(v.z, ef, f.z * kick);
        x.x = fma(v.x, dt, x.x);
        x.y = fma(v.y, dt, x.y);
        x.z = fma(v.z, dt, x.z);
        vel[i] = v;
        if (xchg_has(X, XK_POS)) {  // the new position goes straight into every rank's copy (NVLink peer stores)
            for (int r = 0; r < X.world; r++) peer_xyzq[r][i] = x;
        } else {
            xyzq[i] = x;
        }
        if (K.out) {
            double *po = K.out + 3 * (size_t) i;
            po[0] = x.x; po[1] = x.y; po[2] = x.z;
        }
    }
    if (xchg_has(X, XK_POS)) xchg_signal(X, XK_POS, gridDim.x);
}

// ------------------------------------------------------------------------------------------------
// k_mesh: blocks [0, n_face_blocks) sweep faces (INTERFACE::dressup: direction, area, volume and
// their totals, src/interface.cpp:33-67, src/face.cpp:16-45); the remaining blocks sweep edges
// (EDGE::bending_forces, src/edge.cpp:38-58, with compute_gradS :60-152 in its closed form
//   grad_p S = sum_{F in {F0,F1}, p in F} (pn_F - pnn_F) x dir_other(F),
// and the bending / stretching energies of src/interface.cpp:990-1015).
//
// Per edge with oriented record (v0,v1,a0,a1):  dir0 = (p1-p0)x(pa-p1), dir1 = (p0-p1)x(pb-p0),
// S = dir0.dir1, |dir| = 2A. With  w0 = dir1 - S/|dir0|^2 dir0,  w1 = dir0 - S/|dir1|^2 dir1  and
// kb = bkappa/(|dir0||dir1|) = bkappa 0.25/(A0 A1), the four bending-force slots are
//   v0: kb[(p1-pa) x w0 + (pb-p1) x w1]    v1: kb[(pa-p0) x w0 + (p0-pb) x w1]
//   a0: kb (p0-p1) x w0                    a1: kb (p1-p0) x w1
// which is  bkappa 0.25/(A0A1) (grad S - S/A0 grad A_F0 - S/A1 grad A_F1)  of the reference with
// grad_c A_F = (pn-pnn) x dir_F / (4 A_F)  (src/face.cpp:69-90).
template <bool ENERGY, bool STORE>
__global__ void __launch_bounds__(kMeshThreads)
k_mesh(const double4 *__restrict__ xyzq, const int4 *__restrict__ face_v, const int4 *__restrict__ edge_rec,
       const double *__restrict__ l0, double4 *__restrict__ bslot, double *__restrict__ partials /* [4][nblocks] */,
       unsigned int *__restrict__ ticket, double *__restrict__ scal, double4 *__restrict__ face_dir_area,
       double *__restrict__ face_vol, double *__restrict__ edge_S, double2 *__restrict__ edge_E, int n_face_blocks,
       const MeshParams P) {
    __shared__ double red[4 * 32];
    __shared__ bool s_last;
    const int nblocks = gridDim.x;
    double acc[4] = {0, 0, 0, 0};  // area, volume, bending energy, stretching energy
    if ((int) blockIdx.x < n_face_blocks) {
        const int f = blockIdx.x * blockDim.x + threadIdx.x;
        if (f < P.nf) {
            const int4 fv = __ldg(face_v + f);
            const V3 a = ld3(xyzq, fv.x), b = ld3(xyzq, fv.y), c = ld3(xyzq, fv.z);
            const V3 dir = cross(sub(b, a), sub(c, b));
            const double area = 0.5 * sqrt(dot(dir, dir));
            const double vol = (1. / 6.) * dot(a, cross(b, c));
            acc[0] = area;
            acc[1] = vol;
            if (STORE) {
                face_dir_area[f] = make_double4(dir.x, dir.y, dir.z, area);
                face_vol[f] = vol;
            }
        }
    } else {
        const int e = (blockIdx.x - n_face_blocks) * blockDim.x + threadIdx.x;
        if (e < P.ne) {
            const int4 r = __ldg(edge_rec + e);
            const V3 p0 = ld3(xyzq, r.x), p1 = ld3(xyzq, r.y), pa = ld3(xyzq, r.z), pb = ld3(xyzq, r.w);
            const V3 e01 = sub(p1, p0);
            const V3 dir0 = cross(e01, sub(pa, p1));
            const V3 dir1 = cross(sub(p0, p1), sub(pb, p0));
            const double S = dot(dir0, dir1);
            const double n0 = dot(dir0, dir0), n1 = dot(dir1, dir1);
            const double inv_n0n1 = 1.0 / sqrt(n0 * n1);
            const double kb = P.bkappa * inv_n0n1;
            const double c0 = S / n0, c1 = S / n1;
            const V3 w0 = mk(dir1.x - c0 * dir0.x, dir1.y - c0 * dir0.y, dir1.z - c0 * dir0.z);
            const V3 w1 = mk(dir0.x - c1 * dir1.x, dir0.y - c1 * dir1.y, dir0.z - c1 * dir1.z);
            const V3 t00 = cross(sub(p1, pa), w0), t01 = cross(sub(pb, p1), w1);
            const V3 t10 = cross(sub(pa, p0), w0), t11 = cross(sub(p0, pb), w1);
            const V3 ta = cross(sub(p0, p1), w0), tb = cross(e01, w1);
            bslot[e] = make_double4(kb * (t00.x + t01.x), kb * (t00.y + t01.y), kb * (t00.z + t01.z), 0.0);
            bslot[P.ne + e] = make_double4(kb * (t10.x + t11.x), kb * (t10.y + t11.y), kb * (t10.z + t11.z), 0.0);
            bslot[2 * P.ne + e] = make_double4(kb * ta.x, kb * ta.y, kb * ta.z, 0.0);
            bslot[3 * P.ne + e] = make_double4(kb * tb.x, kb * tb.y, kb * tb.z, 0.0);
            if (STORE) {
                edge_S[e] = S;
                // per-edge energies of the local maps (INTERFACE::compute_local_energies, src/interface.cpp:1198-1209):
                // bending, and stretching always in its l0 form whatever -B says
                const double le = __ldg(l0 + e), st = sqrt(dot(e01, e01)) - le;
                edge_E[e] = make_double2(P.bkappa * (1.0 - S * inv_n0n1),
                                         (0.5 * P.sconstant * st * st / (le * le)) * (P.avg_edge * P.avg_edge));
            }
            if (ENERGY) {
                acc[2] = P.bkappa * (1.0 - S * inv_n0n1);
                const double l = sqrt(dot(e01, e01));
                if (!P.buckling) {
                    const double le = __ldg(l0 + e), st = l - le;
                    acc[3] = 0.5 * P.sconstant * st * st / (le * le);
                } else {
                    const double st = l - P.avg_edge;
                    acc[3] = 0.5 * P.sconstant * st * st;
                }
            }
        }
    }
    block_sum<4>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) partials[k * nblocks + blockIdx.x] = acc[k];
        s_last = take_last_ticket(ticket, nblocks);
    }
    __syncthreads();
    if (!s_last) return;
    // ordered final sums, one warp-strided pass per quantity, fixed order
    double tot[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 4; k++) tot[k] += __ldcg(partials + k * nblocks + b);
    }
    __syncthreads();
    block_sum<4>(tot, red);
    if (threadIdx.x == 0) {
        scal[SC_AREA] = tot[0];
        scal[SC_VOL] = tot[1];
        if (ENERGY) {
            scal[SC_BE] = tot[2];
            scal[SC_SE] = P.buckling ? tot[3] : tot[3] * P.avg_edge * P.avg_edge;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_vertex_mesh: gather-by-vertex of the mesh-elastic force terms of the owned vertices. Runs on the side
// stream right after k_mesh, i.e. concurrently with the pair kernel, and leaves their sum in fmesh.
//   bending    sum of the vertex's 2*degree slots                     src/edge.cpp:38-58
//   stretching -s avg^2 sum_e (l-l0)/l (x_v-x_o)/l0^2  (or -B y form)  src/vertex.cpp:56-89
//   tension    sigma_a * sum_F -grad A_F ; 2 sigma_v (V-V0) sum_F -grad V_F   src/vertex.cpp:34-53,92-106
// k_vertex: pair term + fmesh = forvec (src/newforces.cpp:279-280), then (STEP) the second half-kick
//   v <- v expfac + f dt/2 sqrt(expfac) and ke = 1/2 v.v (src/newmd.cpp:146-157) and the thermostat, or
//   (FORCE) the pair force component and the pair energies.
enum { VMODE_STEP = 0, VMODE_FORCE = 1 };

struct VertexTables {
    const int *degree;      // [nv]
    const int *ring;        // [kMaxDeg][nv]   -1 padded
    const double *ring_l0;  // [kMaxDeg][nv]
    const int *bslot_ref;   // [2*kMaxDeg][nv] -1 padded
};

struct ForceComponents {  // FORCE mode only
    double4 *pair, *bend, *stretch, *tension, *voltension;
    double2 *vertex_e;  // (electrostatic, LJ) energy of the vertex: half its sum over all partners
};

template <int MODE>
__global__ void __launch_bounds__(kMeshThreads)
k_vertex_mesh(const double4 *__restrict__ xyzq, const double4 *__restrict__ bslot, const VertexTables T,
              const ForceComponents C, const double *__restrict__ scal, double4 *__restrict__ fmesh,
              double4 *__restrict__ ngv /* VERTEX::neg_GradVi, only with the volume constraint */, const MeshParams P) {
    const int i = P.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.row_hi) return;
    const int nv = P.nv;
    const V3 p = ld3(xyzq, i);
    // ---- bending: gather the slots written by k_mesh
    V3 fb = mk(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 2 * kMaxDeg; k++) {
        const int s = __ldg(T.bslot_ref + k * nv + i);
        if (s >= 0) {
            const double4 a = ldcg4(bslot + s);
            fb.x += a.x; fb.y += a.y; fb.z += a.z;
        }
    }
    // ---- one-ring: stretching, surface tension, volume tension
    V3 fs = mk(0, 0, 0), gA = mk(0, 0, 0), gV = mk(0, 0, 0);
    const int deg = __ldg(T.degree + i);
    const V3 pfirst = ld3(xyzq, __ldg(T.ring + i));
    V3 pn = pfirst;
#pragma unroll
    for (int k = 0; k < kMaxDeg; k++) {
        if (k >= deg) break;
        // pn = position of ring[k]; pnn = position of ring[k+1] (cyclic)
        const V3 pnn = (k + 1 < deg) ? ld3(xyzq, __ldg(T.ring + (k + 1) * nv + i)) : pfirst;
        // stretching along edge (v, ring[k])
        const V3 d = sub(p, pn);
        const double l = sqrt(dot(d, d));
        double c;
        if (!P.buckling) {
            const double le = __ldg(T.ring_l0 + k * nv + i);
            c = ((l - le) / l) * (1.0 / (le * le));
        } else {
            c = (l - P.avg_edge) / l;
        }
        fs.x += c * d.x; fs.y += c * d.y; fs.z += c * d.z;
        // face (p, pn, pnn): FACE::computegradients, src/face.cpp:47-91
        const V3 cv = cross(pn, pnn);
        gV.x -= (1. / 6.) * cv.x; gV.y -= (1. / 6.) * cv.y; gV.z -= (1. / 6.) * cv.z;
        const V3 cp = cross(sub(pn, p), sub(pnn, p));
        const V3 num = cross(sub(pn, pnn), cp);
        const double inv_den = 1.0 / (2.0 * sqrt(dot(cp, cp)));
        gA.x -= num.x * inv_den; gA.y -= num.y * inv_den; gA.z -= num.z * inv_den;
        pn = pnn;
    }
    const double ks = P.buckling ? -P.sconstant : -P.sconstant * P.avg_edge * P.avg_edge;
    fs.x *= ks; fs.y *= ks; fs.z *= ks;
    const double kv = 2.0 * P.sigma_v * (__ldcg(scal + SC_VOL) - P.ref_volume);
    const V3 ft = mk(P.sigma_a * gA.x, P.sigma_a * gA.y, P.sigma_a * gA.z);
    const V3 fv = mk(kv * gV.x, kv * gV.y, kv * gV.z);
    fmesh[i] = make_double4(fb.x + fs.x + ft.x + fv.x, fb.y + fs.y + ft.y + fv.y, fb.z + fs.z + ft.z + fv.z, 0.0);
    if (P.constraint) ngv[i] = make_double4(gV.x, gV.y, gV.z, 0.0);
    if (MODE == VMODE_FORCE) {
        C.bend[i] = make_double4(fb.x, fb.y, fb.z, 0.0);
        C.stretch[i] = make_double4(fs.x, fs.y, fs.z, 0.0);
        C.tension[i] = make_double4(ft.x, ft.y, ft.z, 0.0);
        C.voltension[i] = make_double4(fv.x, fv.y, fv.z, 0.0);
    }
}

// Tail shared by k_vertex and k_rattle_apply: per-warp kinetic-energy partials, (FORCE) pair-energy sums,
// exchange signal, and in the last block the total kinetic energy and the thermostat.
template <bool ENERGIES, bool STEP>
__device__ __forceinline__ bool vertex_tail(double (&acc)[3], double *__restrict__ partials, unsigned int *__restrict__ ticket,
                                            double *__restrict__ scal, double *__restrict__ ke_warp, int n_ke_warps,
                                            Thermo *__restrict__ th_mid, Thermo *__restrict__ th_cur,
                                            double *const *__restrict__ peer_ke, const Xchg &X, const MeshParams &P,
                                            double *red /* [3*32] */, bool *s_last_p) {
    bool &s_last = *s_last_p;
    const int nblocks = gridDim.x;
    // kinetic energy: one partial per warp at its global warp index (see canonical_sum)
    {
        double ke = acc[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ke += __shfl_down_sync(0xffffffffu, ke, o);
        const int i0 = P.row_lo + blockIdx.x * blockDim.x + (threadIdx.x & ~31);
        if ((threadIdx.x & 31) == 0 && i0 < P.row_hi) {
            if (xchg_has(X, XK_KE)) {
                for (int r = 0; r < X.world; r++) peer_ke[r][i0 >> 5] = ke;
            } else {
                ke_warp[i0 >> 5] = ke;
            }
        }
    }
    if (ENERGIES) {  // pair energies of the owned rows (summed over ranks by the host, rank order)
        double e2[2] = {acc[1], acc[2]};
        block_sum<2>(e2, red);
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = e2[0];
            partials[nblocks + blockIdx.x] = e2[1];
        }
    }
    if (xchg_has(X, XK_KE)) xchg_signal(X, XK_KE, nblocks);
    __syncthreads();  // every warp's ke_warp store precedes thread 0's fence + ticket
    if (threadIdx.x == 0) s_last = take_last_ticket(ticket, nblocks);
    __syncthreads();
    if (!s_last) return false;
    if (ENERGIES) {
        double tot[2] = {0, 0};
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            tot[0] += __ldcg(partials + b);
            tot[1] += __ldcg(partials + nblocks + b);
        }
        __syncthreads();
        block_sum<2>(tot, red);
        if (threadIdx.x == 0) {
            scal[SC_ES] = tot[0];
            scal[SC_LJ] = tot[1];
        }
        __syncthreads();
    }
    if (P.n_ranks > 1) return false;  // sharded: k_ke_post finishes after the exchange of ke_warp
    const double ke = canonical_sum(ke_warp, n_ke_warps, red);
    if (threadIdx.x == 0) {
        scal[SC_KE] = ke;
        red[0] = ke;
    }
    __syncthreads();
    thermo_finish(th_mid, th_cur, scal, red[0], STEP, P);
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(kMeshThreads)
k_vertex(const double4 *__restrict__ xyzq, double4 *__restrict__ vel, double4 *__restrict__ force,
         const double4 *__restrict__ fmesh, const double4 *__restrict__ pair_partial,
         const double *__restrict__ pair_fv, const double4 *__restrict__ lj_out, const int *__restrict__ lj_hit,
         const ForceComponents C, double *__restrict__ partials /* [2][nblocks] */, unsigned int *__restrict__ ticket,
         double *__restrict__ scal, double *__restrict__ ke_warp /* [n_ke_warps] */, int n_ke_warps,
         Thermo *__restrict__ th_mid, Thermo *__restrict__ th_cur, double *const *__restrict__ peer_ke,
         const Xchg X, const MeshParams P, const double4 *__restrict__ ion_mesh, const double *__restrict__ ion_mesh_lj,
         const Packed K) {
    __shared__ double red[3 * 32];
    __shared__ bool s_last;
    __shared__ size_t s_fv_off;
    const int i = P.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0, 0, 0};  // ke, es energy, lj energy
    // Sums of the virtual ranks that arrive over peer memory: wait here (one warp per CTA polls the arrival
    // flags of this force evaluation's exchange) instead of in a kernel of its own.
    if (P.use_pair && P.pair_sym && xchg_has(X, XK_FV)) {
        if (threadIdx.x < 32) {
            const unsigned long long e = xchg_wait_own_epoch(X, XK_FV);
            if (threadIdx.x == 0) s_fv_off = (size_t) ((e - 1) & 1ull) * (8 * 3) * P.npad;
        }
        __syncthreads();
        pair_fv += s_fv_off;
    }
    if (i < P.row_hi) {
        // ---- pair term: fixed-order sum of the partial results
        V3 fp = mk(0, 0, 0);
        double ue = 0;
        if (P.use_pair) {
            if (P.pair_sym) {  // the 8 virtual ranks (pair_sym_kernel.cuh)
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    const double *f = pair_fv + (size_t) v * 3 * P.npad + i;
                    fp.x += __ldcg(f); fp.y += __ldcg(f + P.npad); fp.z += __ldcg(f + 2 * (size_t) P.npad);
                }
                if (MODE == VMODE_FORCE)
                    for (int s = 0; s < P.n_splits; s++)
                        ue += ldcg4(pair_partial + (size_t) s * P.row_stride + (i - P.row_lo)).w;
            } else {  // the column chunks of the ordered kernel (pair_kernel.cuh)
                for (int s = 0; s < P.n_splits; s++) {
                    const double4 a = ldcg4(pair_partial + (size_t) s * P.row_stride + (i - P.row_lo));
                    fp.x += a.x; fp.y += a.y; fp.z += a.z; ue += a.w;
                }
            }
            const double cq = P.pair_coef * __ldcg(&xyzq[i].w);
            fp.x *= cq; fp.y *= cq; fp.z *= cq;
            ue *= 0.5 * cq;
        }
        double ulj = 0;
        if (!P.use_pair || __ldcg(lj_hit) != 0) {
            const double4 a = ldcg4(lj_out + (i - P.row_lo));
            fp.x += a.x; fp.y += a.y; fp.z += a.z; ulj = a.w;
        }
        double ue_ion = 0, ulj_ion = 0;
        if (P.n_ions > 0) {  // force of the counterions on the vertex and its mesh-ion energies: k_ion_mesh's ion tiles, in order
            double ax = 0, ay = 0, az = 0;
            for (int tile = 0; tile < (P.n_ions + 127) / 128; tile++) {
                const double4 a = ldcg4(ion_mesh + (size_t) tile * P.nv + i);
                ax += a.x; ay += a.y; az += a.z;
                ue_ion += a.w;
                ulj_ion += __ldcg(ion_mesh_lj + (size_t) tile * P.nv + i);
            }
            fp.x += ax; fp.y += ay; fp.z += az;
        }
        const double4 fm = ldcg4(fmesh + i);
        const V3 f = mk(fp.x + fm.x, fp.y + fm.y, fp.z + fm.z);
        force[i] = make_double4(f.x, f.y, f.z, 0.0);
        double4 v = vel[i];
        if (MODE == VMODE_STEP) {
            const double ef = __ldcg(scal + SC_EXPFAC), kick = __ldcg(scal + SC_KICK);
            v.x = fma(v.x, ef, f.x * kick);
            v.y = fma(v.y, ef, f.y * kick);
            v.z = fma(v.z, ef, f.z * kick);
            vel[i] = v;
            if (K.out) {  // packed host-state step: velvec and forvec go straight into the block that is copied back
                double *pv = K.out + 3 * ((size_t) P.nv + i), *pf = pv + 3 * (size_t) P.nv;
                pv[0] = v.x; pv[1] = v.y; pv[2] = v.z;
                pf[0] = f.x; pf[1] = f.y; pf[2] = f.z;
            }
        } else {
            C.pair[i] = make_double4(fp.x, fp.y, fp.z, 0.0);
            C.vertex_e[i] = make_double2(ue, ulj);  // mesh-mesh only: what the local energy maps use
            acc[1] = ue + ue_ion;
            acc[2] = ulj + ulj_ion;
        }
        acc[0] = 0.5 * (v.x * v.x + v.y * v.y + v.z * v.z);
    }
    if (MODE == VMODE_STEP && P.constraint) return;  // RATTLE follows: k_rattle_apply finishes the step
    const bool done = vertex_tail<MODE == VMODE_FORCE, MODE == VMODE_STEP>(acc, partials, ticket, scal, ke_warp, n_ke_warps,
                                                                            th_mid, th_cur, peer_ke, X, P, red, &s_last);
    if (MODE == VMODE_STEP && K.out && done && threadIdx.x == 0) K.out[9 * (size_t) P.nv] = red[0];  // kinetic energy
}

// Sharded runs: after the all-gather of every rank's warp partials, one block sums them in the
// canonical order (identical arithmetic on every rank) and, inside a step, runs the forward half-chain.
__global__ void __launch_bounds__(kMeshThreads)
k_ke_post(const double *__restrict__ ke_warp, int n_ke_warps, double *__restrict__ scal,
          Thermo *__restrict__ th_mid, Thermo *__restrict__ th_cur, int step, const Xchg X, const MeshParams P) {
    __shared__ double red[32];
    if (xchg_has(X, XK_KE)) {
        if (threadIdx.x < 32) xchg_wait_warp(X, XK_KE);
        __syncthreads();
    }
    __shared__ double s_ke;
    const double ke = canonical_sum(ke_warp, n_ke_warps, red);
    if (threadIdx.x == 0) {
        scal[SC_KE] = ke;
        s_ke = ke;
    }
    __syncthreads();
    thermo_finish(th_mid, th_cur, scal, s_ke, step != 0, P);
}

// ------------------------------------------------------------------------------------------------
// Rigid volume constraint (-G V, linear form): SHAKE after the drift, RATTLE after the second half-kick
// (src/functions.h:33-171, called from src/newmd.cpp:33-38,126,152). g_i = VERTEX::neg_GradVi of the most
// recent force evaluation (k_vertex_mesh leaves it in ngv); masses are 1 as in the reference.
//
// Smallest real root of x^3 + a x^2 + b x + c = 0 the way GSL's gsl_poly_solve_cubic reports it in x0
// (trigonometric / Cardano solution, Numerical Recipes 5.6; three real roots are sorted ascending).
__device__ __forceinline__ double cubic_x0(const double a, const double b, const double c) {
    const double q = a * a - 3 * b;
    const double r = 2 * a * a * a - 9 * a * b + 27 * c;
    const double Q = q / 9, R = r / 54;
    const double Q3 = Q * Q * Q, R2 = R * R;
    const double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
    if (R == 0 && Q == 0) return -a / 3;
    if (CR2 == CQ3) {
        const double sqrtQ = sqrt(Q);
        return (R > 0 ? -2 * sqrtQ : -sqrtQ) - a / 3;
    }
    if (R2 < Q3) {
        const double sgnR = R >= 0 ? 1 : -1;
        const double theta = acos(sgnR * sqrt(R2 / Q3));
        const double norm = -2 * sqrt(Q);
        const double pi = 3.14159265358979323846;
        const double r0 = norm * cos(theta / 3) - a / 3;
        const double r1 = norm * cos((theta + 2.0 * pi) / 3) - a / 3;
        const double r2 = norm * cos((theta - 2.0 * pi) / 3) - a / 3;
        return fmin(r0, fmin(r1, r2));
    }
    const double sgnR = R >= 0 ? 1 : -1;
    const double A = -sgnR * pow(fabs(R) + sqrt(R2 - Q3), 1.0 / 3.0);
    return A + Q / A - a / 3;
}

__device__ __forceinline__ double triple(V3 a, V3 b, V3 c) { return dot(a, cross(b, c)); }

// Face sums of SHAKE (which == 0; 8 triple products, then the cubic for the multiplier) or RATTLE
// (which == 1; d/dt of the volume and its derivative along g). One thread per face, ordered reduction,
// the last block leaves the multiplier in scal[SC_LAMBDA] / scal[SC_MU].
template <int WHICH>
__global__ void __launch_bounds__(kMeshThreads)
k_constraint_sums(const double4 *__restrict__ xyzq, const double4 *__restrict__ vel, const double4 *__restrict__ ngv,
                  const int4 *__restrict__ face_v, double *__restrict__ partials /* [8][nblocks] */,
                  unsigned int *__restrict__ ticket, double *__restrict__ scal, const MeshParams P) {
    __shared__ double red[8 * 32];
    __shared__ bool s_last;
    const int nblocks = gridDim.x;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (f < P.nf) {
        const int4 fv = __ldg(face_v + f);
        const V3 p0 = ld3(xyzq, fv.x), p1 = ld3(xyzq, fv.y), p2 = ld3(xyzq, fv.z);
        const double4 a0 = ldcg4(ngv + fv.x), a1 = ldcg4(ngv + fv.y), a2 = ldcg4(ngv + fv.z);
        const V3 g0 = mk(a0.x, a0.y, a0.z), g1 = mk(a1.x, a1.y, a1.z), g2 = mk(a2.x, a2.y, a2.z);
        if (WHICH == 0) {
            acc[0] = triple(g0, g1, g2);
            acc[1] = triple(g0, p1, g2);
            acc[2] = triple(g0, g1, p2);
            acc[3] = triple(p0, g1, g2);
            acc[4] = triple(g0, p1, p2);
            acc[5] = triple(p0, p1, g2);
            acc[6] = triple(p0, g1, p2);
            acc[7] = triple(p0, p1, p2);
        } else {
            const double4 b0 = ldcg4(vel + fv.x), b1 = ldcg4(vel + fv.y), b2 = ldcg4(vel + fv.z);
            const V3 v0 = mk(b0.x, b0.y, b0.z), v1 = mk(b1.x, b1.y, b1.z), v2 = mk(b2.x, b2.y, b2.z);
            acc[0] = triple(v0, p1, p2) + triple(p0, p1, v2) + triple(p0, v1, p2);
            acc[1] = triple(g0, p1, p2) + triple(p0, p1, g2) + triple(p0, g1, p2);
        }
    }
    block_sum<8>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) partials[k * nblocks + blockIdx.x] = acc[k];
        s_last = take_last_ticket(ticket, nblocks);
    }
    __syncthreads();
    if (!s_last) return;
    double tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 8; k++) tot[k] += __ldcg(partials + k * nblocks + b);
    }
    __syncthreads();
    block_sum<8>(tot, red);
    if (threadIdx.x == 0) {
        if (WHICH == 0) {
            const double f0 = tot[0], idt = 1.0 / P.dt;
            const double a = idt * (tot[1] + tot[2] + tot[3]) / f0;
            const double b = idt * idt * (tot[4] + tot[5] + tot[6]) / f0;
            const double c = idt * idt * idt * (tot[7] - 6 * P.ref_volume) / f0;
            scal[SC_LAMBDA] = (c == 0 || f0 == 0) ? 0.0 : cubic_x0(a, b, c);  // already satisfied: bypass
        } else {
            scal[SC_MU] = (tot[0] == 0 || tot[1] == 0) ? 0.0 : -(tot[0] / tot[1]);
        }
    }
}

// SHAKE update: v += lambda g, x += lambda dt g.
__global__ void __launch_bounds__(kMeshThreads)
k_shake_apply(double4 *__restrict__ xyzq, double4 *__restrict__ vel, const double4 *__restrict__ ngv,
              const double *__restrict__ scal, const MeshParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nv) return;
    const double lam = __ldcg(scal + SC_LAMBDA);
    if (lam == 0.0) return;
    const double4 g = ngv[i];
    double4 v = vel[i], x = xyzq[i];
    v.x = v.x + lam * g.x; v.y = v.y + lam * g.y; v.z = v.z + lam * g.z;
    const double ld = lam * P.dt;
    x.x = x.x + ld * g.x; x.y = x.y + ld * g.y; x.z = x.z + ld * g.z;
    vel[i] = v;
    xyzq[i] = x;
}

// RATTLE update v += mu g, then what k_vertex does after the second half-kick: kinetic energy and thermostat
// (INIT: the constraint pass before the first step, src/newmd.cpp:33-38 -- no forward half-chain).
template <bool INIT>
__global__ void __launch_bounds__(kMeshThreads)
k_rattle_apply(double4 *__restrict__ vel, const double4 *__restrict__ ngv, double *__restrict__ partials,
               unsigned int *__restrict__ ticket, double *__restrict__ scal, double *__restrict__ ke_warp, int n_ke_warps,
               Thermo *__restrict__ th_mid, Thermo *__restrict__ th_cur, double *const *__restrict__ peer_ke, const Xchg X,
               const MeshParams P) {
    __shared__ double red[3 * 32];
    __shared__ bool s_last;
    const int i = P.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0, 0, 0};
    if (i < P.row_hi) {
        const double mu = __ldcg(scal + SC_MU);
        const double4 g = ngv[i];
        double4 v = vel[i];
        v.x = v.x + mu * g.x; v.y = v.y + mu * g.y; v.z = v.z + mu * g.z;
        vel[i] = v;
        acc[0] = 0.5 * (v.x * v.x + v.y * v.y + v.z * v.z);
    }
    vertex_tail<false, !INIT>(acc, partials, ticket, scal, ke_warp, n_ke_warps, th_mid, th_cur, peer_ke, X, P, red, &s_last);
}

// ------------------------------------------------------------------------------------------------
// Data feeds of the reference's writers.
//
// k_vertex_geometry: VERTEX::compute_area_normal (src/vertex.cpp:7-18) -- area = 1/3 of the incident face areas,
// normal = normalised sum of area-weighted face normals (normal_F * A_F = direction_F / 2, src/face.cpp:16-26) --
// gathered over the one-ring; consumed by interface_movie (Va and Vq/a columns of p.lammpstrj, src/functions.cpp:50-84).
__global__ void __launch_bounds__(kMeshThreads)
k_vertex_geometry(const double4 *__restrict__ xyzq, const VertexTables T, double4 *__restrict__ out /* nx ny nz area */,
                  const MeshParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nv) return;
    const int nv = P.nv;
    const V3 p = ld3(xyzq, i);
    const int deg = __ldg(T.degree + i);
    const V3 pfirst = ld3(xyzq, __ldg(T.ring + i));
    V3 pn = pfirst, ns = mk(0, 0, 0);
    double area = 0;
#pragma unroll
    for (int k = 0; k < kMaxDeg; k++) {
        if (k >= deg) break;
        const V3 pnn = (k + 1 < deg) ? ld3(xyzq, __ldg(T.ring + (k + 1) * nv + i)) : pfirst;
        const V3 dir = cross(sub(pn, p), sub(pnn, pn));  // face (p, pn, pnn) in its stored orientation
        ns.x += 0.5 * dir.x; ns.y += 0.5 * dir.y; ns.z += 0.5 * dir.z;
        area += 0.5 * sqrt(dot(dir, dir));
        pn = pnn;
    }
    const double inv = 1.0 / sqrt(dot(ns, ns));
    out[i] = make_double4(ns.x * inv, ns.y * inv, ns.z * inv, area * (1. / 3.));
}

// k_local_faces: the per-face energy maps of INTERFACE::compute_local_energies / _by_component
// (src/interface.cpp:1140-1337). The reference adds every unordered pair energy to the faces of both vertices and
// every edge energy to the edge's two faces; gathered by face that is
//   es[F] = sum_{v in F} sum_{j != v} E_es(v, j) = sum_{v in F} 2 * (vertex_e of the energy pass),
//   bending[F] / stretching[F] = sum over the three edges of F,   elastic = stretching + bending.
// face_e holds the three edge indices of a face in ascending order (the order in which the reference's edge loop
// reaches the face).
__global__ void __launch_bounds__(kMeshThreads)
k_local_faces(const int4 *__restrict__ face_v, const int4 *__restrict__ face_e, const double2 *__restrict__ vertex_e,
              const double2 *__restrict__ edge_E, double4 *__restrict__ out /* es elastic bending stretching */,
              const MeshParams P) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.nf) return;
    const int4 v = __ldg(face_v + f), e = __ldg(face_e + f);
    const double es = 2.0 * vertex_e[v.x].x + 2.0 * vertex_e[v.y].x + 2.0 * vertex_e[v.z].x;
    const double2 a = edge_E[e.x], b = edge_E[e.y], c = edge_E[e.z];
    double el = 0;
    el += a.y; el += a.x; el += b.y; el += b.x; el += c.y; el += c.x;  // stretching then bending, edge by edge
    out[f] = make_double4(es, el, a.x + b.x + c.x, a.y + b.y + c.y);
}

// ------------------------------------------------------------------------------------------------
// AoS <-> record conversion on the device (the boundary hands over packed [n][3] doubles, the
// layout of an array of VECTOR3D-as-double; the kernels work on 32-byte records).
__global__ void k_scatter3(double4 *__restrict__ rec, const double *__restrict__ packed, int n, int keep_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double w = keep_w ? rec[i].w : 0.0;
    rec[i] = make_double4(packed[3 * i], packed[3 * i + 1], packed[3 * i + 2], w);
}
__global__ void k_scatter_w(double4 *__restrict__ rec, const double *__restrict__ w, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rec[i].w = w[i];
}
__global__ void k_gather3(const double4 *__restrict__ rec, double *__restrict__ packed, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 v = rec[i];
    packed[3 * i] = v.x;
    packed[3 * i + 1] = v.y;
    packed[3 * i + 2] = v.z;
}

// packed block [pos | vel | force (| KE)] <-> records, one launch for the three arrays (npsl_md_step_packed)
__global__ void k_unpack_state(double4 *__restrict__ xyzq, double4 *__restrict__ vel, double4 *__restrict__ force,
                               const double *__restrict__ packed, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *px = packed + 3 * (size_t) i, *pv = px + 3 * (size_t) n, *pf = pv + 3 * (size_t) n;
    xyzq[i] = make_double4(px[0], px[1], px[2], xyzq[i].w);
    vel[i] = make_double4(pv[0], pv[1], pv[2], 0.0);
    force[i] = make_double4(pf[0], pf[1], pf[2], 0.0);
}
__global__ void k_pack_state(const double4 *__restrict__ xyzq, const double4 *__restrict__ vel,
                             const double4 *__restrict__ force, const double *__restrict__ scal, double *__restrict__ packed,
                             int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) packed[9 * (size_t) n] = scal[SC_KE];
    if (i >= n) return;
    double *px = packed + 3 * (size_t) i, *pv = px + 3 * (size_t) n, *pf = pv + 3 * (size_t) n;
    const double4 x = xyzq[i], v = vel[i], f = force[i];
    px[0] = x.x; px[1] = x.y; px[2] = x.z;
    pv[0] = v.x; pv[1] = v.y; pv[2] = v.z;
    pf[0] = f.x; pf[1] = f.y; pf[2] = f.z;
}

}  // namespace npsl
