// Explicit counterions, FP64, sm_100a.
//
// The reference walks, besides the mesh-mesh pairs, every mesh-ion pair twice (once for the force on the vertex,
// once for the force on the ion; here once) and every ordered ion-ion pair (src/newforces.cpp:44-65,179-250), and repeats the
// walks for the energies (src/interface.cpp:1059-1089). Ions are few (|q| / valency, hundreds to a few thousand), so
// these are two small kernels beside the mesh-mesh pair kernel: the screened-Coulomb part uses the pair kernel's
// rsqrt / table-exp (<= 3e-16 relative error, pair_kernel.cuh), the WCA branch -- live here, ions do touch the membrane
// and each other -- the reference's expression with per-class LJ lengths (INTERFACE::lj_length_mesh_ions for
// mesh-ion, the ion diameter for ion-ion). Fixed summation orders, no atomics.
//   k_ion_mesh       every mesh-ion pair once: force of all ions on each vertex, the vertex's mesh-ion energies (counted
//                    once, on the mesh side, like the reference), and per CTA the vertices' pull on every ion
//   k_ion_rows       one CTA per ion: all other ions on it plus the per-CTA sums of k_ion_mesh (ordered block sum),
//                    its ion-ion energies (halved)
//   k_ion_kick_drift first half-kick and drift of the ions (PARTICLE::update_real_velocity / update_real_position,
//                    src/particle.h:66-83, src/newmd.cpp:116-123)
//   k_ion_finish     second half-kick, the ions' kinetic energy (src/newenergies.h:18-26) and the totals of the ion-ion
//                    energies, one CTA, ordered; runs before the kernel that advances the Nose-Hoover chains.
#pragma once
#include "mesh_kernels.cuh"
#include "pair_kernel.cuh"

namespace npsl {

struct IonParams {
    int n_ions, nv;
    double coef;        // scalefactor / em
    double inv_lambda;  // 1 / inv_kappa_out
    double elj;
    double d2_mi, cut2_mi;  // mesh-ion LJ length squared and dcut2 * that
    double d2_ii, cut2_ii;  // ion-ion (ion diameter)
    double dt;
    double c2s;             // -log2(e) / lambda * 2^TB, the argument scale of exp2_neg
};

// One pair, d = p1 - p2, qq = coef q1 q2. s: force on particle 1 is s * d. es, lj: pair energies.
__device__ __forceinline__ void ion_pair(const double dx, const double dy, const double dz, const double qq,
                                         const double inv_lambda, const double c2s, const double *__restrict__ tab,
                                         const double elj, const double d2, const double cut2, double &s, double &es,
                                         double &lj) {
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const double rinv = rsqrt_fast(r2);
    const double qe = qq * exp2_neg(r2 * rinv, c2s, tab);  // coef q1 q2 e^{-r/lambda}
    es = qe * rinv;
    s = qe * (rinv * rinv) * (rinv + inv_lambda);
    lj = 0.0;
    if (r2 < cut2) {
        const double r6 = r2 * r2 * r2, r12 = r6 * r6, d6 = d2 * d2 * d2, d12 = d6 * d6;
        s += 48 * elj * ((d12 / r12) - 0.5 * (d6 / r6)) * (1 / r2);
        lj = 4 * elj * (d6 / r6) * ((d6 / r6) - 1) + elj;
    }
}

// Every mesh-ion pair once (Newton's third law), the way k_pair_sym treats the mesh-mesh pairs: a CTA owns 128
// vertices (one per thread, force and energies in registers) and walks the ions in tiles of 128 staged in shared memory
// together with one force accumulator per ion. The four warps work on the four 32-ion groups of a tile and rotate
// groups between four phases; inside a phase lane l visits ion (l + s) mod 32 at step s, so an ion's accumulator is
// touched by exactly one thread at a time: no atomics, fixed order. Grid = (vertex blocks, ion tiles): a CTA works on
// one tile, so that a few hundred ions still give every SM enough warps. The vertices' sums go to out[tile][vertex]
// (k_vertex adds the tiles in order), the ions' sums to ion_part[vertex block][3][n_ions] (k_ion_rows adds the blocks
// in order).
constexpr int kIonTile = 128;

__global__ void __launch_bounds__(128)
k_ion_mesh(const double4 *__restrict__ xyzq, const double4 *__restrict__ ion_xyzq, double4 *__restrict__ out /* f, es */,
           double *__restrict__ out_lj, double *__restrict__ ion_part, const IonParams P) {
    __shared__ double sx[kIonTile], sy[kIonTile], sz[kIonTile], sq[kIonTile];
    __shared__ double sax[kIonTile], say[kIonTile], saz[kIonTile];
    __shared__ double tab[kExpTab];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int k = t; k < kExpTab; k += blockDim.x) tab[k] = c_exp2_tab[k];
    const int i = blockIdx.x * blockDim.x + t;
    const double4 pi = xyzq[min(i, P.nv - 1)];
    const double qi = P.coef * pi.w;
    double fx = 0, fy = 0, fz = 0, es = 0, lj = 0;
    {
        const int j0 = blockIdx.y * kIonTile;
        const int j = j0 + t;
        if (j < P.n_ions) {
            const double4 v = ion_xyzq[j];
            sx[t] = v.x; sy[t] = v.y; sz[t] = v.z; sq[t] = v.w;
        } else {  // padding ions: far away, no charge
            sx[t] = 1.0e3 + t; sy[t] = 0.0; sz[t] = 0.0; sq[t] = 0.0;
        }
        sax[t] = 0.0; say[t] = 0.0; saz[t] = 0.0;
        __syncthreads();
        for (int p = 0; p < 4; p++) {
            const int o = ((warp + p) & 3) * 32;
            if (j0 + o < P.n_ions) {
#pragma unroll 2
                for (int s = 0; s < 32; s++) {
                    const int col = o + ((lane + s) & 31);
                    const double dx = pi.x - sx[col], dy = pi.y - sy[col], dz = pi.z - sz[col];
                    double f, e1, e2;
                    ion_pair(dx, dy, dz, qi * sq[col], P.inv_lambda, P.c2s, tab, P.elj, P.d2_mi, P.cut2_mi, f, e1, e2);
                    if (i >= P.nv || j0 + col >= P.n_ions) { f = 0.0; e1 = 0.0; e2 = 0.0; }  // rows / ions past the end
                    const double gx = f * dx, gy = f * dy, gz = f * dz;
                    fx += gx; fy += gy; fz += gz;
                    es += e1; lj += e2;
                    sax[col] -= gx; say[col] -= gy; saz[col] -= gz;
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (j < P.n_ions) {
            double *q = ion_part + (size_t) blockIdx.x * 3 * P.n_ions + j;
            q[0] = sax[t];
            q[P.n_ions] = say[t];
            q[2 * (size_t) P.n_ions] = saz[t];
        }
    }
    if (i < P.nv) {
        out[(size_t) blockIdx.y * P.nv + i] = make_double4(fx, fy, fz, es);
        out_lj[(size_t) blockIdx.y * P.nv + i] = lj;
    }
}

__global__ void __launch_bounds__(256)
k_ion_rows(const double4 *__restrict__ ion_xyzq, const double *__restrict__ ion_part, int n_mesh_ctas,
           double4 *__restrict__ ion_force, double2 *__restrict__ ion_e, const IonParams P) {
    __shared__ double red[5 * 32];
    __shared__ double tab[kExpTab];
    for (int k = threadIdx.x; k < kExpTab; k += blockDim.x) tab[k] = c_exp2_tab[k];
    __syncthreads();
    const int i = blockIdx.x;
    const double4 pi = ion_xyzq[i];
    double acc[5] = {0, 0, 0, 0, 0};  // fx fy fz es_ii lj_ii
#pragma unroll 2
    for (int j = threadIdx.x; j < P.n_ions; j += blockDim.x) {  // ion-ion, src/newforces.cpp:203-226
        if (j == i) continue;
        const double4 pj = ion_xyzq[j];
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        double s, e1, e2;
        ion_pair(dx, dy, dz, P.coef * pi.w * pj.w, P.inv_lambda, P.c2s, tab, P.elj, P.d2_ii, P.cut2_ii, s, e1, e2);
        acc[0] += s * dx; acc[1] += s * dy; acc[2] += s * dz;
        acc[3] += e1; acc[4] += e2;
    }
    for (int b = threadIdx.x; b < n_mesh_ctas; b += blockDim.x) {  // the vertices' pull: per-CTA sums of k_ion_mesh
        const double *q = ion_part + (size_t) b * 3 * P.n_ions + i;
        acc[0] += __ldcg(q); acc[1] += __ldcg(q + P.n_ions); acc[2] += __ldcg(q + 2 * (size_t) P.n_ions);
    }
    block_sum<5>(acc, red);
    if (threadIdx.x == 0) {
        ion_force[i] = make_double4(acc[0], acc[1], acc[2], 0.0);
        ion_e[i] = make_double2(0.5 * acc[3], 0.5 * acc[4]);
    }
}

__global__ void __launch_bounds__(128)
k_ion_kick_drift(double4 *__restrict__ ion_xyzq, double4 *__restrict__ ion_vel, const double4 *__restrict__ ion_force,
                 const double *__restrict__ scal, const IonParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_ions) return;
    const double ef = __ldcg(scal + SC_EXPFAC_IONS), kick = __ldcg(scal + SC_KICK_IONS);
    double4 v = ion_vel[i], x = ion_xyzq[i];
    const double4 f = ion_force[i];
    v.x = fma(v.x, ef, f.x * kick);
    v.y = fma(v.y, ef, f.y * kick);
    v.z = fma(v.z, ef, f.z * kick);
    x.x = fma(v.x, P.dt, x.x);
    x.y = fma(v.y, P.dt, x.y);
    x.z = fma(v.z, P.dt, x.z);
    ion_vel[i] = v;
    ion_xyzq[i] = x;
}

// One CTA. STEP: second half-kick first. Leaves the ions' kinetic energy and the ion-ion energy totals in the scalar
// block (ordered: per-thread strided sums, then the fixed block tree).
template <bool STEP>
__global__ void __launch_bounds__(256)
k_ion_finish(double4 *__restrict__ ion_vel, const double4 *__restrict__ ion_force, const double2 *__restrict__ ion_e,
             double *__restrict__ scal, const IonParams P) {
    __shared__ double red[3 * 32];
    const double ef = __ldcg(scal + SC_EXPFAC_IONS), kick = __ldcg(scal + SC_KICK_IONS);
    double acc[3] = {0, 0, 0};  // ke, es, lj
    for (int i = threadIdx.x; i < P.n_ions; i += blockDim.x) {
        double4 v = ion_vel[i];
        if (STEP) {
            const double4 f = ion_force[i];
            v.x = fma(v.x, ef, f.x * kick);
            v.y = fma(v.y, ef, f.y * kick);
            v.z = fma(v.z, ef, f.z * kick);
            ion_vel[i] = v;
        }
        acc[0] += 0.5 * (v.x * v.x + v.y * v.y + v.z * v.z);
        const double2 e = ion_e[i];
        acc[1] += e.x; acc[2] += e.y;
    }
    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        scal[SC_KE_IONS] = acc[0];
        scal[SC_ES_IONS] = acc[1];
        scal[SC_LJ_IONS] = acc[2];
    }
}

}  // namespace npsl
