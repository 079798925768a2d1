// Explicit counterions, FP64, sm_100a.
//
// The reference walks, besides the mesh-mesh pairs, every mesh-ion pair twice (once for the force on the vertex,
// once for the force on the ion) and every ordered ion-ion pair (src/newforces.cpp:44-65,179-250), and repeats the
// walks for the energies (src/interface.cpp:1059-1089). Ions are few (|q| / valency, hundreds to a few thousand), so
// these are two small kernels beside the mesh-mesh pair kernel: the screened-Coulomb part uses the pair kernel's
// rsqrt / table-exp (<= 3e-16 relative error, pair_kernel.cuh), the WCA branch -- live here, ions do touch the membrane
// and each other -- the reference's expression with per-class LJ lengths (INTERFACE::lj_length_mesh_ions for
// mesh-ion, the ion diameter for ion-ion). Fixed summation orders, no atomics.
//   k_ion_mesh_rows  four lanes per mesh vertex, ions staged through shared memory: force of all ions on the vertex
//                    and the vertex's mesh-ion energies (counted once, on the mesh side, like the reference)
//   k_ion_rows       one CTA per ion: force of all vertices and all other ions on it (ordered block sum), its
//                    ion-ion energies (halved)
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

// Four lanes per vertex: lane s of a vertex's quad takes the ions s, s + 4, ... of every staged tile; the four partial
// sums are combined in the fixed tree (s0 + s1) + (s2 + s3). 32 vertices per CTA.
__global__ void __launch_bounds__(128)
k_ion_mesh_rows(const double4 *__restrict__ xyzq, const double4 *__restrict__ ion_xyzq, double4 *__restrict__ out /* f, es */,
                double *__restrict__ out_lj, const IonParams P) {
    __shared__ double4 sj[256];
    __shared__ double tab[kExpTab];
    for (int k = threadIdx.x; k < kExpTab; k += blockDim.x) tab[k] = c_exp2_tab[k];
    const int i = blockIdx.x * 32 + (threadIdx.x >> 2), sub = threadIdx.x & 3;
    const double4 pi = xyzq[min(i, P.nv - 1)];
    double acc[5] = {0, 0, 0, 0, 0};  // fx fy fz es lj
    for (int j0 = 0; j0 < P.n_ions; j0 += 256) {
        const int cnt = min(256, P.n_ions - j0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) sj[k] = ion_xyzq[j0 + k];
        __syncthreads();
#pragma unroll 4  // independent rsqrt / exp chains in flight; each lane's sums stay in ion order
        for (int k = sub; k < cnt; k += 4) {
            const double4 pj = sj[k];
            const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            double s, e1, e2;
            ion_pair(dx, dy, dz, P.coef * pi.w * pj.w, P.inv_lambda, P.c2s, tab, P.elj, P.d2_mi, P.cut2_mi, s, e1, e2);
            acc[0] += s * dx; acc[1] += s * dy; acc[2] += s * dz;
            acc[3] += e1; acc[4] += e2;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
    }
    if (sub == 0 && i < P.nv) {
        out[i] = make_double4(acc[0], acc[1], acc[2], acc[3]);
        out_lj[i] = acc[4];
    }
}

__global__ void __launch_bounds__(256)
k_ion_rows(const double4 *__restrict__ xyzq, const double4 *__restrict__ ion_xyzq, double4 *__restrict__ ion_force,
           double2 *__restrict__ ion_e, const IonParams P) {
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
#pragma unroll 4
    for (int j = threadIdx.x; j < P.nv; j += blockDim.x) {  // mesh-ion (for ions), src/newforces.cpp:228-249
        const double4 pj = ldg4(xyzq + j);
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        double s, e1, e2;
        ion_pair(dx, dy, dz, P.coef * pi.w * pj.w, P.inv_lambda, P.c2s, tab, P.elj, P.d2_mi, P.cut2_mi, s, e1, e2);
        acc[0] += s * dx; acc[1] += s * dy; acc[2] += s * dz;
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
