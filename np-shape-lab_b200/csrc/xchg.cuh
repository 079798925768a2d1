// One-shot exchanges between the GPUs of a sharded job over NVLink peer memory (sm_100a).
//
// The reference exchanges per-rank slices with boost::mpi::all_gather (src/newforces.cpp:268-271).
// Here each rank's exchange buffers are mapped into every other rank's address space (CUDA IPC), the
// producing kernel stores its results straight into the consumers' buffers, and arrival is signalled with
// one epoch counter per (exchange kind, source rank) in the consumer's memory:
//   producer: after a CTA barrier one thread per CTA fences the CTA's remote stores (fence.sys, cumulative)
//             and takes a ticket; the last CTA of the kernel bumps the kind's epoch and release-stores it
//             into flags[kind][my rank] of every rank;
//   consumer: one warp spins (acquire loads) until flags[kind][r] has reached the next epoch for every r.
// Epochs only grow and both sides count exchanges from zero, so nothing is reset between steps and the
// kernels can be replayed from a CUDA graph unchanged.
//   Row-sharded plan, three kinds per MD step: positions (k_kick_drift -> everyone), pair-force sums of the
//   virtual ranks (k_pair_reduce -> owner of the vertex), per-warp kinetic energies (k_vertex -> everyone).
//   A rank cannot run ahead by more than one exchange, which makes the buffers safe to overwrite.
//   Replicated-integrator plan, one kind per step: every rank broadcasts the sums of its virtual ranks
//   (k_pair_reduce -> everyone) and integrates all vertices itself. The receive buffer is double-buffered
//   on the parity of the epoch: a rank can be at most one exchange ahead of the slowest reader.
// Every rank produces each exchange exactly once per force evaluation, so a consumer's next epoch is its
// own producer epoch (Multi-GPU).
#pragma once
#include <cuda_runtime.h>

namespace npsl {

constexpr int kMaxRanks = 8;
// XK_LJ0 / XK_LJ1: the two parities of the collision-flag exchange (k_lj_flag); never part of Xchg::on.
// XK_STATE: the packed host state of npsl_md_step_packed, every rank uploads its own row slice and shares it.
enum { XK_POS = 0, XK_FV = 1, XK_KE = 2, XK_LJ0 = 3, XK_LJ1 = 4, XK_STATE = 5, XK_KINDS = 6 };
constexpr unsigned long long kXchgTimeoutNs = 30ull * 1000000000ull;  // a wait that long means a peer is gone

struct Xchg {
    int world, rank;
    int on;                                   // bit k set: exchange kind k goes over peer memory (0: single GPU or NCCL)
    int fv_bcast;                             // 1: every rank receives every virtual-rank sum (replicated integrator)
    int cta_fence_sys;                        // 1: every CTA fences at system scope before its ticket (see xchg_signal_cta)
    unsigned long long *const *peer_flags;    // [world] -> each rank's flags[XK_KINDS][kMaxRanks]
    unsigned long long *state;                // local: [kind] producer epoch, [XK_KINDS + kind] consumer epoch
    unsigned int *tickets;                    // local: [kind]
    int *timed_out;                           // local: set when a wait gave up (reported as NPSL_ERR_STATE by the host)
};

__device__ __forceinline__ bool xchg_has(const Xchg &X, const int kind) { return (X.on >> kind) & 1; }

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One thread per CTA of the producing kernel, after a CTA barrier that follows the CTA's remote stores.
// The CTA's stores are ordered before its ticket by a device-scope fence (fences are cumulative over the barrier);
// the CTA that draws the last ticket has thereby observed every CTA's stores, fences once at system scope and
// then publishes the new epoch to every rank with relaxed stores (fence + store = release pattern; causality order
// is transitive across the two scopes). X.cta_fence_sys = 1 restores a system-scope fence in every CTA.
__device__ __forceinline__ void xchg_signal_cta(const Xchg &X, const int kind, const unsigned int nblocks) {
    if (X.cta_fence_sys) __threadfence_system(); else __threadfence();
    const unsigned int t = atomicAdd(X.tickets + kind, 1u);
    if (t == nblocks - 1) {
        X.tickets[kind] = 0;
        const unsigned long long e = X.state[kind] + 1;
        X.state[kind] = e;
        __threadfence_system();
        for (int r = 0; r < X.world; r++) st_relaxed_sys(X.peer_flags[r] + kind * kMaxRanks + X.rank, e);
    }
}

// Called by every thread of every CTA (1-D block) of the producing kernel, after its remote stores.
__device__ __forceinline__ void xchg_signal(const Xchg &X, const int kind, const unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) xchg_signal_cta(X, kind, nblocks);
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spins (acquire loads) until *f >> shift has reached e. Every rank has to issue the same sequence of force
// evaluations (include/npsl.h, "collective calls"); if a peer never arrives the wait gives up after kXchgTimeoutNs
// and leaves a mark that the host turns into NPSL_ERR_STATE, instead of hanging the GPU.
__device__ __forceinline__ unsigned long long xchg_spin(const Xchg &X, const unsigned long long *f,
                                                        const unsigned long long e, const int shift, const unsigned ns) {
    unsigned long long v = ld_acquire_sys(f);
    if ((v >> shift) >= e) return v;
    const unsigned long long t0 = global_ns();
    unsigned int it = 0;
    while (((v = ld_acquire_sys(f)) >> shift) < e) {
        __nanosleep(ns);
        if ((++it & 1023u) == 0 && global_ns() - t0 > kXchgTimeoutNs) {
            *X.timed_out = 1;
            break;
        }
    }
    return v;
}

// One warp: waits until every rank's signal of the next epoch of `kind` has arrived. Returns to all lanes.
__device__ __forceinline__ void xchg_wait_warp(const Xchg &X, const int kind) {
    const int lane = threadIdx.x & 31;
    const unsigned long long e = X.state[XK_KINDS + kind] + 1;
    if (lane < X.world) xchg_spin(X, X.peer_flags[X.rank] + kind * kMaxRanks + lane, e, 0, 64);
    __syncwarp();
    if (lane == 0) X.state[XK_KINDS + kind] = e;
    __threadfence();
}

// One warp of any CTA of the consuming kernel: waits for every rank's signal of the exchange this rank itself
// produced last (its own producer epoch; valid because every rank produces each exchange exactly once per force
// evaluation and this rank's producer precedes the consumer in stream order). No state is written, so any
// number of CTAs may call it. Returns the epoch to all lanes.
__device__ __forceinline__ unsigned long long xchg_wait_own_epoch(const Xchg &X, const int kind) {
    const int lane = threadIdx.x & 31;
    const unsigned long long e = X.state[kind];
    if (lane < X.world) xchg_spin(X, X.peer_flags[X.rank] + kind * kMaxRanks + lane, e, 0, 32);
    __syncwarp();
    __threadfence();
    return e;
}

__global__ void k_xchg_wait(const Xchg X, const int kind) {
    if (threadIdx.x < 32) xchg_wait_warp(X, kind);
}

// The WCA / LJ collision detector of the symmetric pair kernel is per rank (a rank only sees the pairs of its own
// units), but the reference applies the WCA term on every rank's rows unconditionally (src/newforces.cpp:165-171), so
// the exact pass k_lj and the gate in k_vertex must act on the OR over all ranks. One warp, right after k_pair_sym on
// the stream of k_lj: lane r stores (epoch << 1 | my flag) into rank r's slot for this rank, then waits for rank r's
// slot here; the OR replaces the local flag. The slots alternate with the parity of the epoch: a rank can be at most
// one force evaluation ahead of the slowest one (it needs everybody's pair-force sums to finish a step).
__global__ void k_lj_flag(const Xchg X, int *__restrict__ lj_hit) {
    const int lane = threadIdx.x & 31;
    const unsigned long long e = X.state[XK_LJ0] + 1;
    const int kind = XK_LJ0 + (int) (e & 1ull);
    const unsigned long long mine = (e << 1) | (unsigned long long) (*(volatile int *) lj_hit != 0);
    int hit = 0;
    if (lane < X.world) {
        st_release_sys(X.peer_flags[lane] + kind * kMaxRanks + X.rank, mine);
        hit = (int) (xchg_spin(X, X.peer_flags[X.rank] + kind * kMaxRanks + lane, e, 1, 64) & 1ull);
    }
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) {
        *lj_hit = hit ? 1 : 0;
        X.state[XK_LJ0] = e;
    }
}

// Packed host state (npsl_md_step_packed on a replicated-plan context): every rank has copied only ITS row slice
// [lo, hi) of posvec / velvec / forvec from its host block into its packed device image; this kernel stores the slice
// into every peer's image over NVLink, so that each PCIe link carries 1 / world of the state instead of all of it.
// image = [pos | vel | force], n vertices each; one thread per double of the slice.
__global__ void __launch_bounds__(256)
k_state_share(const double *__restrict__ image, double *const *__restrict__ peer_image, int n, int lo, int hi, const Xchg X) {
    const int len = 3 * (hi - lo);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * len) {
        const size_t off = (size_t) (k / len) * 3 * n + (size_t) 3 * lo + (k % len);
        const double v = image[off];
        for (int r = 0; r < X.world; r++)
            if (r != X.rank) peer_image[r][off] = v;
    }
    xchg_signal(X, XK_STATE, gridDim.x);
}

// NCCL exchanges: the flags of all ranks were all-gathered into all[world]
__global__ void k_lj_or(const int *__restrict__ all, int world, int *__restrict__ lj_hit) {
    int hit = 0;
    for (int r = threadIdx.x; r < world; r += 32) hit |= all[r];
    hit = __any_sync(0xffffffffu, hit != 0);
    if (threadIdx.x == 0) *lj_hit = hit ? 1 : 0;
}

}  // namespace npsl
