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
// kernels can be replayed from a CUDA graph unchanged. Three kinds per MD step: positions (k_kick_drift
// -> everyone), pair-force sums of the virtual ranks (k_pair_reduce -> owner of the vertex), per-warp
// kinetic energies (k_vertex -> everyone). A rank cannot run ahead by more than one exchange, which is
// what makes the buffers safe to overwrite (Multi-GPU).
#pragma once
#include <cuda_runtime.h>

namespace npsl {

constexpr int kMaxRanks = 8;
enum { XK_POS = 0, XK_FV = 1, XK_KE = 2, XK_KINDS = 3 };

struct Xchg {
    int world, rank;
    int on;                                   // 0: single GPU or NCCL exchanges
    unsigned long long *const *peer_flags;    // [world] -> each rank's flags[XK_KINDS][kMaxRanks]
    unsigned long long *state;                // local: [kind] producer epoch, [XK_KINDS + kind] consumer epoch
    unsigned int *tickets;                    // local: [kind]
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by every thread of every CTA of the producing kernel, after its remote stores.
__device__ __forceinline__ void xchg_signal(const Xchg &X, const int kind, const unsigned int nblocks) {
    __syncthreads();  // the CTA's remote stores happen-before thread 0's system-scope fence (fences are cumulative)
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(X.tickets + kind, 1u);
        if (t == nblocks - 1) {
            X.tickets[kind] = 0;
            __threadfence();
            const unsigned long long e = X.state[kind] + 1;
            X.state[kind] = e;
            __threadfence_system();
            for (int r = 0; r < X.world; r++) st_release_sys(X.peer_flags[r] + kind * kMaxRanks + X.rank, e);
        }
    }
}

// One warp: waits until every rank's signal of the next epoch of `kind` has arrived. Returns to all lanes.
__device__ __forceinline__ void xchg_wait_warp(const Xchg &X, const int kind) {
    const int lane = threadIdx.x & 31;
    const unsigned long long e = X.state[XK_KINDS + kind] + 1;
    if (lane < X.world) {
        const unsigned long long *f = X.peer_flags[X.rank] + kind * kMaxRanks + lane;
        while (ld_acquire_sys(f) < e) __nanosleep(64);
    }
    __syncwarp();
    if (lane == 0) X.state[XK_KINDS + kind] = e;
    __threadfence();
}

__global__ void k_xchg_wait(const Xchg X, const int kind) {
    if (threadIdx.x < 32) xchg_wait_warp(X, kind);
}

}  // namespace npsl
