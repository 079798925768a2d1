// Development microbenchmarks (GPU box): what limits the FP64 pipe on B200 for operand patterns
// like those of the pair kernel. Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// A: fma(v, b, a) with a, b kernel parameters (constant / uniform operands)
template <int CH>
__global__ void __launch_bounds__(256) kA(double *out, int iters, double a, double b) {
    double v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = a + k + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = fma(v[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 12345.678) out[0] = s;
}
// B: three distinct register operands per DFMA
template <int CH>
__global__ void __launch_bounds__(256) kB(double *out, int iters, double a, double b) {
    double v[CH], w[CH], u[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { v[k] = a + k + threadIdx.x; w[k] = b + 1e-9 * (k + threadIdx.x); u[k] = a * 1e-3 * (k + 1 + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) v[k] = fma(v[k], w[k], u[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k] + w[k] + u[k];
    if (s == 12345.678) out[0] = s;
}
// C: mix of DMUL / DADD / DFMA with register operands (ratio 1:1:2... roughly the pair kernel's 32:16:84)
template <int CH>
__global__ void __launch_bounds__(256) kC(double *out, int iters, double a, double b) {
    double v[CH], w[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { v[k] = a + k + threadIdx.x; w[k] = b + 1e-9 * (k + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            double t = v[k] * w[k];          // DMUL
            double s = t + w[k];             // DADD
            v[k] = fma(s, w[k], v[k]);       // DFMA
            v[k] = fma(v[k], b, t);          // DFMA
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k] + w[k];
    if (s == 12345.678) out[0] = s;
}
// D: DFMA + interleaved integer ALU work (1 IMAD per 2 DFMA)
template <int CH>
__global__ void __launch_bounds__(256) kD(double *out, int iters, double a, double b) {
    double v[CH];
    int q[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { v[k] = a + k + threadIdx.x; q[k] = k + threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            v[k] = fma(v[k], b, a);
            v[k] = fma(v[k], b, a);
            q[k] = q[k] * 3 + it;
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k] + q[k];
    if (s == 12345.678) out[0] = s;
}
// E: DFMA + MUFU.RSQ64H (1 per 32 DFMA) like the pair kernel
template <int CH>
__global__ void __launch_bounds__(256) kE(double *out, int iters, double a, double b) {
    double v[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) v[k] = a + k + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
#pragma unroll
            for (int m = 0; m < 8; m++) v[k] = fma(v[k], b, a);
        }
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v[0]));
        v[1] += y;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += v[k];
    if (s == 12345.678) out[0] = s;
}

template <class F>
double run(const char *name, F launch, double dfma_per_thread_iter, int blocks, int threads, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(64);
    CK(cudaDeviceSynchronize());
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        launch(iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double) blocks * threads * dfma_per_thread_iter * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    printf("%-44s blocks/SM=%d threads=%d  %.4e fp64-instr/s  (%.1f%% of 148*64*1.965e9)\n", name, blocks / 148, threads, best,
           100.0 * best / (148.0 * 64 * 1.965e9));
    return best;
}

int main() {
    double *out; CK(cudaMalloc(&out, 8));
    const int iters = 1 << 14;
    for (int bps : {1, 2, 4, 8}) {
        int blocks = 148 * bps;
        run("A<16> fma(v,const,const)", [&](int it) { kA<16><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 16, blocks, 256, iters);
        run("A<4>  fma(v,const,const)", [&](int it) { kA<4><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 4, blocks, 256, iters);
        run("A<2>  fma(v,const,const)", [&](int it) { kA<2><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 2, blocks, 256, iters);
        run("A<1>  fma(v,const,const)", [&](int it) { kA<1><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 1, blocks, 256, iters);
        run("B<8>  fma(v,w,u) 3 reg operands", [&](int it) { kB<8><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 8, blocks, 256, iters);
        run("B<4>  fma(v,w,u) 3 reg operands", [&](int it) { kB<4><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 4, blocks, 256, iters);
        run("C<8>  mul/add/fma/fma reg operands", [&](int it) { kC<8><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 32, blocks, 256, iters);
        run("C<4>  mul/add/fma/fma reg operands", [&](int it) { kC<4><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 16, blocks, 256, iters);
        run("D<8>  2 fma + 1 imad", [&](int it) { kD<8><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 16, blocks, 256, iters);
        run("E<4>  32 fma + 1 mufu.rsq64h", [&](int it) { kE<4><<<blocks, 256>>>(out, it, 1.0000001, 0.9999999); }, 32, blocks, 256, iters);
    }
    return 0;
}
