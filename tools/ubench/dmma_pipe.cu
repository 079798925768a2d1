// Development microbenchmark (GPU box): does the FP64 tensor instruction (mma.sync m8n8k4 f64, "DMMA") run beside
// the FP64 FMA pipe on B200 or on the same units? Times a DFMA-only loop, a DMMA-only loop and loops that interleave the
// two. If the mixed loop takes max(A, B) the pair kernel's accumulations could move to DMMA for free; if it takes A + B
// they share the units and nothing is gained. Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_pipe dmma_pipe.cu && ./dmma_pipe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double &d0, double &d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NF independent DFMA chains and NM independent DMMA accumulators per thread per iteration
template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
    double v[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int k = 0; k < NF; k++) v[k] = a + k + threadIdx.x;
#pragma unroll
    for (int k = 0; k < NM; k++) { c0[k] = k; c1[k] = -k; }
    const double fa = a * 1e-3 + threadIdx.x * 1e-9, fb = b + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < (NF > NM ? NF : NM); k++) {
            if (k < NF) v[k] = fma(v[k], b, a);
            if (k < NM) dmma(c0[k], c1[k], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NF; k++) s += v[k];
#pragma unroll
    for (int k = 0; k < NM; k++) s += c0[k] + c1[k];
    if (s == 12345.678) out[0] = s;
}

template <int NF, int NM>
static void run(const char *name, double *out, int sms) {
    const int iters = 20000, blocks = sms * 8;
    k_mix<NF, NM><<<blocks, 256>>>(out, 100, 0.5, 0.999);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_mix<NF, NM><<<blocks, 256>>>(out, iters, 0.5, 0.999);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double warps = (double) blocks * 8, n = warps * iters;
    const double dfma = n * NF * 32, dmma_fma = n * NM * 256;
    printf("%-28s %8.3f ms  DFMA %7.3f T/s  DMMA-FMA %7.3f T/s  total FMA %7.3f T/s\n", name, ms, dfma / ms * 1e-9,
           dmma_fma / ms * 1e-9, (dfma + dmma_fma) / ms * 1e-9);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double *out;
    CK(cudaMalloc(&out, 64));
    const int sms = p.multiProcessorCount;
    run<8, 0>("DFMA x8", out, sms);
    run<16, 0>("DFMA x16", out, sms);
    run<0, 4>("DMMA x4", out, sms);
    run<0, 8>("DMMA x8", out, sms);
    run<8, 1>("DFMA x8 + DMMA x1", out, sms);
    run<8, 2>("DFMA x8 + DMMA x2", out, sms);
    run<16, 2>("DFMA x16 + DMMA x2", out, sms);
    run<8, 4>("DFMA x8 + DMMA x4", out, sms);
    run<8, 8>("DFMA x8 + DMMA x8", out, sms);
    return 0;
}
