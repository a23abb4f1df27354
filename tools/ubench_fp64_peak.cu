// Cross-check of the FP64 roofline denominator: DFMA throughput of the whole GPU by wall clock (CUDA events), for
// 4 / 8 / 16 / 32 resident warps per SM with 8 independent chains each.  Peak = 64 lanes per SM per cycle
// (148 x 64 x 2 x 1.965 GHz = 37.2 TFLOP/s); bench.py's DFMA probe and ncu's pipe-utilisation metric use the same.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub3 tools/ubench_fp64_peak.cu && /tmp/ub3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, (size_t)148 * 32 * 1024 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int block = 128, grid = 148 * warps / 4;
        k<<<grid, block>>>(out, iters);
        cudaEventRecord(e0); k<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)grid * block * iters * 16 * 8 * 2;
        printf("%2d warps/SM: %.3f ms, %.2f TFLOP/s\n", warps, ms, flops / ms / 1e9);
    }
    return 0;
}
