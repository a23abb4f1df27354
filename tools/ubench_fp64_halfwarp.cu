// Microbenchmark: does a warp with only 16 active lanes issue DFMA faster than a full warp on B200?
// (FP64 pipe = 16 lanes per SM sub-partition: a full-warp DFMA occupies it for 2 cycles.)  Decides whether 16
// trajectories per warp would help the strong-scaling rollouts.  One warp per SM sub-partition is not controllable, so
// we launch 1 block of 32 threads per SM (1 warp/SM) and time with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub tools/ubench_fp64_halfwarp.cu && /tmp/ub
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int active, int iters, int chains) {
    const int lane = threadIdx.x & 31;
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + lane * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = 0, t1 = 0;
    if (lane < active) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (chains == 1) { a[0] = fma(a[0], m, c); }
                else if (chains == 2) { a[0] = fma(a[0], m, c); a[1] = fma(a[1], m, c); }
                else if (chains == 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = fma(a[i], m, c);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
                }
            }
        }
        t1 = clock64();
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    for (int warps : {1, 2, 4, 8}) for (int chains : {1, 2, 4, 8}) for (int active : {32, 16, 8}) {
        k<<<148, 32 * warps>>>(out, cyc, active, iters, chains);
        k<<<148, 32 * warps>>>(out, cyc, active, iters, chains);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        const double dfma_per_warp = (double)iters * 16 * chains;
        printf("warps/SM %d chains %d active %2d: %.2f cycles per DFMA warp-instr (per warp), %.2f DFMA/cycle/SM\n", warps, chains, active,
               h[0] / dfma_per_warp, dfma_per_warp * warps / h[0]);
    }
    return 0;
}
