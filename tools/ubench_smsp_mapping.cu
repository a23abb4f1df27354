// Microbenchmark: how are the warps of SMALL blocks spread over the four SM sub-partitions (each with its own FP64 pipe)?
// k blocks of w warps per SM, every warp runs 8 independent DFMA chains (enough ILP to saturate one sub-partition's pipe
// alone: 0.47 DFMA warp-instr/cycle, tools/ubench_fp64_halfwarp.cu).  DFMA/cycle/SM = 0.47 x (sub-partitions in use).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub2 tools/ubench_smsp_mapping.cu && /tmp/ub2
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, unsigned* smid, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; unsigned id; asm("mov.u32 %0, %%smid;" : "=r"(id)); smid[blockIdx.x] = id; }
}
int main() {
    double* out; long long* cyc; unsigned* smid;
    cudaMalloc(&out, 148 * 64 * 256 * sizeof(double)); cudaMalloc(&cyc, 148 * 64 * sizeof(long long)); cudaMalloc(&smid, 148 * 64 * sizeof(unsigned));
    const int iters = 2000;
    for (int w : {1, 2, 4}) for (int kb : {1, 2, 3, 4, 6, 8}) {
        if (w * kb > 16) continue;
        const int grid = 148 * kb;
        k<<<grid, 32 * w>>>(out, cyc, smid, iters);
        k<<<grid, 32 * w>>>(out, cyc, smid, iters);
        cudaDeviceSynchronize();
        static long long h[148 * 64]; static unsigned sm[148 * 64];
        cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(sm, smid, grid * sizeof(unsigned), cudaMemcpyDeviceToHost);
        int per_sm[256] = {0}; int mx = 0; long long worst = 0; double avg = 0;
        for (int b = 0; b < grid; ++b) { per_sm[sm[b] & 255]++; if (h[b] > worst) worst = h[b]; avg += h[b]; }
        for (int s = 0; s < 256; ++s) if (per_sm[s] > mx) mx = per_sm[s];
        avg /= grid;
        const double dfma_per_warp = (double)iters * 16 * 8;
        printf("warps/block %d blocks/SM %d (max seen on one SM %d): avg %.0f cycles, %.2f DFMA warp-instr/cycle/SM (by avg block time)\n", w, kb, mx,
               avg, dfma_per_warp * w * kb / avg);
    }
    return 0;
}
