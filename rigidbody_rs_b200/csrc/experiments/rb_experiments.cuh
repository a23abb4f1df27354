// experiments/rb_experiments.cuh -- measured dead ends, kept buildable but OUT of the product headers.
//
// Neither file of this directory is part of the default build, nor of the source bundle the run-time compiler embeds
// (Makefile: JIT_SRCS).  They are compiled only with -DRB_STREAM=1 / -DRB_PREFETCH=1 (make EXTRA_NVFLAGS=...), which
// is how the numbers in profiles/r1_kbench_streaming.jsonl and profiles/r1_kbench_register_prefetch.jsonl were made:
//   * rb_stream_kernel: persistent, TMA-fed (cp.async.bulk + mbarrier) RNEA / FD -- 0.958 vs 0.811 ms, 1.49 vs 1.30 ms;
//   * rb_rnea_pf_kernel / rb_fd_pf_kernel: persistent kernels with register prefetch -- 0.856 vs 0.801, 1.403 vs 1.286 ms.
// Included by rb_kernels.cuh (after the product kernels) only when one of the two macros is set.
#pragma once
#include "rb_tma.cuh"

// ------------------------------------------------------------------ persistent variants with register prefetch (experiment)
// A thread walks states s, s + T, s + 2T, ... and issues the loads of its NEXT state before computing the current one,
// so every warp always has arithmetic to overlap its own memory latency (the one-shot kernels rely on other warps).
#if RB_PREFETCH
#ifndef RB_PF_MINB_RNEA
#define RB_PF_MINB_RNEA 3
#endif
#ifndef RB_PF_MINB_FD
#define RB_PF_MINB_FD 3
#endif
template <class M>
__global__ void __launch_bounds__(RB_BLOCK, RB_PF_MINB_RNEA)
rb_rnea_pf_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
                  const double* __restrict__ ddq, double* __restrict__ tau, size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N];
    rb_load<N>(q, ld, s, a); rb_load<N>(dq, ld, s, b); rb_load<N>(ddq, ld, s, c);
    while (true) {
        const size_t s2 = s + nthr;
        const bool more = s2 < B;
        double a2[N], b2[N], c2[N];
        if (more) { rb_load<N>(q, ld, s2, a2); rb_load<N>(dq, ld, s2, b2); rb_load<N>(ddq, ld, s2, c2); }
        double sn[N], cs[N], t[N];
        rb_sincos_all<N>(a, sn, cs);
        rb_rnea<M, true>(p, sn, cs, b, c, t);
        rb_store<N>(tau, ld, s, t);
        if (!more) break;
#pragma unroll
        for (int i = 0; i < N; ++i) { a[i] = a2[i]; b[i] = b2[i]; c[i] = c2[i]; }
        s = s2;
    }
}
template <class M>
__global__ void __launch_bounds__(RB_BLOCK, RB_PF_MINB_FD)
rb_fd_pf_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
                const double* __restrict__ tau, double* __restrict__ qdd, size_t B, size_t ld, int* __restrict__ status) {
    constexpr int N = M::N;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N];
    rb_load<N>(q, ld, s, a); rb_load<N>(dq, ld, s, b); rb_load<N>(tau, ld, s, c);
    bool all_ok = true;
    while (true) {
        const size_t s2 = s + nthr;
        const bool more = s2 < B;
        double a2[N], b2[N], c2[N];
        if (more) { rb_load<N>(q, ld, s2, a2); rb_load<N>(dq, ld, s2, b2); rb_load<N>(tau, ld, s2, c2); }
        double sn[N], cs[N], x[N];
        rb_sincos_all<N>(a, sn, cs);
        const bool ok = rb_forward_dynamics<M>(p, sn, cs, b, c, x);
        if (!ok) {
            all_ok = false;
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = rb_nan<double>();
        }
        rb_store<N>(qdd, ld, s, x);
        if (!more) break;
#pragma unroll
        for (int i = 0; i < N; ++i) { a[i] = a2[i]; b[i] = b2[i]; c[i] = c2[i]; }
        s = s2;
    }
    if (!all_ok) atomicOr(status, RB_STATUS_NOT_SPD);
}
#endif

// ------------------------------------------------------------------ streaming (persistent, TMA-fed) RNEA / FD
// The one-tile-per-block kernels above leave each warp's 21 input loads exposed at the start of its life, so
// HBM latency is hidden only by other resident warps -- and registers cap those at 16-20 per SM.  Here a
// persistent block walks tiles of RB_BLOCK states; one elected thread asks the TMA unit to bulk-copy the
// next tile's 3N input rows (RB_BLOCK*8 = 1 KiB contiguous each) into a 2-stage shared-memory ring while all
// warps compute the current tile, and an mbarrier (transaction bytes) says when a stage has landed.  The
// FP64 pipe then sees compute-phase warps only.  Requires 16-byte aligned rows (pointers % 16, ld % 2).
#define RB_STAGES 2
template <class M, int MINB, bool IS_FD>
__global__ void __launch_bounds__(RB_BLOCK, MINB)
rb_stream_kernel(const __grid_constant__ typename M::Param p, const RB_R* __restrict__ in0, const RB_R* __restrict__ in1,
                 const RB_R* __restrict__ in2, RB_R* __restrict__ out, unsigned num_tiles, size_t ld, int* __restrict__ status) {
    constexpr int N = M::N, ROWS = 3 * N;
    extern __shared__ __align__(128) double rb_stage[];      // [RB_STAGES][ROWS][RB_BLOCK]
    __shared__ __align__(8) uint64_t bar[RB_STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < RB_STAGES; ++k) rb_mbar_init(&bar[k], 1);
        rb_fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](unsigned tile, int st) {
        const size_t s0 = (size_t)tile * RB_BLOCK;
        RB_R* dst = rb_stage + (size_t)st * ROWS * RB_BLOCK;
        rb_mbar_expect_tx(&bar[st], ROWS * RB_BLOCK * (uint32_t)sizeof(RB_R));
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const RB_R* src = (r < N ? in0 : (r < 2 * N ? in1 : in2)) + (size_t)(r % N) * ld + s0;
            rb_bulk_g2s(dst + r * RB_BLOCK, src, RB_BLOCK * (uint32_t)sizeof(RB_R), &bar[st]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < RB_STAGES; ++k) {
            const unsigned t = blockIdx.x + k * gridDim.x;
            if (t < num_tiles) issue(t, k);
        }
    }
    bool ok = true;
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int st = it % RB_STAGES;
        rb_mbar_wait(&bar[st], (it / RB_STAGES) & 1);
        const RB_R* src = rb_stage + (size_t)st * ROWS * RB_BLOCK + tid;
        RB_R a[N], b[N], c[N], sn[N], cs[N], x[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            a[i] = src[i * RB_BLOCK]; b[i] = src[(N + i) * RB_BLOCK]; c[i] = src[(2 * N + i) * RB_BLOCK];
        }
        __syncthreads();                                     // every thread has drained this stage
        if (tid == 0) {
            const unsigned nxt = tile + RB_STAGES * gridDim.x;
            if (nxt < num_tiles) issue(nxt, st);
        }
        rb_sincos_all<N>(a, sn, cs);
        if constexpr (IS_FD) {
            if (!rb_forward_dynamics<M>(p, sn, cs, b, c, x)) {
                ok = false;
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = rb_nan<RB_R>();
            }
        } else {
            rb_rnea<M, true>(p, sn, cs, b, c, x);
        }
        rb_store<N>(out, ld, (size_t)tile * RB_BLOCK + tid, x);
    }
    if constexpr (IS_FD) { if (!ok) atomicOr(status, RB_STATUS_NOT_SPD); }
}


#ifndef RB_DEVICE_ONLY
// Launch hooks called by RbLaunch<M>::rnea / fd when an experiment macro is set: `*done` = states served here (the
// product kernel takes the rest).
template <class M>
struct RbExperimentLaunch {
    using P = typename M::Param;
    // Full tiles go through the persistent TMA-fed kernel when rows are 16-byte aligned; the ragged tail
    // (and unaligned or tiny batches) through the one-tile-per-block kernel.
    template <int MINB, bool IS_FD>
    static cudaError_t stream3(const P& p, const double* a, const double* b, const double* c, double* out,
                               size_t B, size_t ld, int* status, cudaStream_t st, int sms, size_t* done) {
        *done = 0;
#if RB_STREAM
        const bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15u) == 0 && (ld & 1u) == 0;
        const size_t tiles = B / RB_BLOCK;
        const unsigned cap = (unsigned)sms * MINB;
        if (!aligned || tiles < 2 * (size_t)cap || tiles > 0xFFFFFFF0u) return cudaSuccess;
        auto k = rb_stream_kernel<M, MINB, IS_FD>;
        constexpr size_t smem = (size_t)RB_STAGES * 3 * M::N * RB_BLOCK * sizeof(double);
        static std::atomic<bool> configured{false};
        if (!configured.load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            configured.store(true, std::memory_order_release);
        }
        k<<<cap, RB_BLOCK, smem, st>>>(p, a, b, c, out, (unsigned)tiles, ld, status);
        *done = tiles * RB_BLOCK;
        return cudaGetLastError();
#else
        return cudaSuccess;
#endif
    }
    static cudaError_t rnea(const P& p, const double* q, const double* dq, const double* ddq, double* tau,
                            size_t B, size_t ld, cudaStream_t st, int sms, size_t* done) {
        *done = 0;
#if RB_PREFETCH
        if constexpr (std::is_same<typename M::Real, double>::value) {
            const size_t cap = (size_t)sms * RB_PF_MINB_RNEA, want = (B + RB_BLOCK - 1) / RB_BLOCK;
            rb_rnea_pf_kernel<M><<<(unsigned)(want < cap ? want : cap), RB_BLOCK, 0, st>>>(p, q, dq, ddq, tau, B, ld);
            *done = B;
            return cudaGetLastError();
        }
#endif
        return stream3<RB_MINB_RNEA, false>(p, q, dq, ddq, tau, B, ld, nullptr, st, sms, done);
    }
    static cudaError_t fd(const P& p, const double* q, const double* dq, const double* tau, double* qdd,
                          size_t B, size_t ld, int* status, cudaStream_t st, int sms, size_t* done) {
        *done = 0;
#if RB_PREFETCH
        if constexpr (std::is_same<typename M::Real, double>::value) {
            const size_t cap = (size_t)sms * RB_PF_MINB_FD, want = (B + RB_BLOCK - 1) / RB_BLOCK;
            rb_fd_pf_kernel<M><<<(unsigned)(want < cap ? want : cap), RB_BLOCK, 0, st>>>(p, q, dq, tau, qdd, B, ld, status);
            *done = B;
            return cudaGetLastError();
        }
#endif
        return stream3<RB_MINB_FD, true>(p, q, dq, tau, qdd, B, ld, status, st, sms, done);
    }
};
#endif  // RB_DEVICE_ONLY
