// rb_kernels_long.cuh -- unrolled kernels for chains too long to keep per-state data in registers
// (BASELINE.json configs[4]: the synthetic 32-joint serial chain).
//
// Same model-policy templates as rb_kernels.cuh (CtModel / RtModel, constants from the instruction stream or the
// constant bank, never from global memory: the first loop-based 32-joint kernels were bound by the ~14
// model/sin-cos load instructions each force transform needed).  What changes for long chains:
//   * RNEA: the 6(N-1) force components live in the thread's local frame (ptxas spills what does not fit;
//     spills are private, interleaved per thread, hence coalesced);
//   * forward dynamics: H (528 doubles at N = 32) is never held by a thread.  rb_fd_prepare_kernel streams the
//     packed upper triangle to HBM once ([k][state], coalesced) together with rhs = tau - bias, and
//     rb_launch_ldlt_tiles (rb_kernels_n.cu) factorises tiles of 32 states in shared memory.
#pragma once
#include "rb_kernels.cuh"
#include "rb_util.cuh"

// 1 = forward dynamics through rb_fd_prepare_kernel + the tile solver; 0 (default) = leave it to the run-time-n
// family's warp-per-state kernel (rb_kernels_warp.cu), which keeps H on the SM.
#ifndef RB_LONG_FD
#define RB_LONG_FD 0
#endif
#ifndef RB_MINB_LONG
#define RB_MINB_LONG 2
#endif

// Parameter block of a long-chain family: the model parameter followed by the H chunk buffer.
template <class M>
struct RbLongParam {
    typename M::Param model;
    double* hpk;             // packed upper triangles, tile-major [state / 32][N(N+1)/2][state % 32]
    size_t hpk_states;
};

template <class M>
__global__ void __launch_bounds__(RB_BLOCK, RB_MINB_LONG)
rb_long_rnea_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
                    const double* __restrict__ ddq, double* __restrict__ tau, size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N], sn[N], cs[N], t[N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_load<N>(dq, ld, s, b);
    rb_load<N>(ddq, ld, s, c);
    rb_rnea<M, true>(p, sn, cs, b, c, t);
    rb_store<N>(tau, ld, s, t);
}

// rhs = tau - rnea(q, dq, 0) -> qdd;  packed upper triangle of crba(q) -> hpk (chunk-local state index s).
template <class M>
__global__ void __launch_bounds__(RB_BLOCK, RB_MINB_LONG)
rb_fd_prepare_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
                     const double* __restrict__ tau, double* __restrict__ qdd, double* __restrict__ hpk, size_t hpk_states,
                     size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], sn[N], cs[N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    {
        double b[N], bias[N];
        rb_load<N>(dq, ld, s, b);
        rb_rnea<M, false>(p, sn, cs, b, b /*unused*/, bias);
#pragma unroll
        for (int i = 0; i < N; ++i) __stcs(qdd + (size_t)i * ld + s, __ldcs(tau + (size_t)i * ld + s) - bias[i]);
    }
    double* hp = hpk + (s >> 5) * (size_t)(N * (N + 1) / 2) * 32 + (s & 31);      // tile-major: [s / 32][k][s % 32]
    rb_crba_put<M>(p, sn, cs, [&](auto jc, auto ic, double v) {
        constexpr int J = decltype(jc)::value, I = decltype(ic)::value;
        __stcs(hp + (size_t)(J * N - J * (J - 1) / 2 + (I - J)) * 32, v);
    });
}

// H out in the reference convention (entry r + N*c, strict lower triangle 0), written entry by entry.
template <class M>
__global__ void __launch_bounds__(RB_BLOCK, RB_MINB_LONG)
rb_long_crba_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, double* __restrict__ Hout,
                    size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], sn[N], cs[N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    for (int c = 0; c < N; ++c)
        for (int r = c + 1; r < N; ++r) __stcs(Hout + (size_t)(r + N * c) * ld + s, 0.0);
    rb_crba_put<M>(p, sn, cs, [&](auto jc, auto ic, double v) {
        constexpr int J = decltype(jc)::value, I = decltype(ic)::value;
        __stcs(Hout + (size_t)(J + N * I) * ld + s, v);
    });
}

template <class M>
struct RbLaunchLong {
    using LP = RbLongParam<M>;
    static unsigned grid(size_t B) { return (unsigned)((B + RB_BLOCK - 1) / RB_BLOCK); }
    static cudaError_t rnea(const void* param, const double* q, const double* dq, const double* ddq, double* tau,
                            size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_long_rnea_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(((const LP*)param)->model, q, dq, ddq, tau, B, ld);
        return cudaGetLastError();
    }
    static cudaError_t fd(const void* param, const double* q, const double* dq, const double* tau, double* qdd,
                          size_t B, size_t ld, int* status, cudaStream_t st) {
        const LP* P = (const LP*)param;
        for (size_t off = 0; off < B; off += P->hpk_states) {
            const size_t cnt = (B - off < P->hpk_states) ? (B - off) : P->hpk_states;
            rb_fd_prepare_kernel<M><<<grid(cnt), RB_BLOCK, 0, st>>>(P->model, q + off, dq + off, tau + off, qdd + off,
                                                                   P->hpk, P->hpk_states, cnt, ld);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            e = rb_launch_ldlt_tiles(M::N, P->hpk, P->hpk_states, qdd + off, cnt, ld, status, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    static cudaError_t crba(const void* param, const double* q, double* H, size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_long_crba_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(((const LP*)param)->model, q, H, B, ld);
        return cudaGetLastError();
    }
    // fwd_kin / jac / rollout are served by the run-time-n family (null here = use the fallback table).
    static RbOps ops(const char* name) {
        RbOps o;
        o.name = name; o.n = M::N; o.param_bytes = sizeof(LP); o.shared_scratch = true;      // the H chunk buffer
        o.rnea = &rnea; o.fd = RB_LONG_FD ? &fd : nullptr; o.rnea_aos = nullptr; o.fd_aos = nullptr; o.crba = &crba; o.fwd_kin = nullptr; o.jac = nullptr; o.rollout = nullptr;
        o.rnea_f32 = nullptr; o.fd_f32 = nullptr;
        return o;
    }
};
