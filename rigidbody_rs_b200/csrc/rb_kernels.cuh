// rb_kernels.cuh -- __global__ kernels over SoA batches for an unrolled model policy M, and the RbOps
// table through which the C ABI (rb_api.cu) reaches whichever kernel family serves a chain.
//
// Data layout in HBM: joint-major SoA.  Array x of a batch is [n][ld] doubles; state s of joint i is
// x[i*ld + s].  Thread t of the grid owns state s = t: consecutive lanes read consecutive doubles, so each
// warp-level load/store is one fully used 256-byte run, and each array is streamed exactly once
// (ld.global.cs / st.global.cs, no reuse, no smem staging needed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rb_dyn.cuh"

#define RB_BLOCK 128

struct RbOps {
    const char* name;
    int n;                 // joints this table serves (0 = any, run-time n)
    size_t param_bytes;    // bytes of model parameter the launchers expect behind `param`
    cudaError_t (*rnea)(const void* param, const double* q, const double* dq, const double* ddq, double* tau,
                        size_t B, size_t ld, cudaStream_t st);
    cudaError_t (*fd)(const void* param, const double* q, const double* dq, const double* tau, double* qdd,
                      size_t B, size_t ld, int* status, cudaStream_t st);
    cudaError_t (*crba)(const void* param, const double* q, double* H, size_t B, size_t ld, cudaStream_t st);
    cudaError_t (*fwd_kin)(const void* param, const double* q, double* xyz, size_t B, size_t ld, cudaStream_t st);
    cudaError_t (*jac)(const void* param, const double* q, double* J, size_t B, size_t ld, cudaStream_t st);
    cudaError_t (*rollout)(const void* param, const double* q0, const double* dq0, const double* tau, double dt,
                           int horizon, double* q_traj, double* dq_traj, double* q_fin, double* dq_fin,
                           size_t B, size_t ld, int* status, cudaStream_t st);
};

const RbOps* rb_ops_fr3();        // compile-time FR3 model (rb_kernels_fr3.cu)
const RbOps* rb_ops_rt7();        // any 7-joint chain, run-time constants (rb_kernels_rt.cu)
const RbOps* rb_ops_generic_n();  // any chain length (rb_kernels_n.cu)
const double* rb_fr3_table();     // the 7x24 table + 3 gravity doubles the FR3 kernels were compiled for

#define RB_STATUS_NOT_SPD 1

template <int N>
RB_DI void rb_load(const double* __restrict__ x, size_t ld, size_t s, double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __ldcs(x + (size_t)i * ld + s);
}
template <int N>
RB_DI void rb_store(double* __restrict__ x, size_t ld, size_t s, const double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) __stcs(x + (size_t)i * ld + s, v[i]);
}

template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_rnea_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
               const double* __restrict__ ddq, double* __restrict__ tau, size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N], sn[N], cs[N], t[N];
    rb_load<N>(q, ld, s, a);
    rb_load<N>(dq, ld, s, b);
    rb_load<N>(ddq, ld, s, c);
    rb_sincos_all<N>(a, sn, cs);
    rb_rnea<M, true>(p, sn, cs, b, c, t);
    rb_store<N>(tau, ld, s, t);
}

template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_fd_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, const double* __restrict__ dq,
             const double* __restrict__ tau, double* __restrict__ qdd, size_t B, size_t ld, int* __restrict__ status) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N], sn[N], cs[N], x[N];
    rb_load<N>(q, ld, s, a);
    rb_load<N>(dq, ld, s, b);
    rb_load<N>(tau, ld, s, c);
    rb_sincos_all<N>(a, sn, cs);
    const bool ok = rb_forward_dynamics<M>(p, sn, cs, b, c, x);
    if (!ok) {
        atomicOr(status, RB_STATUS_NOT_SPD);
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = __longlong_as_double(0x7ff8000000000000LL);
    }
    rb_store<N>(qdd, ld, s, x);
}

// H out: reference convention, n*n entries per state, entry k = r + n*c, upper filled, strict lower 0.
template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_crba_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, double* __restrict__ Hout,
               size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], sn[N], cs[N], H[N][N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_crba<M>(p, sn, cs, H);
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) __stcs(Hout + (size_t)(r + N * c) * ld + s, r <= c ? H[r][c] : 0.0);
}

template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_fwd_kin_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, double* __restrict__ xyz,
                  size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], sn[N], cs[N], pos[3];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_fwd_kin<M>(p, sn, cs, pos);
    rb_store<3>(xyz, ld, s, pos);
}

template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_jac_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q, double* __restrict__ Jout,
              size_t B, size_t ld) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], sn[N], cs[N], J[N][6];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_jac<M>(p, sn, cs, J);
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) __stcs(Jout + (size_t)(r + 6 * c) * ld + s, J[c][r]);
}

// MPC rollout: the trajectory's (q, dq) stay in registers for the whole horizon; per step the
// kernel reads tau[t] (n doubles) and writes the new (q, dq) (2n doubles).
template <class M>
__global__ void __launch_bounds__(RB_BLOCK)
rb_rollout_kernel(const __grid_constant__ typename M::Param p, const double* __restrict__ q0, const double* __restrict__ dq0,
                  const double* __restrict__ tau, double dt, int horizon, double* __restrict__ q_traj,
                  double* __restrict__ dq_traj, double* __restrict__ q_fin, double* __restrict__ dq_fin,
                  size_t B, size_t ld, int* __restrict__ status) {
    constexpr int N = M::N;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double q[N], dq[N];
    rb_load<N>(q0, ld, s, q);
    rb_load<N>(dq0, ld, s, dq);
    bool ok = true;
    const size_t step = (size_t)N * ld;
    double u[N];
    rb_load<N>(tau, ld, s, u);
    for (int t = 0; t < horizon; ++t) {
        double sn[N], cs[N], qdd[N], un[N];
        // prefetch the next step's torques so the load latency hides behind this step's arithmetic
        if (t + 1 < horizon) rb_load<N>(tau + (size_t)(t + 1) * step, ld, s, un);
        rb_sincos_all<N>(q, sn, cs);
        ok = rb_forward_dynamics<M>(p, sn, cs, dq, u, qdd) && ok;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            dq[i] = fma(dt, qdd[i], dq[i]);
            q[i] = fma(dt, dq[i], q[i]);
        }
        if (q_traj) rb_store<N>(q_traj + (size_t)t * step, ld, s, q);
        if (dq_traj) rb_store<N>(dq_traj + (size_t)t * step, ld, s, dq);
        if (t + 1 < horizon) {
#pragma unroll
            for (int i = 0; i < N; ++i) u[i] = un[i];
        }
    }
    if (q_fin) rb_store<N>(q_fin, ld, s, q);
    if (dq_fin) rb_store<N>(dq_fin, ld, s, dq);
    if (!ok) atomicOr(status, RB_STATUS_NOT_SPD);
}

// ------------------------------------------------------------------ launchers for policy M
template <class M>
struct RbLaunch {
    using P = typename M::Param;
    static unsigned grid(size_t B) { return (unsigned)((B + RB_BLOCK - 1) / RB_BLOCK); }
    static cudaError_t rnea(const void* param, const double* q, const double* dq, const double* ddq, double* tau,
                            size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_rnea_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q, dq, ddq, tau, B, ld);
        return cudaGetLastError();
    }
    static cudaError_t fd(const void* param, const double* q, const double* dq, const double* tau, double* qdd,
                          size_t B, size_t ld, int* status, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_fd_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q, dq, tau, qdd, B, ld, status);
        return cudaGetLastError();
    }
    static cudaError_t crba(const void* param, const double* q, double* H, size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_crba_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q, H, B, ld);
        return cudaGetLastError();
    }
    static cudaError_t fwd_kin(const void* param, const double* q, double* xyz, size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_fwd_kin_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q, xyz, B, ld);
        return cudaGetLastError();
    }
    static cudaError_t jac(const void* param, const double* q, double* J, size_t B, size_t ld, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_jac_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q, J, B, ld);
        return cudaGetLastError();
    }
    static cudaError_t rollout(const void* param, const double* q0, const double* dq0, const double* tau, double dt,
                               int horizon, double* q_traj, double* dq_traj, double* q_fin, double* dq_fin,
                               size_t B, size_t ld, int* status, cudaStream_t st) {
        if (B == 0) return cudaSuccess;
        rb_rollout_kernel<M><<<grid(B), RB_BLOCK, 0, st>>>(*(const P*)param, q0, dq0, tau, dt, horizon, q_traj, dq_traj,
                                                          q_fin, dq_fin, B, ld, status);
        return cudaGetLastError();
    }
    static RbOps ops(const char* name) {
        RbOps o;
        o.name = name; o.n = M::N; o.param_bytes = sizeof(P);
        o.rnea = &rnea; o.fd = &fd; o.crba = &crba; o.fwd_kin = &fwd_kin; o.jac = &jac; o.rollout = &rollout;
        return o;
    }
};
