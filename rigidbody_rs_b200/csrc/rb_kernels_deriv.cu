// rb_kernels_deriv.cu -- analytical-derivative kernels (rb_deriv.cuh), one instantiation per chain length
// (inverse dynamics: 1..32 joints; forward dynamics, which also unrolls FD itself: 1..12).
// Unlike the other register-resident kernels these are NOT specialised on the model: the algebra runs in the world
// frame, where the zeros of a particular robot's fixed rotations buy little, and real loops over the joints keep the
// code small and the register allocation sane.  The model travels as a __grid_constant__ parameter (<= 6.2 KB).
#include "rb_kernels.cuh"
#include "rb_deriv.cuh"
#include "rb_util.cuh"

#ifndef RB_MINB_DERIV
#define RB_MINB_DERIV 2
#endif

// out: [2 N^2][ld]: d tau_r / d q_c at entry r + N c, then d tau_r / d dq_c at N^2 + r + N c.
template <int N>
__global__ void __launch_bounds__(RB_BLOCK, RB_MINB_DERIV)
rb_rnea_deriv_kernel(const __grid_constant__ RbModelK<N> p, const double* __restrict__ q, const double* __restrict__ dq,
                     const double* __restrict__ ddq, double* __restrict__ out, size_t B, size_t ld) {
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], c[N], sn[N], cs[N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_load<N>(dq, ld, s, b);
    rb_load<N>(ddq, ld, s, c);
    double* o = out + s;
    rb_rnea_derivatives<N>(p, sn, cs, b, c, [&](int blk, int r, int col, double v) { __stcs(o + (size_t)(blk * N * N + r + N * col) * ld, v); });
}

// out: [3 N^2][ld]: d qdd / d q, d qdd / d dq, H^-1 (= d qdd / d tau), each entry r + N c.
//   qdd = FD(q, dq, tau);  d qdd / d x = -H^-1 (d rnea / d x at ddq = qdd).
// The two d rnea matrices pass through `out` (written, then read back column by column by the same thread); the
// factor of H is rebuilt after the derivative sweep rather than kept alive across it (35 doubles of registers).
template <int N>
__global__ void __launch_bounds__(RB_BLOCK, RB_MINB_DERIV)
rb_fd_deriv_kernel(const __grid_constant__ RbModelK<N> p, const double* __restrict__ q, const double* __restrict__ dq,
                   const double* __restrict__ tau, double* __restrict__ out, size_t B, size_t ld, int* __restrict__ status) {
    using M = RtModel<N>;
    const size_t s = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    if (s >= B) return;
    double a[N], b[N], x[N], sn[N], cs[N];
    rb_load<N>(q, ld, s, a);
    rb_sincos_all<N>(a, sn, cs);
    rb_load<N>(dq, ld, s, b);
    rb_load<N>(tau, ld, s, a);
    rb_forward_dynamics<M>(p, sn, cs, b, a, x);              // qdd
    double* o = out + s;
    rb_rnea_derivatives<N>(p, sn, cs, b, x, [&](int blk, int r, int col, double v) { o[(size_t)(blk * N * N + r + N * col) * ld] = v; });
    double H[N][N], dinv[N];
    rb_crba<M>(p, sn, cs, H);
    const bool ok = rb_ldlt_factor<N>(H, dinv);
    if (!ok) atomicOr(status, RB_STATUS_NOT_SPD);
#pragma unroll 1
    for (int c = 0; c < 3 * N; ++c) {                        // columns of the three blocks
        double y[N];
#pragma unroll
        for (int r = 0; r < N; ++r) y[r] = c < 2 * N ? -o[(size_t)(r + N * c) * ld] : (r == c - 2 * N ? 1.0 : 0.0);
        rb_ldlt_apply<N>(H, dinv, y);
#pragma unroll
        for (int r = 0; r < N; ++r) __stcs(o + (size_t)(r + N * c) * ld, ok ? y[r] : rb_nan<double>());
    }
}

namespace {
unsigned dgrid(size_t B) { return (unsigned)((B + RB_BLOCK - 1) / RB_BLOCK); }
template <int N>
cudaError_t launch_rnea(const double* flat, const double* q, const double* dq, const double* ddq, double* out, size_t B, size_t ld, cudaStream_t st) {
    RbModelK<N> p;
    memcpy(&p, flat, sizeof(p));
    rb_rnea_deriv_kernel<N><<<dgrid(B), RB_BLOCK, 0, st>>>(p, q, dq, ddq, out, B, ld);
    return cudaGetLastError();
}
template <int N>
cudaError_t launch_fd(const double* flat, const double* q, const double* dq, const double* tau, double* out, size_t B, size_t ld, int* status, cudaStream_t st) {
    RbModelK<N> p;
    memcpy(&p, flat, sizeof(p));
    rb_fd_deriv_kernel<N><<<dgrid(B), RB_BLOCK, 0, st>>>(p, q, dq, tau, out, B, ld, status);
    return cudaGetLastError();
}
}  // namespace

#define RB_DERIV_CASES(F, ...) \
    switch (n) { \
        case 1: return F<1>(__VA_ARGS__); case 2: return F<2>(__VA_ARGS__); case 3: return F<3>(__VA_ARGS__); \
        case 4: return F<4>(__VA_ARGS__); case 5: return F<5>(__VA_ARGS__); case 6: return F<6>(__VA_ARGS__); \
        case 7: return F<7>(__VA_ARGS__); case 8: return F<8>(__VA_ARGS__); case 9: return F<9>(__VA_ARGS__); \
        case 10: return F<10>(__VA_ARGS__); case 11: return F<11>(__VA_ARGS__); case 12: return F<12>(__VA_ARGS__); \
        default: return cudaErrorInvalidValue; \
    }

cudaError_t rb_launch_rnea_deriv(int n, const double* flat_model, const double* q, const double* dq, const double* ddq,
                                 double* out, size_t B, size_t ld, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    switch (n) {                                             // the rolled recursion has no per-joint state: any length
        case 13: return launch_rnea<13>(flat_model, q, dq, ddq, out, B, ld, st); case 14: return launch_rnea<14>(flat_model, q, dq, ddq, out, B, ld, st);
        case 15: return launch_rnea<15>(flat_model, q, dq, ddq, out, B, ld, st); case 16: return launch_rnea<16>(flat_model, q, dq, ddq, out, B, ld, st);
        case 17: return launch_rnea<17>(flat_model, q, dq, ddq, out, B, ld, st); case 18: return launch_rnea<18>(flat_model, q, dq, ddq, out, B, ld, st);
        case 19: return launch_rnea<19>(flat_model, q, dq, ddq, out, B, ld, st); case 20: return launch_rnea<20>(flat_model, q, dq, ddq, out, B, ld, st);
        case 21: return launch_rnea<21>(flat_model, q, dq, ddq, out, B, ld, st); case 22: return launch_rnea<22>(flat_model, q, dq, ddq, out, B, ld, st);
        case 23: return launch_rnea<23>(flat_model, q, dq, ddq, out, B, ld, st); case 24: return launch_rnea<24>(flat_model, q, dq, ddq, out, B, ld, st);
        case 25: return launch_rnea<25>(flat_model, q, dq, ddq, out, B, ld, st); case 26: return launch_rnea<26>(flat_model, q, dq, ddq, out, B, ld, st);
        case 27: return launch_rnea<27>(flat_model, q, dq, ddq, out, B, ld, st); case 28: return launch_rnea<28>(flat_model, q, dq, ddq, out, B, ld, st);
        case 29: return launch_rnea<29>(flat_model, q, dq, ddq, out, B, ld, st); case 30: return launch_rnea<30>(flat_model, q, dq, ddq, out, B, ld, st);
        case 31: return launch_rnea<31>(flat_model, q, dq, ddq, out, B, ld, st); case 32: return launch_rnea<32>(flat_model, q, dq, ddq, out, B, ld, st);
        default: break;
    }
    RB_DERIV_CASES(launch_rnea, flat_model, q, dq, ddq, out, B, ld, st)
}
cudaError_t rb_launch_fd_deriv(int n, const double* flat_model, const double* q, const double* dq, const double* tau,
                               double* out, size_t B, size_t ld, int* status, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    RB_DERIV_CASES(launch_fd, flat_model, q, dq, tau, out, B, ld, status, st)
}
