// rb_kernels_warp.cu -- forward dynamics of long chains (13..32 joints): one WARP per state, one LANE per joint.
//
// A 32-joint state needs a 528-entry mass matrix: no thread can hold it, and streaming it through HBM made the
// first long-chain path spend its time on memory and on a 400 KB unrolled CRBA.  Here a warp owns a state and
// nothing but q, dq, tau (in) and qdd (out) ever leaves the SM:
//
//   1. lane i builds its joint's local transform X_i(q_i) = (R_p Rz(q_i), t)      (joint.rs:36-38; one sincos per lane)
//   2. a 5-step prefix product over the lanes gives every joint's world pose (R_i, p_i)
//   3. with everything expressed in the world frame about the world origin, the recursions of rnea
//      (multibody.rs:122-150) and of crba's composite inertias (multibody.rs:157-171) become plain sums along the
//      chain: prefix sums for omega, alpha and the linear acceleration, suffix sums for the link wrenches and for
//      the 10-parameter inertias (m, m c, I about the origin).  bias_i = s_i . f^c_i with the joint screw
//      s_i = (z_i ; p_i x z_i)
//   4. H[j][i] = s_j . (I^c_i s_i) for j <= i: lane i keeps I^c_i s_i in 6 registers, the 32 screws are broadcast
//      from shared memory, and lane i ends up with row i of the (symmetric) matrix in registers -- the layout the
//      elimination wants, so H is never written anywhere
//   5. right-looking LDL^T with lane = row: at step k every lane writes its scaled entry of column k to shared
//      memory (one store), reads the column back as broadcasts and updates its row; the rhs rides along; the back
//      substitution walks the stored columns of L.
//
// Same result as  qdd = solve(sym(crba(q)), tau - rnea(q, dq, 0))   up to rounding: the world-frame
// sums associate differently from the reference's link-frame recursion (measured 5e-14 relative on the 32-joint
// chain, cond(H) ~ 1e4).
#include "rb_kernels.cuh"
#include "rb_util.cuh"

#ifndef RBW_WARPS
#define RBW_WARPS 12
#endif
#ifndef RBW_GROUP
#define RBW_GROUP 4                       // states a warp stages at once (32-byte runs of the joint-major arrays)
#endif
#ifndef RBW_MODEL_SMEM
#define RBW_MODEL_SMEM 0                  // 1 = per-lane model constants re-read from shared memory (frees ~46 registers)
#endif
#define RBW_LDL 34                        // row stride of the stored L columns: even, so pairs are 16-byte aligned
#define RBW_IOS (RBW_GROUP + 1)           // padded stride of the staging rows

namespace {
constexpr unsigned FULL = 0xffffffffu;
constexpr int RBW_PER_WARP = 32 * RBW_LDL + 32 * 6 + 4 * 32 * RBW_IOS;   // doubles of shared memory per warp
constexpr int RBW_MODEL_DOUBLES = RBW_MODEL_SMEM ? 23 * 32 : 0;          // block-shared copy of the per-lane constants

__device__ __forceinline__ double up(double v, int d) { return __shfl_up_sync(FULL, v, d); }
__device__ __forceinline__ double dn(double v, int d) { return __shfl_down_sync(FULL, v, d); }

template <int K>
__device__ __forceinline__ void prefix_sum(double (&x)[K], int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const double t = up(x[c], d);
            if (lane >= d) x[c] += t;
        }
    }
}
template <int K>
__device__ __forceinline__ void suffix_sum(double (&x)[K], int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const double t = dn(x[c], d);
            if (lane + d < 32) x[c] += t;
        }
    }
}
__device__ __forceinline__ void cross(const double (&a)[3], const double (&b)[3], double (&o)[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// o += a x b
__device__ __forceinline__ void cross_acc(const double (&a)[3], const double (&b)[3], double (&o)[3]) {
    o[0] = fma(a[1], b[2], fma(-a[2], b[1], o[0]));
    o[1] = fma(a[2], b[0], fma(-a[0], b[2], o[1]));
    o[2] = fma(a[0], b[1], fma(-a[1], b[0], o[2]));
}
// symmetric 3x3 (xx xy xz yy yz zz) times vector
__device__ __forceinline__ void sym_mul(const double (&S)[6], const double (&v)[3], double (&o)[3]) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}

__global__ void __launch_bounds__(32 * RBW_WARPS, 1)
rbw_fd_kernel(const double* __restrict__ model, int n, const double* __restrict__ q, const double* __restrict__ dq,
              const double* __restrict__ tau, double* __restrict__ qdd, size_t B, size_t ld, int* __restrict__ status) {
    extern __shared__ __align__(16) double rbw_sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* Lt = rbw_sm + (size_t)w * RBW_PER_WARP;          // [32][RBW_LDL]  column k of L at row k
    double* Sb = Lt + 32 * RBW_LDL;                          // [32][6]        joint screws (z ; p x z)
    double* io = Sb + 32 * 6;                                // [3][32][RBW_IOS] q, dq, tau of the staged states
    double* ob = io + 3 * 32 * RBW_IOS;                      // [32][RBW_IOS]  qdd of the staged states
    const bool act = lane < n;

    // this lane's joint: fixed placement, link inertia about the joint origin (rb_model.h RbJointK), composite mass
    // index: 0-8 R_p, 9-11 t, 12 m, 13-15 h = m c, 16-21 I_o (xx xy xz yy yz zz), 22 composite mass
    double mdl[23] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (act) {
        const double* row = model + (size_t)lane * 24;
#pragma unroll
        for (int e = 0; e < 13; ++e) mdl[e] = row[e];
#pragma unroll
        for (int e = 0; e < 9; ++e) mdl[13 + e] = row[14 + e];
    }
    {
        double mc[1] = {mdl[12]};
        suffix_sum<1>(mc, lane);                             // composite mass of the sub-chain from joint i on
        mdl[22] = mc[0];
    }
#if RBW_MODEL_SMEM
    double* msm = rbw_sm + (size_t)RBW_WARPS * RBW_PER_WARP;
    if (w == 0) {
#pragma unroll
        for (int e = 0; e < 23; ++e) msm[e * 32 + lane] = mdl[e];
    }
    __syncthreads();
#define MDL(e) msm[(e) * 32 + lane]
#else
#define MDL(e) mdl[e]
#endif
    const double g[3] = {model[(size_t)n * 24], model[(size_t)n * 24 + 1], model[(size_t)n * 24 + 2]};

    const size_t groups = (B + RBW_GROUP - 1) / RBW_GROUP;
    bool all_ok = true;
    for (size_t grp = (size_t)blockIdx.x * RBW_WARPS + w; grp < groups; grp += (size_t)gridDim.x * RBW_WARPS) {
        const size_t s0 = grp * RBW_GROUP;
        {   // stage q, dq, tau of RBW_GROUP states: lane -> (joint lane / GROUP + 8 it, state lane % GROUP)
            const int sj = lane / RBW_GROUP, ss = lane % RBW_GROUP;
            const bool sin = s0 + ss < B;
#pragma unroll
            for (int it = 0; it < 32 / (32 / RBW_GROUP); ++it) {
                const int i = it * (32 / RBW_GROUP) + sj;
                const bool ld_ok = sin && i < n;
                const size_t off = (size_t)i * ld + s0 + ss;
                io[(0 * 32 + i) * RBW_IOS + ss] = ld_ok ? __ldcs(q + off) : 0.0;
                io[(1 * 32 + i) * RBW_IOS + ss] = ld_ok ? __ldcs(dq + off) : 0.0;
                io[(2 * 32 + i) * RBW_IOS + ss] = ld_ok ? __ldcs(tau + off) : 0.0;
            }
        }
        __syncwarp();
        const int in_group = (int)(B - s0 < RBW_GROUP ? B - s0 : RBW_GROUP);
        for (int st = 0; st < in_group; ++st) {
            const double qi = io[(0 * 32 + lane) * RBW_IOS + st];
            const double dqi = io[(1 * 32 + lane) * RBW_IOS + st];
            const double ti = io[(2 * 32 + lane) * RBW_IOS + st];

            // ---- 1, 2: world pose of every joint frame
            double sn, cs;
            sincos(qi, &sn, &cs);
            double R[9], p[3] = {MDL(9), MDL(10), MDL(11)};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double r0 = MDL(3 * r), r1 = MDL(3 * r + 1);
                R[3 * r + 0] = fma(cs, r0, sn * r1);
                R[3 * r + 1] = fma(cs, r1, -sn * r0);
                R[3 * r + 2] = MDL(3 * r + 2);
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double A[9], u[3];
#pragma unroll
                for (int e = 0; e < 9; ++e) A[e] = up(R[e], d);
#pragma unroll
                for (int e = 0; e < 3; ++e) u[e] = up(p[e], d);
                if (lane >= d) {                             // (A, u) o (R, p) = (A R, u + A p)
                    double nR[9], np[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        np[r] = fma(A[3 * r], p[0], fma(A[3 * r + 1], p[1], fma(A[3 * r + 2], p[2], u[r])));
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            nR[3 * r + c] = fma(A[3 * r], R[c], fma(A[3 * r + 1], R[3 + c], A[3 * r + 2] * R[6 + c]));
                    }
#pragma unroll
                    for (int e = 0; e < 9; ++e) R[e] = nR[e];
#pragma unroll
                    for (int e = 0; e < 3; ++e) p[e] = np[e];
                }
            }
            // joint screw about the world origin: s = (z ; p x z); idle lanes get a zero screw
            const double z[3] = {act ? R[2] : 0.0, act ? R[5] : 0.0, act ? R[8] : 0.0};
            double v[3];
            cross(p, z, v);

            // ---- 3a: velocities and velocity-product accelerations (ddq = 0), prefix sums along the chain
            const double zq[3] = {z[0] * dqi, z[1] * dqi, z[2] * dqi};
            double om[3] = {zq[0], zq[1], zq[2]};
            prefix_sum<3>(om, lane);
            double al[3];
            cross(om, zq, al);                               // omega_{i-1} x z dq = omega_i x z dq
            prefix_sum<3>(al, lane);
            double acc[3];
            {
                double omp[3], alp[3], d[3], w1[3];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const double a = up(om[e], 1), b = up(al[e], 1), c = up(p[e], 1);
                    omp[e] = lane ? a : 0.0;
                    alp[e] = lane ? b : 0.0;
                    d[e] = p[e] - (lane ? c : 0.0);
                }
                cross(omp, d, w1);
                cross(alp, d, acc);
                cross_acc(omp, w1, acc);
#pragma unroll
                for (int e = 0; e < 3; ++e) acc[e] += lane ? 0.0 : g[e];     // base acceleration (multibody.rs:118)
            }
            prefix_sum<3>(acc, lane);                        // classical acceleration of joint origin i

            // ---- 3b: link inertia in world orientation, link wrench about the world origin
            double hw[3], Iw[6];
            const double m = MDL(12);
            {
                const double h[3] = {MDL(13), MDL(14), MDL(15)};
                const double Io[6] = {MDL(16), MDL(17), MDL(18), MDL(19), MDL(20), MDL(21)};
#pragma unroll
                for (int r = 0; r < 3; ++r) hw[r] = R[3 * r] * h[0] + R[3 * r + 1] * h[1] + R[3 * r + 2] * h[2];
                double T[9];                                 // T = R Io
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
                    T[3 * r + 0] = a * Io[0] + b * Io[1] + c * Io[2];
                    T[3 * r + 1] = a * Io[1] + b * Io[3] + c * Io[4];
                    T[3 * r + 2] = a * Io[2] + b * Io[4] + c * Io[5];
                }
                Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
                Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
                Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
                Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
                Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
                Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
            }
            double fw[6];                                    // (F ; N about the world origin)
            {
                double w1[3], F[3], N[3], Iom[3];
                cross(om, hw, w1);
                F[0] = m * acc[0]; F[1] = m * acc[1]; F[2] = m * acc[2];
                cross_acc(al, hw, F);
                cross_acc(om, w1, F);
                sym_mul(Iw, al, N);
                sym_mul(Iw, om, Iom);
                cross_acc(om, Iom, N);
                cross_acc(hw, acc, N);
                cross_acc(p, F, N);
                fw[0] = F[0]; fw[1] = F[1]; fw[2] = F[2]; fw[3] = N[0]; fw[4] = N[1]; fw[5] = N[2];
            }
            suffix_sum<6>(fw, lane);
            const double bias = z[0] * fw[3] + z[1] * fw[4] + z[2] * fw[5] + v[0] * fw[0] + v[1] * fw[1] + v[2] * fw[2];
            double b = ti - bias;                            // idle lanes: 0

            // ---- 3c: composite inertias about the world origin (first moment, 6 inertia entries; mass is a constant)
            double ci[9];
            {
                const double u[3] = {fma(0.5 * m, p[0], hw[0]), fma(0.5 * m, p[1], hw[1]), fma(0.5 * m, p[2], hw[2])};
                const double pu2 = 2.0 * (p[0] * u[0] + p[1] * u[1] + p[2] * u[2]);
                ci[0] = fma(m, p[0], hw[0]); ci[1] = fma(m, p[1], hw[1]); ci[2] = fma(m, p[2], hw[2]);
                ci[3] = Iw[0] - 2.0 * p[0] * u[0] + pu2;
                ci[4] = Iw[1] - (p[0] * u[1] + u[0] * p[1]);
                ci[5] = Iw[2] - (p[0] * u[2] + u[0] * p[2]);
                ci[6] = Iw[3] - 2.0 * p[1] * u[1] + pu2;
                ci[7] = Iw[4] - (p[1] * u[2] + u[1] * p[2]);
                ci[8] = Iw[5] - 2.0 * p[2] * u[2] + pu2;
            }
            suffix_sum<9>(ci, lane);
            double Ff[3], Fn[3];                             // I^c_i s_i = (force ; moment about the origin)
            {
                const double Hc[3] = {ci[0], ci[1], ci[2]};
                const double IO[6] = {ci[3], ci[4], ci[5], ci[6], ci[7], ci[8]};
                const double mc = MDL(22);
                Ff[0] = mc * v[0]; Ff[1] = mc * v[1]; Ff[2] = mc * v[2];
                cross_acc(z, Hc, Ff);
                sym_mul(IO, z, Fn);
                cross_acc(Hc, v, Fn);
            }

            // ---- 4: row `lane` of H from the broadcast screws (entries j <= lane are H[j][lane]; the rest is never read)
            {
                double2* S2 = reinterpret_cast<double2*>(Sb + lane * 6);
                S2[0] = make_double2(z[0], z[1]);
                S2[1] = make_double2(z[2], v[0]);
                S2[2] = make_double2(v[1], v[2]);
            }
            __syncwarp();
            double a[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const double2* S2 = reinterpret_cast<const double2*>(Sb + j * 6);
                const double2 s01 = S2[0], s23 = S2[1], s45 = S2[2];
                const double hji = fma(s01.x, Fn[0], fma(s01.y, Fn[1], fma(s23.x, Fn[2], fma(s23.y, Ff[0], fma(s45.x, Ff[1], s45.y * Ff[2])))));
                a[j] = hji;
            }
            if (n < 32) {                                    // idle lanes: identity rows
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = (!act && j == lane) ? 1.0 : a[j];
            }

            // ---- 5: LDL^T, lane = row; forward substitution rides along
            bool ok = true;
            double mydinv = 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                Lt[k * RBW_LDL + lane] = a[k];               // column k of L D (entries of lanes < k are never read)
                const double bk = __shfl_sync(FULL, b, k);
                __syncwarp();
                const double d = Lt[k * RBW_LDL + k];
                ok = ok && (d > 0.0);
                const double dinv = rb_rcp_pos(d);
                if (lane == k) mydinv = dinv;
                const double nl = lane > k ? -a[k] * dinv : 0.0;     // -l_rk
                b = fma(nl, bk, b);
                if (((k + 1) & 1) && k + 1 < 32) a[(k + 1) & 31] = fma(nl, Lt[k * RBW_LDL + k + 1], a[(k + 1) & 31]);
#pragma unroll
                for (int i = (k + 2) & ~1; i < 32; i += 2) {
                    const double2 c2 = *reinterpret_cast<const double2*>(Lt + k * RBW_LDL + i);
                    a[i] = fma(nl, c2.x, a[i]);
                    a[i + 1] = fma(nl, c2.y, a[i + 1]);
                }
            }
            // lane k: b = y_k (L y = rhs).  x = L^-T D^-1 y:  x_k = (y_k - sum_{i>k} (l_ik d_k) x_i) / d_k
            double x = b;
#pragma unroll
            for (int i = 31; i >= 1; --i) {
                const double xi = __shfl_sync(FULL, lane == i ? x * mydinv : x, i);
                const double cik = Lt[lane * RBW_LDL + i];
                if (lane == i) x = xi;
                if (lane < i) x = fma(-cik, xi, x);
            }
            if (lane == 0) x *= mydinv;
            all_ok = all_ok && ok;
            ob[lane * RBW_IOS + st] = ok ? x : rb_nan<double>();
            __syncwarp();
        }
        {   // coalesced store of the staged results
            const int sj = lane / RBW_GROUP, ss = lane % RBW_GROUP;
            if (s0 + ss < B) {
#pragma unroll
                for (int it = 0; it < 32 / (32 / RBW_GROUP); ++it) {
                    const int i = it * (32 / RBW_GROUP) + sj;
                    if (i < n) __stcs(qdd + (size_t)i * ld + s0 + ss, ob[i * RBW_IOS + ss]);
                }
            }
        }
        __syncwarp();
    }
    if (!all_ok && lane == 0) atomicOr(status, RB_STATUS_NOT_SPD);
}
}  // namespace

// qdd = FD(q, dq, tau) for a chain of n <= 32 joints whose model rows (rb_model.h layout) are at `model` on the device.
cudaError_t rb_launch_warp_fd(const double* model, int n, const double* q, const double* dq, const double* tau,
                              double* qdd, size_t B, size_t ld, int* status, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    if (n < 1 || n > 32) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr size_t smem = ((size_t)RBW_WARPS * RBW_PER_WARP + RBW_MODEL_DOUBLES) * sizeof(double);
    static bool configured[64] = {false};
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(rbw_fd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const size_t groups = (B + RBW_GROUP - 1) / RBW_GROUP;
    const size_t want = (groups + RBW_WARPS - 1) / RBW_WARPS;
    const unsigned grid = (unsigned)(want < (size_t)sms ? want : (size_t)sms);
    rbw_fd_kernel<<<grid, 32 * RBW_WARPS, smem, st>>>(model, n, q, dq, tau, qdd, B, ld, status);
    return cudaGetLastError();
}
