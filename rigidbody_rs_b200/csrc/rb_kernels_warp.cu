// rb_kernels_warp.cu -- forward dynamics of long chains (13..32 joints): the chain spread over the lanes of a warp
// (rbq_fd_kernel).
//
// A 32-joint state needs a 528-entry mass matrix: no thread can hold it, and streaming it through HBM made the
// first long-chain path spend its time on memory and on a 400 KB unrolled CRBA.  Here a warp owns four states and
// nothing but q, dq, tau (in) and qdd (out) ever leaves the SM:
//
//   chain phase, 16 lanes per state (lane r <-> joints 2r, 2r + 1), two states at a time:
//   1. every joint's local transform X_i(q_i) = (R_p Rz(q_i), t) (joint.rs:36-38) is built by its lane, and a prefix
//      product over the lanes gives every joint's world pose (R_i, p_i)
//   2. with everything expressed in the world frame about the world origin, the recursions of rnea
//      (multibody.rs:122-150) and of crba's composite inertias (multibody.rs:157-171) become plain sums along the
//      chain: prefix sums for omega, alpha and the linear acceleration, suffix sums for the link wrenches and for
//      the 10-parameter inertias (m, m c, I about the origin).  bias_i = s_i . f^c_i with the joint screw
//      s_i = (z_i ; p_i x z_i)
//   3. the screws s_j, the vectors I^c_i s_i and tau - bias go to shared memory (the hand-over)
//   matrix phase, 8 lanes per state (lane (s, r) <-> rows r, r + 8, r + 16, r + 24 of state s), four states at once:
//   4. H[i][j] = s_j . (I^c_i s_i) for j <= i: the lane keeps I^c_i s_i of its four rows in registers, the screws are
//      8-byte broadcasts from shared memory, and the lower triangle of the symmetric matrix is born in registers
//      (8 + 16 + 24 + 32 entries per lane) -- the layout the elimination wants, so H is never written anywhere
//   5. right-looking LDL^T with rows in registers: at step k every lane writes its entries of column k to shared
//      memory, the row's owner adds a 16-byte header {d_k, y_k}, the column comes back as broadcasts and every lane
//      updates its rows; the rhs rides along; the back substitution walks the stored columns of L.
//
// Same result as  qdd = solve(sym(crba(q)), tau - rnea(q, dq, 0))   up to rounding: the world-frame
// sums associate differently from the reference's link-frame recursion (measured 5e-14 relative on the 32-joint
// chain, cond(H) ~ 1e4).
//
// History (numbers: 32-joint chain, 2^20 states on one B200; profiles/r1_kbench_chain32_warp_fd.jsonl,
// profiles/r2_kbench_chain32.jsonl): a warp per state (lane = joint = row) 0.148 G evals/s; half a warp per state in both
// phases 0.224; eight lanes per state in the matrix phase 0.268; warp-uniform loop bounds, mask-FMA scans, 8-byte
// broadcasts, asynchronous staging, no fifth scan shuffle 0.308; team barriers, odd offset between the states' column
// areas 0.322 (17.3 TFLOP/s by SURVEY 8d flops = 0.51 of the measured FP64 peak).  The two older kernels live in
// experiments/rb_warp_fd_old.cuh (build with -DRBW_OLD_KERNELS=1; $RIGIDBODY_B200_WARP_FD=half selects the second).
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "rb_kernels.cuh"
#include "rb_util.cuh"

#ifndef RBW_GROUP
#define RBW_GROUP 4                       // states a warp stages at once (32-byte runs of the joint-major arrays)
#endif
#ifndef RBW_OLD_KERNELS
#define RBW_OLD_KERNELS 0                 // 1 = also build the warp-per-state and half-warp-per-state kernels (experiments/)
#endif
#define RBW_IOS (RBW_GROUP + 1)           // padded stride of the staging rows

namespace {
constexpr unsigned FULL = 0xffffffffu;
constexpr int RBH_MODEL = 24 * 32;        // block-shared model constants [e][32], [22][*] = composite mass

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

__device__ __forceinline__ double up16(double v, int d) { return __shfl_up_sync(FULL, v, d, 16); }
__device__ __forceinline__ double dn16(double v, int d) { return __shfl_down_sync(FULL, v, d, 16); }

// Scans along the chain without selects: a lane whose source lane lies outside its 16-lane segment multiplies what
// the shuffle returned (its own value) by 0.0, every other lane by 1.0 -- the add of a scan step becomes an FMA with a
// per-lane mask, one FP64 instruction as before, and the compare + two FSELs per step disappear (they were 10 % of the
// instructions the kernel executed; ptxas turns a predicated add back into selects).
struct ScanMasks { double up[4], dn[4]; };
__device__ __forceinline__ ScanMasks scan_masks(int r) {
    ScanMasks m;
#pragma unroll
    for (int i = 0; i < 4; ++i) { m.up[i] = r >= (1 << i) ? 1.0 : 0.0; m.dn[i] = r + (1 << i) < 16 ? 1.0 : 0.0; }
    return m;
}
template <int K>
__device__ __forceinline__ void prefix2p(double (&x0)[K], double (&x1)[K], const ScanMasks& m) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const double own = x0[c] + x1[c];
        double t = own;
#pragma unroll
        for (int i = 0; i < 4; ++i) t = fma(up16(t, 1 << i), m.up[i], t);
        x0[c] += t - own;                    // the lanes before this one: inclusive total minus the own pair (no fifth shuffle;
                                             // the difference is off by at most half an ulp of the total)
        x1[c] += x0[c];
    }
}
template <int K>
__device__ __forceinline__ void suffix2p(double (&x0)[K], double (&x1)[K], const ScanMasks& m) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const double own = x0[c] + x1[c];
        double t = own;
#pragma unroll
        for (int i = 0; i < 4; ++i) t = fma(dn16(t, 1 << i), m.dn[i], t);
        x1[c] += t - own;
        x0[c] += x1[c];
    }
}
// (A, u) o (R, p) = (A R, u + A p)
__device__ __forceinline__ void compose(const double (&A)[9], const double (&u)[3], const double (&R)[9], const double (&p)[3],
                                        double (&oR)[9], double (&op)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        op[r] = fma(A[3 * r], p[0], fma(A[3 * r + 1], p[1], fma(A[3 * r + 2], p[2], u[r])));
#pragma unroll
        for (int c = 0; c < 3; ++c) oR[3 * r + c] = fma(A[3 * r], R[c], fma(A[3 * r + 1], R[3 + c], A[3 * r + 2] * R[6 + c]));
    }
}

// Per-joint part of the world-frame pass: link wrench about the origin (fw) and the 9 inertia parameters (ci).
__device__ __forceinline__ void link_terms(const double (&R)[9], const double (&p)[3], double m, const double (&h)[3], const double (&Io)[6],
                                           const double (&om)[3], const double (&al)[3], const double (&acc)[3],
                                           double (&fw)[6], double (&ci)[9]) {
    double hw[3], Iw[6], T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        hw[r] = a * h[0] + b * h[1] + c * h[2];
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
// I^c s for the screw s = (z ; v): force Ff and moment Fn about the origin.
__device__ __forceinline__ void comp_times_screw(double mc, const double (&ci)[9], const double (&z)[3], const double (&v)[3],
                                                 double (&Fn)[3], double (&Ff)[3]) {
    const double Hc[3] = {ci[0], ci[1], ci[2]};
    const double IO[6] = {ci[3], ci[4], ci[5], ci[6], ci[7], ci[8]};
    Ff[0] = mc * v[0]; Ff[1] = mc * v[1]; Ff[2] = mc * v[2];
    cross_acc(z, Hc, Ff);
    sym_mul(IO, z, Fn);
    cross_acc(Hc, v, Fn);
}

#if RBW_OLD_KERNELS
#include "experiments/rb_warp_fd_old.cuh"
#endif

// ------------------------------------------------------------------ a quarter of a warp per state in the matrix phase
// rbq_fd_kernel: the chain phase runs on 16 lanes per state (twice for the four staged states), the mass matrix is built
// and factorised by EIGHT lanes per state, four states per warp: lane (s, r) owns rows r, r + 8,
// r + 16, r + 24 of state s, lower triangle padded to the row group's width (8 + 16 + 24 + 32 = 80 register entries).
// Against 16 lanes per state every broadcast load, every stored pivot column, every reciprocal and every shuffle of the
// elimination now serves four states instead of two, and the rows of a lane are closer to the triangle (the FP64 work
// per state drops by a quarter: 230 instead of 308 warp-level FMAs in the elimination).  The right-hand side and the
// pivot travel together in one 16-byte header per column (no shuffle in the elimination).
//   shared memory per state: columns of L D (column k holds rows 8 (k / 8) .. 31), then the 32 headers {d_k, y_k};
//   the header's d_k is overwritten by 1 / d_k one step later (read by the back substitution).
#ifndef RBQ_WARPS
#define RBQ_WARPS 8
#endif
constexpr int rbq_colofs(int k) {
    return k < 8 ? 34 * k : k < 16 ? 272 + 26 * (k - 8) : k < 24 ? 480 + 18 * (k - 16) : 624 + 8 * (k - 24);
}
constexpr int RBQ_COLS = 688;                         // = rbq_colofs(32)
// The four states' column areas are 3 doubles (mod 16) apart: an 8-byte broadcast (one address per state) touches four
// different bank pairs = one wavefront, and the back substitution's column reads (lane r of a state reads every 34th /
// 26th / 18th double: even bank pairs) interleave with the neighbouring state's (odd bank pairs).  Measured with the
// team barriers in place (deterministic timing, profiles/r2_kbench_chain32.jsonl): odd offsets 3, 5, 7, 9 -> 3.26 ms,
// even offsets 2, 4, 6 -> 3.36 ms, offset 1 -> 3.62 ms, offsets 0, 8, 2, 10 (conflict-free column stores) -> 3.59 ms.
// Nothing in the column area is accessed 16 bytes at a time, so the odd offset costs no alignment.
constexpr int RBQ_SS = RBQ_COLS + 64;                 // columns + headers of one state
constexpr int RBQ_HS = 32 * 6 + 7 * 32;               // hand-over of one state: screws + component-major I^c s and rhs
#ifndef RBQ_SHIFT
#define RBQ_SHIFT 3
#endif
__device__ __forceinline__ constexpr int rbq_bank_shift(int st) { return RBQ_SHIFT * st; }
constexpr int RBQ_REGION = 4 * RBQ_SS + 32;
constexpr int RBQ_IO = 3 * 32 * RBW_IOS;              // staged q, dq, tau
constexpr int RBQ_PER_WARP = RBQ_REGION + RBQ_IO;
static_assert(RBQ_SS % 16 == 0 && RBQ_HS % 16 == 0 && RBQ_HS <= RBQ_SS, "hand-over buffers alias the column storage");
static_assert(RBW_GROUP == 4, "rbq_fd_kernel factorises the four staged states together");

// A broadcast read of shared memory that stays an 8-byte load: nvcc pairs adjacent 8-byte loads whose alignment it can
// prove into one 16-byte load, and a 16-byte load is served a quarter of a warp at a time -- four wavefronts for the four
// states' addresses where two 8-byte loads take one each (measured: 3.8 wavefronts per LDS.128 in the elimination).
#ifndef RBQ_LDS64
#define RBQ_LDS64 1
#endif
__device__ __forceinline__ void rbq_sts(double* p, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v) : "memory");
}
__device__ __forceinline__ double rbq_lds(const double* p) {
#if RBQ_LDS64
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
#else
    return *p;
#endif
}

// Chain phase of one state on 16 lanes (lane r <-> joints 2r, 2r + 1): world poses by a prefix product, the recursions of
// rnea / crba as sums along the chain (see the file header), then the hand-over: screws to Sb[32][6], I^c s and
// tau - bias to Fb[7][32] (component-major).
__device__ __forceinline__ void rbq_chain_phase(const double* __restrict__ msm, const double* __restrict__ io, int st, int r, int n,
                                                const double (&g)[3], double* __restrict__ Sb, double* __restrict__ Fb) {
    const int j0 = 2 * r;
    const bool act0 = j0 < n, act1 = j0 + 1 < n;
    // model constants [e][parity][16]: the lane's two joints are two 8-byte loads of 16 consecutive doubles each
    auto mdl2 = [&](int e) { return make_double2(msm[e * 32 + r], msm[e * 32 + 16 + r]); };
    double R0[9], p0[3], R1[9], p1[3];
    {
        double sn, cs;
        sincos(io[(0 * 32 + j0) * RBW_IOS + st], &sn, &cs);
        double sn1, cs1;
        sincos(io[(0 * 32 + j0 + 1) * RBW_IOS + st], &sn1, &cs1);
        double T0[9], T1[9], t0[3], t1[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const double2 a = mdl2(3 * rr), b = mdl2(3 * rr + 1), c = mdl2(3 * rr + 2), t = mdl2(9 + rr);
            T0[3 * rr + 0] = fma(cs, a.x, sn * b.x);   T1[3 * rr + 0] = fma(cs1, a.y, sn1 * b.y);
            T0[3 * rr + 1] = fma(cs, b.x, -sn * a.x);  T1[3 * rr + 1] = fma(cs1, b.y, -sn1 * a.y);
            T0[3 * rr + 2] = c.x;                      T1[3 * rr + 2] = c.y;
            t0[rr] = t.x;                              t1[rr] = t.y;
        }
        double C[9], cp[3];
        compose(T0, t0, T1, t1, C, cp);
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            double A[9], u[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) A[e] = up16(C[e], d);
#pragma unroll
            for (int e = 0; e < 3; ++e) u[e] = up16(cp[e], d);
            if (r >= d) {
                double nC[9], np[3];
                compose(A, u, C, cp, nC, np);
#pragma unroll
                for (int e = 0; e < 9; ++e) C[e] = nC[e];
#pragma unroll
                for (int e = 0; e < 3; ++e) cp[e] = np[e];
            }
        }
        double E[9], ep[3];
#pragma unroll
        for (int e = 0; e < 9; ++e) { const double x = up16(C[e], 1); E[e] = r ? x : ((e == 0 || e == 4 || e == 8) ? 1.0 : 0.0); }
#pragma unroll
        for (int e = 0; e < 3; ++e) { const double x = up16(cp[e], 1); ep[e] = r ? x : 0.0; }
        compose(E, ep, T0, t0, R0, p0);
        compose(R0, p0, T1, t1, R1, p1);
    }
    const double dq0 = io[(1 * 32 + j0) * RBW_IOS + st], dq1 = io[(1 * 32 + j0 + 1) * RBW_IOS + st];
    double z0[3] = {act0 ? R0[2] : 0.0, act0 ? R0[5] : 0.0, act0 ? R0[8] : 0.0};
    double z1[3] = {act1 ? R1[2] : 0.0, act1 ? R1[5] : 0.0, act1 ? R1[8] : 0.0};
    double v0[3], v1[3];
    cross(p0, z0, v0);
    cross(p1, z1, v1);
    const double zq0[3] = {z0[0] * dq0, z0[1] * dq0, z0[2] * dq0}, zq1[3] = {z1[0] * dq1, z1[1] * dq1, z1[2] * dq1};
    double om0[3] = {zq0[0], zq0[1], zq0[2]}, om1[3] = {zq1[0], zq1[1], zq1[2]};
    const ScanMasks sm = scan_masks(r);
    
    prefix2p<3>(om0, om1, sm);

    double al0[3], al1[3];
    cross(om0, zq0, al0);
    cross(om1, zq1, al1);
    
    prefix2p<3>(al0, al1, sm);

    double ac0[3], ac1[3];
    {
        double omp[3], alp[3], d[3], w1[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {                // joint j0's predecessor is the previous lane's second joint
            const double a = up16(om1[e], 1), b = up16(al1[e], 1), c = up16(p1[e], 1);
            omp[e] = r ? a : 0.0;
            alp[e] = r ? b : 0.0;
            d[e] = p0[e] - (r ? c : 0.0);
        }
        cross(omp, d, w1);
        cross(alp, d, ac0);
        cross_acc(omp, w1, ac0);
#pragma unroll
        for (int e = 0; e < 3; ++e) ac0[e] += r ? 0.0 : g[e];
        const double d1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        cross(om0, d1, w1);
        cross(al0, d1, ac1);
        cross_acc(om0, w1, ac1);
    }
    
    prefix2p<3>(ac0, ac1, sm);

    double fw0[6], fw1[6], ci0[9], ci1[9];
    {
        const double2 m = mdl2(12), h0 = mdl2(13), h1 = mdl2(14), h2 = mdl2(15);
        const double2 i0 = mdl2(16), i1 = mdl2(17), i2 = mdl2(18), i3 = mdl2(19), i4 = mdl2(20), i5 = mdl2(21);
        const double ha[3] = {h0.x, h1.x, h2.x}, hb[3] = {h0.y, h1.y, h2.y};
        const double Ia[6] = {i0.x, i1.x, i2.x, i3.x, i4.x, i5.x}, Ib[6] = {i0.y, i1.y, i2.y, i3.y, i4.y, i5.y};
        link_terms(R0, p0, m.x, ha, Ia, om0, al0, ac0, fw0, ci0);
        link_terms(R1, p1, m.y, hb, Ib, om1, al1, ac1, fw1, ci1);
    }
    
    suffix2p<6>(fw0, fw1, sm);

    const double b0 = io[(2 * 32 + j0) * RBW_IOS + st]
                      - (z0[0] * fw0[3] + z0[1] * fw0[4] + z0[2] * fw0[5] + v0[0] * fw0[0] + v0[1] * fw0[1] + v0[2] * fw0[2]);
    const double b1 = io[(2 * 32 + j0 + 1) * RBW_IOS + st]
                      - (z1[0] * fw1[3] + z1[1] * fw1[4] + z1[2] * fw1[5] + v1[0] * fw1[0] + v1[1] * fw1[1] + v1[2] * fw1[2]);
    
    suffix2p<9>(ci0, ci1, sm);

    const double2 mc = mdl2(22);
    double Fn0[3], Ff0[3], Fn1[3], Ff1[3];
    double2* S2 = reinterpret_cast<double2*>(Sb + j0 * 6);
    comp_times_screw(mc.x, ci0, z0, v0, Fn0, Ff0);
    comp_times_screw(mc.y, ci1, z1, v1, Fn1, Ff1);
    S2[0] = make_double2(z0[0], z0[1]); S2[1] = make_double2(z0[2], v0[0]); S2[2] = make_double2(v0[1], v0[2]);
    S2[3] = make_double2(z1[0], z1[1]); S2[4] = make_double2(z1[2], v1[0]); S2[5] = make_double2(v1[1], v1[2]);
    double2* F2 = reinterpret_cast<double2*>(Fb + j0);
    F2[0 * 16] = make_double2(Fn0[0], Fn1[0]); F2[1 * 16] = make_double2(Fn0[1], Fn1[1]); F2[2 * 16] = make_double2(Fn0[2], Fn1[2]);
    F2[3 * 16] = make_double2(Ff0[0], Ff1[0]); F2[4 * 16] = make_double2(Ff0[1], Ff1[1]); F2[5 * 16] = make_double2(Ff0[2], Ff1[2]);
    F2[6 * 16] = make_double2(act0 ? b0 : 0.0, act1 ? b1 : 0.0);
}

#ifndef RBQ_TEAMS
#define RBQ_TEAMS 2
#endif
// compile-time loops with the index as a constant expression (register arrays of different lengths per row group)
template <int I, int E, class F>
__device__ __forceinline__ void rbq_for(F&& f) {
    if constexpr (I < E) { f(std::integral_constant<int, I>{}); rbq_for<I + 1, E>(f); }
}
template <int I, int E, class F>
__device__ __forceinline__ void rbq_for_down(F&& f) {          // I, I - 1, ..., E
    if constexpr (I >= E) { f(std::integral_constant<int, I>{}); rbq_for_down<I - 1, E>(f); }
}

__global__ void __launch_bounds__(32 * RBQ_WARPS, 1)
rbq_fd_kernel(const double* __restrict__ model, int n, const double* __restrict__ q, const double* __restrict__ dq,
              const double* __restrict__ tau, double* __restrict__ qdd, size_t B, size_t ld, int* __restrict__ status) {
    extern __shared__ __align__(16) double rbw_sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* msm = rbw_sm;                                    // [24][32] model constants, [22][*] = composite mass
    double* wsm = rbw_sm + RBH_MODEL + (size_t)w * RBQ_PER_WARP;
    double* io = wsm + RBQ_REGION;                           // [3][32][RBW_IOS]
    double* ob = wsm;                                        // results [32][RBW_IOS]: over the columns of L, dead after the back substitution
    if (w == 0) {                                            // model -> shared memory; idle joints: identity, no mass
        const bool act = lane < n;
        const double* row = model + (size_t)(act ? lane : 0) * 24;
        double mc[1] = {act ? row[12] : 0.0};
        suffix_sum<1>(mc, lane);
        const int slot = (lane & 1) * 16 + (lane >> 1);      // [parity][joint / 2]
#pragma unroll
        for (int e = 0; e < 13; ++e) msm[e * 32 + slot] = act ? row[e] : ((e == 0 || e == 4 || e == 8) ? 1.0 : 0.0);
#pragma unroll
        for (int e = 0; e < 9; ++e) msm[(13 + e) * 32 + slot] = act ? row[14 + e] : 0.0;
        msm[22 * 32 + slot] = mc[0];
    }
    __syncthreads();
    const double g[3] = {model[(size_t)n * 24], model[(size_t)n * 24 + 1], model[(size_t)n * 24 + 2]};
    // chain phase: half h of the warp, lane rh of 16;  matrix phase: state s of the group, lane r of 8
    const int h = lane >> 4, rh = lane & 15;
    const int s = lane >> 3, r = lane & 7;
    double* Ls = wsm + s * RBQ_SS + rbq_bank_shift(s);       // this state's columns of L D
    // Ls[RBQ_COLS + 2 k], Ls[RBQ_COLS + 2 k + 1]: header of column k = {d_k (later 1 / d_k), y_k}
    const double* Sq = wsm + s * RBQ_HS + 2 * s; // hand-over of state s: screws [32][6] ...
    const double* Fq = Sq + 32 * 6;                          // ... and [7][32] I^c s (6 components) and tau - bias
    int colbase[4];                                          // column r + 8 g of the lane's rows: entry of row i at colbase[g] + i
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
        const int k = r + 8 * gg;
        colbase[gg] = (gg == 0 ? 34 * k : gg == 1 ? 272 + 26 * (k - 8) : gg == 2 ? 480 + 18 * (k - 16) : 624 + 8 * (k - 24)) - 8 * gg;
    }

    // The trip count is the same for every warp of the grid (a warp whose last group does not exist repeats the grid's last
    // group and stores nothing): control flow that depends only on kernel parameters lets the compiler prove the warp
    // converged at every shuffle -- with a per-warp loop bound it re-materialised the member mask and tested for
    // divergence before each of the ~450 shuffles of an iteration (6 % of the instructions executed).
    const size_t groups = (B + RBW_GROUP - 1) / RBW_GROUP;
    const size_t per_sweep = (size_t)gridDim.x * RBQ_WARPS;
    const size_t sweeps = (groups + per_sweep - 1) / per_sweep;
    bool all_ok = true;
    // Staging: a group's q, dq, tau travel global -> shared as 8-byte asynchronous copies (zero-filled where the state or
    // the joint does not exist), issued for the NEXT group as soon as the chain phases have consumed the current one, so
    // the HBM latency hides behind the matrix phase (with plain loads at the top of the iteration it was 9.5 % of the
    // kernel's stall samples).
    auto stage = [&](size_t grp_) {
        const size_t s0_ = grp_ * RBW_GROUP;
        const int sj = lane / RBW_GROUP, ss = lane % RBW_GROUP;
        const bool sin = s0_ + ss < B;
#pragma unroll
        for (int it = 0; it < 32 / (32 / RBW_GROUP); ++it) {
            const int i = it * (32 / RBW_GROUP) + sj;
            const bool ld_ok = sin && i < n;
            const size_t off = ld_ok ? (size_t)i * ld + s0_ + ss : 0;
            const unsigned bytes = ld_ok ? 8u : 0u;
            const unsigned d0 = (unsigned)__cvta_generic_to_shared(io + (0 * 32 + i) * RBW_IOS + ss);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0), "l"(q + off), "r"(bytes) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 32 * RBW_IOS * 8), "l"(dq + off), "r"(bytes) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 2 * 32 * RBW_IOS * 8), "l"(tau + off), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto group_of = [&](size_t sweep_) {
        const size_t want = sweep_ * per_sweep + (size_t)blockIdx.x * RBQ_WARPS + w;
        return want < groups ? want : groups - 1;
    };
    stage(group_of(0));
    for (size_t sweep = 0; sweep < sweeps; ++sweep) {
        const bool valid = sweep * per_sweep + (size_t)blockIdx.x * RBQ_WARPS + w < groups;
        const size_t grp = group_of(sweep);
        const size_t s0 = grp * RBW_GROUP;
        asm volatile("cp.async.wait_all;" ::: "memory");
#if RBQ_TEAMS
        // teams of RBQ_WARPS / RBQ_TEAMS warps meet at a named barrier at the top of every sweep: the kernel is ~96 KB of code
        // against a 32 KB L1.5 instruction cache, and a team streams ONE copy of it.  Measured (2^20 states): no barrier 3.41 ms
        // with 2 % spread, one team of 8 3.52, two teams of 4 3.36 with 0.2 % spread, four teams of 2 3.42; the second team
        // started half an iteration late: 3.62 -- sharing the instruction stream beats mixing phases on the pipes
        asm volatile("bar.sync %0, %1;" ::"r"(1 + w / (RBQ_WARPS / RBQ_TEAMS)), "n"(32 * (RBQ_WARPS / RBQ_TEAMS)) : "memory");
#else
        __syncwarp();
#endif
        const int in_group = (int)(B - s0 < RBW_GROUP ? B - s0 : RBW_GROUP);
        // ================= chain phase, two states at a time (a state that does not exist is computed from zeros: q = dq =
        // tau = 0 were staged for it, its matrix is that of the zero pose and nothing of it is stored)
#pragma unroll 1
        for (int st2 = 0; st2 < 4; st2 += 2) {
            const int st = st2 + h;
            double* ho = wsm + st * RBQ_HS + 2 * st;
            rbq_chain_phase(msm, io, st, rh, n, g, ho, ho + 32 * 6);
        }
        __syncwarp();
        if (sweep + 1 < sweeps) stage(group_of(sweep + 1));  // the staging buffer is free: fetch the next group
        // ================= matrix phase: lane (s, r) <-> rows r, r + 8, r + 16, r + 24 of state s
        double a0[8], a1[16], a2[24], a3[32], b[4];
        {
            double F0[6], F1[6], F2[6], F3[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                F0[c] = Fq[c * 32 + r]; F1[c] = Fq[c * 32 + r + 8]; F2[c] = Fq[c * 32 + r + 16]; F3[c] = Fq[c * 32 + r + 24];
            }
            b[0] = Fq[6 * 32 + r]; b[1] = Fq[6 * 32 + r + 8]; b[2] = Fq[6 * 32 + r + 16]; b[3] = Fq[6 * 32 + r + 24];
            auto dot6 = [](const double2& s01, const double2& s23, const double2& s45, const double (&F)[6]) {
                return fma(s01.x, F[0], fma(s01.y, F[1], fma(s23.x, F[2], fma(s23.y, F[3], fma(s45.x, F[4], s45.y * F[5])))));
            };
            rbq_for<0, 32>([&](auto jc) {
                constexpr int J = decltype(jc)::value;
                // 8-byte broadcasts: the four states' addresses are one wavefront (a 16-byte load is served a quarter
                // of a warp at a time: four wavefronts for the same four addresses, measured)
                const double* Sj = Sq + J * 6;
                const double2 s01 = make_double2(rbq_lds(Sj), rbq_lds(Sj + 1)), s23 = make_double2(rbq_lds(Sj + 2), rbq_lds(Sj + 3)),
                              s45 = make_double2(rbq_lds(Sj + 4), rbq_lds(Sj + 5));
                if constexpr (J < 8) a0[J] = dot6(s01, s23, s45, F0);
                if constexpr (J < 16) a1[J] = dot6(s01, s23, s45, F1);
                if constexpr (J < 24) a2[J] = dot6(s01, s23, s45, F2);
                a3[J] = dot6(s01, s23, s45, F3);
            });
            if (n < 32) {                                    // idle joints: identity rows
                rbq_for<0, 32>([&](auto jc) {
                    constexpr int J = decltype(jc)::value;
                    if constexpr (J < 8) a0[J] = (r >= n && J == r) ? 1.0 : a0[J];
                    if constexpr (J < 16) a1[J] = (r + 8 >= n && J == r + 8) ? 1.0 : a1[J];
                    if constexpr (J < 24) a2[J] = (r + 16 >= n && J == r + 16) ? 1.0 : a2[J];
                    a3[J] = (r + 24 >= n && J == r + 24) ? 1.0 : a3[J];
                });
            }
        }
        __syncwarp();                                        // the hand-over buffers are free: the columns of L go there
        bool ok = true;
        double dinv_prev = 0.0;
        rbq_for<0, 32>([&](auto kc) {
            constexpr int K = decltype(kc)::value, GK = K >> 3, RK = K & 7;
            constexpr int COL = rbq_colofs(K) - 8 * GK;      // entry of row i at COL + i
            // column K: the entries of the lane's rows in groups GK..3 (finished rows of group GK store into dead slots)
            if constexpr (GK <= 0) Ls[COL + r] = a0[K & 7];
            if constexpr (GK <= 1) Ls[COL + r + 8] = a1[K & 15];
            if constexpr (GK <= 2) Ls[COL + r + 16] = a2[K < 24 ? K : 0];
            Ls[COL + r + 24] = a3[K];
            if (r == RK) {
                const double dk = GK == 0 ? a0[K & 7] : GK == 1 ? a1[K & 15] : GK == 2 ? a2[K < 24 ? K : 0] : a3[K];
                rbq_sts(Ls + RBQ_COLS + 2 * K, dk);          // two 8-byte stores: one wavefront each (a 16-byte store by four lanes: four)
                rbq_sts(Ls + RBQ_COLS + 2 * K + 1, b[GK]);
            }
            __syncwarp();
            if constexpr (K > 0) { if (r == ((K - 1) & 7)) Ls[RBQ_COLS + 2 * (K - 1)] = dinv_prev; }
            const double2 hd = make_double2(rbq_lds(Ls + RBQ_COLS + 2 * K), rbq_lds(Ls + RBQ_COLS + 2 * K + 1));
            ok = ok && (hd.x > 0.0);
            const double dinv = rb_rcp_pos(hd.x);
            dinv_prev = dinv;
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3;
            if constexpr (GK <= 0) n0 = -a0[K & 7] * dinv;
            if constexpr (GK <= 1) n1 = -a1[K & 15] * dinv;
            if constexpr (GK <= 2) n2 = -a2[K < 24 ? K : 0] * dinv;
            n3 = -a3[K] * dinv;
            // right-hand side: rows after K only (a finished row's multiplier is meaningless)
            if constexpr (GK == 0) { if (r > RK) b[0] = fma(n0, hd.y, b[0]); b[1] = fma(n1, hd.y, b[1]); b[2] = fma(n2, hd.y, b[2]); b[3] = fma(n3, hd.y, b[3]); }
            if constexpr (GK == 1) { if (r > RK) b[1] = fma(n1, hd.y, b[1]); b[2] = fma(n2, hd.y, b[2]); b[3] = fma(n3, hd.y, b[3]); }
            if constexpr (GK == 2) { if (r > RK) b[2] = fma(n2, hd.y, b[2]); b[3] = fma(n3, hd.y, b[3]); }
            if constexpr (GK == 3) { if (r > RK) b[3] = fma(n3, hd.y, b[3]); }
            auto upd = [&](auto jc, double c) {
                constexpr int J = decltype(jc)::value;
                if constexpr (J < 8 && GK <= 0) a0[J & 7] = fma(n0, c, a0[J & 7]);
                if constexpr (J < 16 && GK <= 1) a1[J & 15] = fma(n1, c, a1[J & 15]);
                if constexpr (J < 24 && GK <= 2) a2[J < 24 ? J : 0] = fma(n2, c, a2[J < 24 ? J : 0]);
                a3[J] = fma(n3, c, a3[J]);
            };
            if constexpr (((K + 1) & 1) && K + 1 < 32) upd(std::integral_constant<int, (K + 1) & 31>{}, rbq_lds(Ls + COL + K + 1));
            rbq_for<((K + 2) & ~1) / 2, 16>([&](auto pc) {
                constexpr int J = 2 * decltype(pc)::value;
                upd(std::integral_constant<int, J>{}, rbq_lds(Ls + COL + J));
                upd(std::integral_constant<int, J + 1>{}, rbq_lds(Ls + COL + J + 1));
            });
        });
        __syncwarp();
        if (r == 7) Ls[RBQ_COLS + 2 * 31] = dinv_prev;
        __syncwarp();
        // back substitution: x_i = (y_i - sum_{j > i} (l_ji d_i) x_j) / d_i; the lane reads its own four columns
        double x0 = b[0], x1 = b[1], x2 = b[2], x3 = b[3];
        const double di0 = Ls[RBQ_COLS + 2 * r], di1 = Ls[RBQ_COLS + 2 * (r + 8)], di2 = Ls[RBQ_COLS + 2 * (r + 16)], di3 = Ls[RBQ_COLS + 2 * (r + 24)];
        rbq_for_down<31, 1>([&](auto ic) {
            constexpr int I = decltype(ic)::value, GI = I >> 3, RI = I & 7;
            const double mine = GI == 0 ? x0 * di0 : GI == 1 ? x1 * di1 : GI == 2 ? x2 * di2 : x3 * di3;
            const double xi = __shfl_sync(FULL, mine, RI, 8);
            if (r == RI) { if constexpr (GI == 0) x0 = xi; if constexpr (GI == 1) x1 = xi; if constexpr (GI == 2) x2 = xi; if constexpr (GI == 3) x3 = xi; }
            // rows before I: every row of the groups below GI, rows r < RI of group GI
            if constexpr (GI > 0) x0 = fma(-Ls[colbase[0] + I], xi, x0);
            if constexpr (GI > 1) x1 = fma(-Ls[colbase[1] + I], xi, x1);
            if constexpr (GI > 2) x2 = fma(-Ls[colbase[2] + I], xi, x2);
            if (r < RI) {
                const double c = Ls[colbase[GI] + I];
                if constexpr (GI == 0) x0 = fma(-c, xi, x0);
                if constexpr (GI == 1) x1 = fma(-c, xi, x1);
                if constexpr (GI == 2) x2 = fma(-c, xi, x2);
                if constexpr (GI == 3) x3 = fma(-c, xi, x3);
            }
        });
        if (r == 0) x0 *= di0;
        const bool live = valid && s < in_group;
        all_ok = all_ok && (ok || !live);
        __syncwarp();                                        // every lane is done with the stored columns (ob lies over them)
        if (live) {
            ob[r * RBW_IOS + s] = ok ? x0 : rb_nan<double>();
            ob[(r + 8) * RBW_IOS + s] = ok ? x1 : rb_nan<double>();
            ob[(r + 16) * RBW_IOS + s] = ok ? x2 : rb_nan<double>();
            ob[(r + 24) * RBW_IOS + s] = ok ? x3 : rb_nan<double>();
        }
        __syncwarp();
        {
            const int sj = lane / RBW_GROUP, ss = lane % RBW_GROUP;
            if (valid && s0 + ss < B) {
#pragma unroll
                for (int it = 0; it < 32 / (32 / RBW_GROUP); ++it) {
                    const int i = it * (32 / RBW_GROUP) + sj;
                    if (i < n) __stcs(qdd + (size_t)i * ld + s0 + ss, ob[i * RBW_IOS + ss]);
                }
            }
        }
        __syncwarp();
    }
    if (!all_ok && r == 0) atomicOr(status, RB_STATUS_NOT_SPD);
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
    const size_t groups = (B + RBW_GROUP - 1) / RBW_GROUP;
#if RBW_OLD_KERNELS
    // $RIGIDBODY_B200_WARP_FD=half / warp selects one of the two earlier kernels (comparisons only)
    static const char old_kernel = [] { const char* e = getenv("RIGIDBODY_B200_WARP_FD"); return e ? e[0] : 'q'; }();
    if (old_kernel == 'h') {
        constexpr size_t hsmem = ((size_t)RBW_HWARPS * RBH_PER_WARP + RBH_MODEL) * sizeof(double);
        static std::atomic<bool> hconfigured[64];
        if (dev >= 0 && dev < 64 && !hconfigured[dev].load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(rbh_fd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem);
            if (e != cudaSuccess) return e;
            hconfigured[dev].store(true, std::memory_order_release);
        }
        const size_t hwant = (groups + RBW_HWARPS - 1) / RBW_HWARPS;
        rbh_fd_kernel<<<(unsigned)(hwant < (size_t)sms ? hwant : (size_t)sms), 32 * RBW_HWARPS, hsmem, st>>>(model, n, q, dq, tau, qdd, B, ld, status);
        return cudaGetLastError();
    }
    if (old_kernel == 'w') {
        constexpr size_t smem = ((size_t)RBW_WARPS * RBW_PER_WARP + RBW_MODEL_DOUBLES) * sizeof(double);
        static std::atomic<bool> configured[64];
        if (dev >= 0 && dev < 64 && !configured[dev].load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(rbw_fd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev].store(true, std::memory_order_release);
        }
        const size_t want = (groups + RBW_WARPS - 1) / RBW_WARPS;
        rbw_fd_kernel<<<(unsigned)(want < (size_t)sms ? want : (size_t)sms), 32 * RBW_WARPS, smem, st>>>(model, n, q, dq, tau, qdd, B, ld, status);
        return cudaGetLastError();
    }
#endif
    constexpr size_t qsmem = ((size_t)RBQ_WARPS * RBQ_PER_WARP + RBH_MODEL) * sizeof(double);
    static_assert(qsmem <= 232448, "rbq_fd_kernel: more than 227 KB of shared memory");
    static std::atomic<bool> qconfigured[64];                // per device; setting the attribute twice is harmless
    if (dev >= 0 && dev < 64 && !qconfigured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(rbq_fd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem);
        if (e != cudaSuccess) return e;
        qconfigured[dev].store(true, std::memory_order_release);
    }
    const size_t qwant = (groups + RBQ_WARPS - 1) / RBQ_WARPS;
    rbq_fd_kernel<<<(unsigned)(qwant < (size_t)sms ? qwant : (size_t)sms), 32 * RBQ_WARPS, qsmem, st>>>(model, n, q, dq, tau, qdd, B, ld, status);
    return cudaGetLastError();
}
