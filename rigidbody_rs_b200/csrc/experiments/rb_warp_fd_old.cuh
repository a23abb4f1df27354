// experiments/rb_warp_fd_old.cuh -- the two earlier long-chain forward-dynamics kernels, kept for comparisons (included
// by rb_kernels_warp.cu inside its anonymous namespace when built with -DRBW_OLD_KERNELS=1):
//   rbw_fd_kernel: one warp per state, lane = joint = row                                   (0.148 G evals/s, 32 joints)
//   rbh_fd_kernel: half a warp per state in both phases, rows r and r + 16 per lane         (0.224 G evals/s)
// The product kernel is rbq_fd_kernel (eight lanes per state in the matrix phase, 0.308 G evals/s).
#ifndef RBW_WARPS
#define RBW_WARPS 12
#endif
#ifndef RBW_MODEL_SMEM
#define RBW_MODEL_SMEM 0                  // 1 = per-lane model constants re-read from shared memory (frees ~46 registers)
#endif
#ifndef RBW_HALF
#define RBW_HALF 1                        // 1 = rbh_fd_kernel, 0 = rbw_fd_kernel when the old kernels are selected
#endif
#define RBW_LDL 34                        // row stride of the stored L columns: even, so pairs are 16-byte aligned
constexpr int RBW_PER_WARP = 32 * RBW_LDL + 32 * 6 + 4 * 32 * RBW_IOS;   // doubles of shared memory per warp
constexpr int RBW_MODEL_DOUBLES = RBW_MODEL_SMEM ? 23 * 32 : 0;          // block-shared copy of the per-lane constants

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

// ------------------------------------------------------------------ half a warp per state
// Same algorithm, two states per warp: lanes 0-15 own one state, lanes 16-31 another.  Every shuffle and every
// shared-memory access then serves two states (a load in which the two half-warps read two different addresses is
// still one wavefront), which halves the data-pipe traffic that bounds rbw_fd_kernel.
//   * sums along the chain: lane r owns joints 2r and 2r+1; pair totals are scanned over 16 lanes (4 steps), the
//     pair is finished locally;
//   * H and the elimination: lane r owns rows r and r+16 (16 + 32 register entries: the work per lane is even);
//     the hand-over between the two layouts goes through shared memory.
#ifndef RBW_HWARPS
#define RBW_HWARPS 8
#endif
constexpr int RBH_LTS = 16 * RBW_LDL;                                     // stored columns of L D of one state, folded
constexpr int RBH_LT = 2 * RBH_LTS;
constexpr int RBH_PER_WARP = RBH_LT + 2 * 32 * 6 + 4 * 32 * RBW_IOS;      // + screws + staging (in 3, out 1)

// Inclusive prefix sums along the chain for the lane's two joints (x0: joint 2r, x1: joint 2r+1).
template <int K>
__device__ __forceinline__ void prefix2(double (&x0)[K], double (&x1)[K], int r) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        double t = x0[c] + x1[c];
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const double u = up16(t, d);
            if (r >= d) t += u;
        }
        const double e = up16(t, 1);
        x0[c] += r ? e : 0.0;
        x1[c] += x0[c];
    }
}
template <int K>
__device__ __forceinline__ void suffix2(double (&x0)[K], double (&x1)[K], int r) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        double t = x0[c] + x1[c];
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const double u = dn16(t, d);
            if (r + d < 16) t += u;
        }
        const double e = dn16(t, 1);
        x1[c] += r < 15 ? e : 0.0;
        x0[c] += x1[c];
    }
}
__global__ void __launch_bounds__(32 * RBW_HWARPS, 1)
rbh_fd_kernel(const double* __restrict__ model, int n, const double* __restrict__ q, const double* __restrict__ dq,
              const double* __restrict__ tau, double* __restrict__ qdd, size_t B, size_t ld, int* __restrict__ status) {
    extern __shared__ __align__(16) double rbw_sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int h = lane >> 4, r = lane & 15;
    double* msm = rbw_sm;                                    // [24][32] model constants, [23][*] = composite mass
    double* wsm = rbw_sm + RBH_MODEL + (size_t)w * RBH_PER_WARP;
    // this state's columns of L D, folded into 16 rows: column k < 16 sits in row k at positions k..31, column k >= 16 in
    // the unused head of row 31-k at positions 0..31-k (its last entry lands on row 31-k's diagonal, long dead by then)
    double* Lt = wsm + h * RBH_LTS;
    double* Sb = wsm + RBH_LT + h * 32 * 6;                  // this state's [32][6] screws
    double* Fb = Lt;                                         // hand-over buffer [7][32] (dead before Lt is written)
    double* io = wsm + RBH_LT + 2 * 32 * 6;                  // [3][32][RBW_IOS]
    double* ob = io + 3 * 32 * RBW_IOS;                      // [32][RBW_IOS]
    if (w == 0) {                                            // model -> shared memory; idle joints: identity, no mass
        const bool act = lane < n;
        const double* row = model + (size_t)(act ? lane : 0) * 24;
        double mc[1] = {act ? row[12] : 0.0};
        suffix_sum<1>(mc, lane);
#pragma unroll
        for (int e = 0; e < 13; ++e) msm[e * 32 + lane] = act ? row[e] : ((e == 0 || e == 4 || e == 8) ? 1.0 : 0.0);
#pragma unroll
        for (int e = 0; e < 9; ++e) msm[(13 + e) * 32 + lane] = act ? row[14 + e] : 0.0;
        msm[22 * 32 + lane] = mc[0];
    }
    __syncthreads();
    const double g[3] = {model[(size_t)n * 24], model[(size_t)n * 24 + 1], model[(size_t)n * 24 + 2]};
    const int j0 = 2 * r;                                    // the lane's joints in the chain phase: j0, j0 + 1
    const bool act0 = j0 < n, act1 = j0 + 1 < n;
    auto mdl2 = [&](int e) { return *reinterpret_cast<const double2*>(msm + e * 32 + j0); };

    const size_t groups = (B + RBW_GROUP - 1) / RBW_GROUP;
    bool all_ok = true;
    for (size_t grp = (size_t)blockIdx.x * RBW_HWARPS + w; grp < groups; grp += (size_t)gridDim.x * RBW_HWARPS) {
        const size_t s0 = grp * RBW_GROUP;
        {
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
        for (int st2 = 0; st2 < in_group; st2 += 2) {
            // a half-warp whose state does not exist (odd tail) recomputes its neighbour's and stores nothing
            const bool live = st2 + h < in_group;
            const int st = live ? st2 + h : st2;
            // ================= chain phase: lane r <-> joints j0, j0 + 1 of state st
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
                // pair product, exclusive prefix product over the 16 lanes, then the two world poses
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
            prefix2<3>(om0, om1, r);
            double al0[3], al1[3];
            cross(om0, zq0, al0);
            cross(om1, zq1, al1);
            prefix2<3>(al0, al1, r);
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
            prefix2<3>(ac0, ac1, r);
            double fw0[6], fw1[6], ci0[9], ci1[9];
            {
                const double2 m = mdl2(12), h0 = mdl2(13), h1 = mdl2(14), h2 = mdl2(15);
                const double2 i0 = mdl2(16), i1 = mdl2(17), i2 = mdl2(18), i3 = mdl2(19), i4 = mdl2(20), i5 = mdl2(21);
                const double ha[3] = {h0.x, h1.x, h2.x}, hb[3] = {h0.y, h1.y, h2.y};
                const double Ia[6] = {i0.x, i1.x, i2.x, i3.x, i4.x, i5.x}, Ib[6] = {i0.y, i1.y, i2.y, i3.y, i4.y, i5.y};
                link_terms(R0, p0, m.x, ha, Ia, om0, al0, ac0, fw0, ci0);
                link_terms(R1, p1, m.y, hb, Ib, om1, al1, ac1, fw1, ci1);
            }
            suffix2<6>(fw0, fw1, r);
            const double b0 = io[(2 * 32 + j0) * RBW_IOS + st]
                              - (z0[0] * fw0[3] + z0[1] * fw0[4] + z0[2] * fw0[5] + v0[0] * fw0[0] + v0[1] * fw0[1] + v0[2] * fw0[2]);
            const double b1 = io[(2 * 32 + j0 + 1) * RBW_IOS + st]
                              - (z1[0] * fw1[3] + z1[1] * fw1[4] + z1[2] * fw1[5] + v1[0] * fw1[0] + v1[1] * fw1[1] + v1[2] * fw1[2]);
            suffix2<9>(ci0, ci1, r);
            {
                const double2 mc = mdl2(22);
                double Fn0[3], Ff0[3], Fn1[3], Ff1[3];
                double2* S2 = reinterpret_cast<double2*>(Sb + j0 * 6);
                comp_times_screw(mc.x, ci0, z0, v0, Fn0, Ff0);
                comp_times_screw(mc.y, ci1, z1, v1, Fn1, Ff1);
                S2[0] = make_double2(z0[0], z0[1]); S2[1] = make_double2(z0[2], v0[0]); S2[2] = make_double2(v0[1], v0[2]);
                S2[3] = make_double2(z1[0], z1[1]); S2[4] = make_double2(z1[2], v1[0]); S2[5] = make_double2(v1[1], v1[2]);
                // hand-over buffer, component-major [7][32]: the lane's two joints are one 16-byte store per component
                // (a joint-major [32][8] layout cost 16-way bank conflicts on both sides)
                double2* F2 = reinterpret_cast<double2*>(Fb + j0);
                F2[0 * 16] = make_double2(Fn0[0], Fn1[0]); F2[1 * 16] = make_double2(Fn0[1], Fn1[1]); F2[2 * 16] = make_double2(Fn0[2], Fn1[2]);
                F2[3 * 16] = make_double2(Ff0[0], Ff1[0]); F2[4 * 16] = make_double2(Ff0[1], Ff1[1]); F2[5 * 16] = make_double2(Ff0[2], Ff1[2]);
                F2[6 * 16] = make_double2(act0 ? b0 : 0.0, act1 ? b1 : 0.0);
            }
            __syncwarp();
            // ================= matrix phase: lane r <-> rows r (lo) and r + 16 (hi) of state st
            double alo[16], ahi[32], blo, bhi;
            {
                double2 l0, l1, l2, h0, h1, h2;
                l0.x = Fb[0 * 32 + r]; l0.y = Fb[1 * 32 + r]; l1.x = Fb[2 * 32 + r]; l1.y = Fb[3 * 32 + r]; l2.x = Fb[4 * 32 + r]; l2.y = Fb[5 * 32 + r];
                h0.x = Fb[0 * 32 + r + 16]; h0.y = Fb[1 * 32 + r + 16]; h1.x = Fb[2 * 32 + r + 16]; h1.y = Fb[3 * 32 + r + 16];
                h2.x = Fb[4 * 32 + r + 16]; h2.y = Fb[5 * 32 + r + 16];
                blo = Fb[6 * 32 + r]; bhi = Fb[6 * 32 + r + 16];
                __syncwarp();                                // Fb (= Lt) is free from here on
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double2* S2 = reinterpret_cast<const double2*>(Sb + j * 6);
                    const double2 s01 = S2[0], s23 = S2[1], s45 = S2[2];
                    ahi[j] = fma(s01.x, h0.x, fma(s01.y, h0.y, fma(s23.x, h1.x, fma(s23.y, h1.y, fma(s45.x, h2.x, s45.y * h2.y)))));
                    if (j < 16)
                        alo[j] = fma(s01.x, l0.x, fma(s01.y, l0.y, fma(s23.x, l1.x, fma(s23.y, l1.y, fma(s45.x, l2.x, s45.y * l2.y)))));
                }
                if (n < 32) {                                // idle joints: identity rows
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        ahi[j] = (r + 16 >= n && j == r + 16) ? 1.0 : ahi[j];
                        if (j < 16) alo[j] = (r >= n && j == r) ? 1.0 : alo[j];
                    }
                }
            }
            bool ok = true;
            double dinv_lo = 0.0, dinv_hi = 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                // column k, entry i at COL + i
                const int COL = k < 16 ? k * RBW_LDL : (31 - k) * RBW_LDL - k;
                if (k < 16) {
                    Lt[COL + r] = alo[k & 15];
                    Lt[COL + r + 16] = ahi[k];
                } else if (r + 16 >= k) {
                    Lt[COL + r + 16] = ahi[k];
                }
                const double bk = __shfl_sync(FULL, k < 16 ? blo : bhi, k & 15, 16);
                __syncwarp();
                const double d = Lt[COL + k];
                ok = ok && (d > 0.0);
                const double dinv = rb_rcp_pos(d);
                if (k < 16) { if (r == k) dinv_lo = dinv; } else { if (r + 16 == k) dinv_hi = dinv; }
                const double nhi = (r + 16 > k) ? -ahi[k] * dinv : 0.0;
                bhi = fma(nhi, bk, bhi);
                double nlo = 0.0;
                if (k < 15) { nlo = (r > k) ? -alo[k & 15] * dinv : 0.0; blo = fma(nlo, bk, blo); }
                const bool aligned = ((COL & 1) == 0);       // pairs (even i, i + 1) are 16-byte aligned
                if (((k + 1) & 1) && k + 1 < 32) {
                    const double c1 = Lt[COL + k + 1];
                    ahi[(k + 1) & 31] = fma(nhi, c1, ahi[(k + 1) & 31]);
                    if (k + 1 < 16) alo[(k + 1) & 15] = fma(nlo, c1, alo[(k + 1) & 15]);
                }
#pragma unroll
                for (int i = (k + 2) & ~1; i < 32; i += 2) {
                    double2 c2;
                    if (aligned) c2 = *reinterpret_cast<const double2*>(Lt + COL + i);
                    else { c2.x = Lt[COL + i]; c2.y = Lt[COL + i + 1]; }
                    ahi[i] = fma(nhi, c2.x, ahi[i]);
                    ahi[i + 1] = fma(nhi, c2.y, ahi[i + 1]);
                    if (i + 1 < 16) {
                        alo[i & 15] = fma(nlo, c2.x, alo[i & 15]);
                        alo[(i + 1) & 15] = fma(nlo, c2.y, alo[(i + 1) & 15]);
                    }
                }
            }
            // back substitution: x_k = (y_k - sum_{i>k} (l_ik d_k) x_i) / d_k, rows r and r + 16 per lane
            double xlo = blo, xhi = bhi;
#pragma unroll
            for (int i = 31; i >= 1; --i) {
                const double mine = i < 16 ? xlo * dinv_lo : xhi * dinv_hi;
                const double xi = __shfl_sync(FULL, mine, i & 15, 16);
                if (i < 16) { if (r == i) xlo = xi; } else { if (r + 16 == i) xhi = xi; }
                if (i > 16 && r + 16 < i) xhi = fma(-Lt[(15 - r) * RBW_LDL - (r + 16) + i], xi, xhi);   // column r + 16, entry i
                const double c = Lt[r * RBW_LDL + i];                                                  // column r, entry i
                if (r < i) xlo = fma(-c, xi, xlo);
            }
            if (r == 0) xlo *= dinv_lo;
            all_ok = all_ok && (ok || !live);
            if (live) {
                ob[r * RBW_IOS + st] = ok ? xlo : rb_nan<double>();
                ob[(r + 16) * RBW_IOS + st] = ok ? xhi : rb_nan<double>();
            }
            __syncwarp();
        }
        {
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
    if (!all_ok && r == 0) atomicOr(status, RB_STATUS_NOT_SPD);
}

