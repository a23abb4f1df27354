// rb_deriv.cuh -- analytical first derivatives of inverse and forward dynamics (the reference
// lists "Differentiability" as not done, README.md:18).
//
//   rb_rnea_derivatives:  d tau / d q  and  d tau / d dq  of  tau = rnea(q, dq, ddq)   (multibody.rs:111-153)
//   forward dynamics:     d qdd / d x = -H^-1 (d tau / d x at ddq = qdd),  d qdd / d tau = H^-1   (rb_kernels.cuh)
//
// Thread per state like the other kernels, but the algebra is done in the WORLD frame about the world origin, where
// the recursions of rnea / crba become sums along the chain and every partial derivative has a closed form.
// Spatial vectors are (angular ; linear-at-origin), forces (moment-about-origin ; force).  With the joint screws
// s_k = (z_k ; p_k x z_k), body velocities v_k = sum_{l<=k} s_l dq_l, accelerations a_k (base: (0 ; g)), world
// inertias I_k, momenta h_k = I_k v_k, wrenches f_k = I_k a_k + v_k x* h_k and the inertia rates
// Idot_k = v_k x* I_k - I_k v_k x (again a 10-parameter object: mass rate 0, first-moment rate, 6 inertia rates),
// all suffix-summed into composites (^c):
//
//   d/d dq_j:  dv = s_j                 da0 = 2 v_j x s_j
//   d/d  q_j:  dv = v_j x s_j           da0 = a_j x s_j + v_j x (v_j x s_j)
//   for both:  d f_k = [s_j x* f_k] + I_k da0 + Idot_k dv + dv x* h_k     (k >= j; bracket for d/dq only)
//   A_m = I^c_m s_m,  B_m = Idot^c_m s_m - s_m x* h^c_m,  G_j = sum_{k>=j} d f_k
//   d tau_m / d x_j = s_m . G_j             (m <= j)
//                   = da0_j . A_m + dv_j . B_m   (m > j)
//
// (the term from d s_m / d q_j cancels against s_j x* F^c_m).  About 8 RNEA evaluations of work for both matrices.
// tests/ check it against complex-step differentiation of an independent link-frame rnea.
#pragma once
#include "rb_dyn.cuh"

RB_DI void rb_cross3(const double (&a)[3], const double (&b)[3], double (&o)[3]) {
    o[0] = fma(a[1], b[2], -(a[2] * b[1]));
    o[1] = fma(a[2], b[0], -(a[0] * b[2]));
    o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}
RB_DI void rb_cross3_acc(const double (&a)[3], const double (&b)[3], double (&o)[3]) {
    o[0] = fma(a[1], b[2], fma(-a[2], b[1], o[0]));
    o[1] = fma(a[2], b[0], fma(-a[0], b[2], o[1]));
    o[2] = fma(a[0], b[1], fma(-a[1], b[0], o[2]));
}
// o = a x b for motion vectors:  (w1 x w2 ; w1 x v2 + v1 x w2)
RB_DI void rb_mcross(const double (&a)[6], const double (&b)[6], double (&o)[6]) {
    const double aw[3] = {a[0], a[1], a[2]}, av[3] = {a[3], a[4], a[5]};
    const double bw[3] = {b[0], b[1], b[2]}, bv[3] = {b[3], b[4], b[5]};
    double w[3], l[3];
    rb_cross3(aw, bw, w);
    rb_cross3(aw, bv, l);
    rb_cross3_acc(av, bw, l);
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2]; o[3] = l[0]; o[4] = l[1]; o[5] = l[2];
}
// o (+)= a x* f for a motion a and a force f = (n ; f):  (w x n + v x f ; w x f)
template <bool ACC>
RB_DI void rb_fcross(const double (&a)[6], const double (&f)[6], double (&o)[6]) {
    const double aw[3] = {a[0], a[1], a[2]}, av[3] = {a[3], a[4], a[5]};
    const double fn[3] = {f[0], f[1], f[2]}, ff[3] = {f[3], f[4], f[5]};
    double n[3] = {ACC ? o[0] : 0.0, ACC ? o[1] : 0.0, ACC ? o[2] : 0.0};
    double l[3] = {ACC ? o[3] : 0.0, ACC ? o[4] : 0.0, ACC ? o[5] : 0.0};
    rb_cross3_acc(aw, fn, n);
    rb_cross3_acc(av, ff, n);
    rb_cross3_acc(aw, ff, l);
    o[0] = n[0]; o[1] = n[1]; o[2] = n[2]; o[3] = l[0]; o[4] = l[1]; o[5] = l[2];
}
// o (+)= I u for a 10-parameter inertia about the origin (mass m, first moment H, inertia S = xx xy xz yy yz zz):
//   n = S w + H x v,  f = m v - H x w.   The inertia RATE is the same object with m = 0.
template <bool ACC>
RB_DI void rb_wimul(double m, const double (&H)[3], const double (&S)[6], const double (&u)[6], double (&o)[6]) {
    const double w[3] = {u[0], u[1], u[2]}, v[3] = {u[3], u[4], u[5]};
    double n[3] = {ACC ? o[0] : 0.0, ACC ? o[1] : 0.0, ACC ? o[2] : 0.0};
    double f[3] = {ACC ? o[3] : 0.0, ACC ? o[4] : 0.0, ACC ? o[5] : 0.0};
    n[0] = fma(S[0], w[0], fma(S[1], w[1], fma(S[2], w[2], n[0])));
    n[1] = fma(S[1], w[0], fma(S[3], w[1], fma(S[4], w[2], n[1])));
    n[2] = fma(S[2], w[0], fma(S[4], w[1], fma(S[5], w[2], n[2])));
    rb_cross3_acc(H, v, n);
    f[0] = fma(m, v[0], f[0]); f[1] = fma(m, v[1], f[1]); f[2] = fma(m, v[2], f[2]);
    rb_cross3_acc(w, H, f);                                  // -H x w = w x H
    o[0] = n[0]; o[1] = n[1]; o[2] = n[2]; o[3] = f[0]; o[4] = f[1]; o[5] = f[2];
}
RB_DI double rb_dot6(const double (&a)[6], const double (&b)[6]) {
    return fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], fma(a[4], b[4], a[5] * b[5])))));
}

// put_q(m, j, v): d tau_m / d q_j;  put_v(m, j, v): d tau_m / d dq_j  (m, j are RbIC constants).
template <class M, class PutQ, class PutV>
RB_DI void rb_rnea_derivatives(const typename M::Param& p, const double (&sn)[M::N], const double (&cs)[M::N],
                               const double (&dq)[M::N], const double (&ddq)[M::N], PutQ&& put_q, PutV&& put_v) {
    constexpr int N = M::N;
    double S[N][6], DV[N][6], DA[N][6];                     // screws; dv and da0 of the d/dq columns (d/d dq: S, 2 DV)
    double R[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}, P[3] = {0.0, 0.0, 0.0};
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double a[6] = {0.0, 0.0, 0.0, M::template g<0>(p), M::template g<1>(p), M::template g<2>(p)};
    // ---- outward: world poses, screws, velocities, accelerations
    rb_for_up<0, N>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        double T[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            P[r] = k_dot3_acc<KC(I, RB_F_T, 0), KC(I, RB_F_T, 1), KC(I, RB_F_T, 2)>(P[r], KV(I, RB_F_T, 0), KV(I, RB_F_T, 1), KV(I, RB_F_T, 2), R[r][0], R[r][1], R[r][2]);
            T[r][0] = k_dot3<KC(I, RB_F_R, 0), KC(I, RB_F_R, 3), KC(I, RB_F_R, 6)>(KV(I, RB_F_R, 0), KV(I, RB_F_R, 3), KV(I, RB_F_R, 6), R[r][0], R[r][1], R[r][2]);
            T[r][1] = k_dot3<KC(I, RB_F_R, 1), KC(I, RB_F_R, 4), KC(I, RB_F_R, 7)>(KV(I, RB_F_R, 1), KV(I, RB_F_R, 4), KV(I, RB_F_R, 7), R[r][0], R[r][1], R[r][2]);
            T[r][2] = k_dot3<KC(I, RB_F_R, 2), KC(I, RB_F_R, 5), KC(I, RB_F_R, 8)>(KV(I, RB_F_R, 2), KV(I, RB_F_R, 5), KV(I, RB_F_R, 8), R[r][0], R[r][1], R[r][2]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            R[r][0] = fma(cs[I], T[r][0], sn[I] * T[r][1]);
            R[r][1] = fma(cs[I], T[r][1], -(sn[I] * T[r][0]));
            R[r][2] = T[r][2];
        }
        const double z[3] = {R[0][2], R[1][2], R[2][2]};
        double pz[3];
        rb_cross3(P, z, pz);
        S[I][0] = z[0]; S[I][1] = z[1]; S[I][2] = z[2]; S[I][3] = pz[0]; S[I][4] = pz[1]; S[I][5] = pz[2];
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = fma(S[I][k], dq[I], v[k]);
        rb_mcross(v, S[I], DV[I]);                           // v_I x s_I
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] = fma(S[I][k], ddq[I], fma(DV[I][k], dq[I], a[k]));
        double t6[6];
        rb_mcross(a, S[I], DA[I]);
        rb_mcross(v, DV[I], t6);
#pragma unroll
        for (int k = 0; k < 6; ++k) DA[I][k] += t6[k];
    });
    // ---- inward: composites, the per-joint vectors, and the entries
    double Hc[3] = {0.0, 0.0, 0.0}, Ic[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};           // I^c (mass: model constant)
    double dHc[3] = {0.0, 0.0, 0.0}, dIc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};         // Idot^c
    double hc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, Fc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    rb_for_down<N - 1>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        {   // link I about the world origin
            const double m = KV(I, RB_F_M, 0);
            const double h[3] = {KV(I, RB_F_H, 0), KV(I, RB_F_H, 1), KV(I, RB_F_H, 2)};
            const double Io[6] = {KV(I, RB_F_I, 0), KV(I, RB_F_I, 1), KV(I, RB_F_I, 2), KV(I, RB_F_I, 3), KV(I, RB_F_I, 4), KV(I, RB_F_I, 5)};
            double hw[3], T[3][3], Hm[3], IO[6];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                hw[r] = fma(R[r][0], h[0], fma(R[r][1], h[1], R[r][2] * h[2]));
                T[r][0] = fma(R[r][0], Io[0], fma(R[r][1], Io[1], R[r][2] * Io[2]));
                T[r][1] = fma(R[r][0], Io[1], fma(R[r][1], Io[3], R[r][2] * Io[4]));
                T[r][2] = fma(R[r][0], Io[2], fma(R[r][1], Io[4], R[r][2] * Io[5]));
            }
            double u[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { Hm[r] = fma(m, P[r], hw[r]); u[r] = fma(0.5 * m, P[r], hw[r]); }
            const double pu2 = 2.0 * fma(P[0], u[0], fma(P[1], u[1], P[2] * u[2]));
            auto rr = [&](int x, int y) { return fma(T[x][0], R[y][0], fma(T[x][1], R[y][1], T[x][2] * R[y][2])); };
            IO[0] = rr(0, 0) - 2.0 * P[0] * u[0] + pu2;
            IO[1] = rr(0, 1) - fma(P[0], u[1], u[0] * P[1]);
            IO[2] = rr(0, 2) - fma(P[0], u[2], u[0] * P[2]);
            IO[3] = rr(1, 1) - 2.0 * P[1] * u[1] + pu2;
            IO[4] = rr(1, 2) - fma(P[1], u[2], u[1] * P[2]);
            IO[5] = rr(2, 2) - 2.0 * P[2] * u[2] + pu2;
            // momentum, wrench, inertia rate of the link; accumulate the composites
            double h6[6], f6[6];
            rb_wimul<false>(m, Hm, IO, v, h6);
            rb_wimul<false>(m, Hm, IO, a, f6);
            rb_fcross<true>(v, h6, f6);
            const double w[3] = {v[0], v[1], v[2]}, vo[3] = {v[3], v[4], v[5]};
            double dH[3] = {m * vo[0], m * vo[1], m * vo[2]};
            rb_cross3_acc(w, Hm, dH);
            // d IO = [w]x IO + ([w]x IO)^T + 2 (H . vo) Id - vo H^T - H vo^T
            const double c0[3] = {IO[0], IO[1], IO[2]}, c1[3] = {IO[1], IO[3], IO[4]}, c2[3] = {IO[2], IO[4], IO[5]};
            double W0[3], W1[3], W2[3];                      // columns of [w]x IO
            rb_cross3(w, c0, W0); rb_cross3(w, c1, W1); rb_cross3(w, c2, W2);
            const double hv2 = 2.0 * fma(Hm[0], vo[0], fma(Hm[1], vo[1], Hm[2] * vo[2]));
            dIc[0] += 2.0 * W0[0] + hv2 - 2.0 * vo[0] * Hm[0];
            dIc[1] += W1[0] + W0[1] - fma(vo[0], Hm[1], Hm[0] * vo[1]);
            dIc[2] += W2[0] + W0[2] - fma(vo[0], Hm[2], Hm[0] * vo[2]);
            dIc[3] += 2.0 * W1[1] + hv2 - 2.0 * vo[1] * Hm[1];
            dIc[4] += W2[1] + W1[2] - fma(vo[1], Hm[2], Hm[1] * vo[2]);
            dIc[5] += 2.0 * W2[2] + hv2 - 2.0 * vo[2] * Hm[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) { Hc[k] += Hm[k]; dHc[k] += dH[k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) { Ic[k] += IO[k]; hc[k] += h6[k]; Fc[k] += f6[k]; }
        }
        const double mc = KV(I, RB_F_M, 1);
        double A[6], Bv[6], G[6], Gp[6], dv2[6];
        rb_wimul<false>(mc, Hc, Ic, S[I], A);
        rb_wimul<false>(0.0, dHc, dIc, S[I], Bv);            // Idot^c s
#pragma unroll
        for (int k = 0; k < 6; ++k) { Gp[k] = Bv[k]; dv2[k] = 2.0 * DV[I][k]; }
        double sh[6];
        rb_fcross<false>(S[I], hc, sh);                      // s x* h^c
#pragma unroll
        for (int k = 0; k < 6; ++k) { Bv[k] -= sh[k]; Gp[k] += sh[k]; }
        rb_wimul<true>(mc, Hc, Ic, dv2, Gp);
        rb_fcross<false>(S[I], Fc, G);
        rb_wimul<true>(mc, Hc, Ic, DA[I], G);
        rb_wimul<true>(0.0, dHc, dIc, DV[I], G);
        rb_fcross<true>(DV[I], hc, G);
        rb_for_up<0, I + 1>([&](auto mcn) {                  // rows m <= I of column I
            constexpr int Mm = decltype(mcn)::value;
            put_q(mcn, ic, rb_dot6(S[Mm], G));
            put_v(mcn, ic, rb_dot6(S[Mm], Gp));
        });
        rb_for_up<0, I>([&](auto jc) {                       // row I of the columns j < I
            constexpr int J = decltype(jc)::value;
            put_q(ic, jc, fma(DA[J][0], A[0], fma(DA[J][1], A[1], fma(DA[J][2], A[2], fma(DA[J][3], A[3], fma(DA[J][4], A[4], DA[J][5] * A[5]))))) + rb_dot6(DV[J], Bv));
            put_v(ic, jc, 2.0 * rb_dot6(DV[J], A) + rb_dot6(S[J], Bv));
        });
        if constexpr (I > 0) {
            // step back to link I-1: velocities, accelerations, pose
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                a[k] = fma(-S[I][k], ddq[I], fma(-DV[I][k], dq[I], a[k]));
                v[k] = fma(-S[I][k], dq[I], v[k]);
            }
            double U[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                U[r][0] = fma(cs[I], R[r][0], -(sn[I] * R[r][1]));
                U[r][1] = fma(sn[I], R[r][0], cs[I] * R[r][1]);
                U[r][2] = R[r][2];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                R[r][0] = k_dot3<KC(I, RB_F_R, 0), KC(I, RB_F_R, 1), KC(I, RB_F_R, 2)>(KV(I, RB_F_R, 0), KV(I, RB_F_R, 1), KV(I, RB_F_R, 2), U[r][0], U[r][1], U[r][2]);
                R[r][1] = k_dot3<KC(I, RB_F_R, 3), KC(I, RB_F_R, 4), KC(I, RB_F_R, 5)>(KV(I, RB_F_R, 3), KV(I, RB_F_R, 4), KV(I, RB_F_R, 5), U[r][0], U[r][1], U[r][2]);
                R[r][2] = k_dot3<KC(I, RB_F_R, 6), KC(I, RB_F_R, 7), KC(I, RB_F_R, 8)>(KV(I, RB_F_R, 6), KV(I, RB_F_R, 7), KV(I, RB_F_R, 8), U[r][0], U[r][1], U[r][2]);
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
                P[r] = -k_dot3_acc<KC(I, RB_F_T, 0), KC(I, RB_F_T, 1), KC(I, RB_F_T, 2)>(-P[r], KV(I, RB_F_T, 0), KV(I, RB_F_T, 1), KV(I, RB_F_T, 2), R[r][0], R[r][1], R[r][2]);
        }
    });
}

// LDL^T of the upper triangle in place (A(j,i) <- L(i,j) for i > j; the diagonal keeps d_j) and its application to
// right-hand sides: the factor-once / solve-many split of rb_ldlt_solve (rb_dyn.cuh), same operation order.
template <int N>
RB_DI bool rb_ldlt_factor(double (&A)[N][N], double (&dinv)[N]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double d = A[j][j];
        ok = ok && (d > 0.0);
        dinv[j] = rb_rcp_pos(d);
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            const double l = A[j][i] * dinv[j];
#pragma unroll
            for (int k = i; k < N; ++k) A[i][k] = fma(-l, A[j][k], A[i][k]);
            A[j][i] = l;
        }
    }
    return ok;
}
template <int N>
RB_DI void rb_ldlt_apply(const double (&A)[N][N], const double (&dinv)[N], double (&x)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = j + 1; i < N; ++i) x[i] = fma(-A[j][i], x[j], x[i]);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] *= dinv[j];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
#pragma unroll
        for (int k = i + 1; k < N; ++k) x[i] = fma(-A[i][k], x[k], x[i]);
    }
}
