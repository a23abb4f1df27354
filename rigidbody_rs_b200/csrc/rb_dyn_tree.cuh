// rb_dyn_tree.cuh -- the unrolled recursions of rb_dyn.cuh for kinematic trees (included at the end of rb_dyn.cuh).
//
// The reference is serial-only (f[i-1], multibody.rs:148; ic[i-1], :170).  A compile-time table that carries a parent
// index per joint (CtModel::parent<I>(), slot 23 of the row) lets the same straight-line code follow a tree: motion
// comes from the parent link, wrenches and composite inertias are accumulated into the parent, and H(j, i) is non-zero
// only where joint j supports link i.  The per-link arrays are indexed with compile-time constants, so they are
// registers; a serial table never reaches this file (M::kTree is false) and keeps its tuned code.
#pragma once

// f(RbIC<I>), then the same for the parent of I, ... up to a root.
template <class M, int I, class F> RB_DI void rb_for_anc(F&& f) {
    if constexpr (I >= 0) { f(RbIC<I>{}); rb_for_anc<M, M::template parent<I>()>(f); }
}

template <class M, bool HAS_DDQ>
RB_DI void rb_rnea_tree(const typename M::Param& p, const RB_R (&s)[M::N], const RB_R (&c)[M::N],
                        const RB_R (&dq)[M::N], const RB_R (&ddq)[M::N], RB_R (&tau)[M::N]) {
    constexpr int N = M::N;
    RB_R fl[N][3], fr[N][3], W[N][3], AL[N][3], AC[N][3];       // wrench; omega, alpha, a' of every link, own frame
    rb_for_up<0, N>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        constexpr int P = M::template parent<I>();
        const RB_R dqi = dq[I];
        RB_R b[3], wn[3] = {0.0, 0.0, 0.0}, aln[3] = {0.0, 0.0, 0.0};
        if constexpr (P < 0) {
            b[0] = M::template g<0>(p); b[1] = M::template g<1>(p); b[2] = M::template g<2>(p);      // base: a' = g (:116-120)
        } else {
            // a'_i = E (a' + alpha x t + w x (w x t)) with the parent's w, alpha, a'
            const RB_R t0 = KV(I, RB_F_T, 0), t1 = KV(I, RB_F_T, 1), t2 = KV(I, RB_F_T, 2);
            constexpr int T0 = KC(I, RB_F_T, 0), T1 = KC(I, RB_F_T, 1), T2 = KC(I, RB_F_T, 2);
            const RB_R u0 = k_fnma<T1>(t1, W[P][2], k_mul<T2>(t2, W[P][1]));
            const RB_R u1 = k_fnma<T2>(t2, W[P][0], k_mul<T0>(t0, W[P][2]));
            const RB_R u2 = k_fnma<T0>(t0, W[P][1], k_mul<T1>(t1, W[P][0]));
            b[0] = k_fma<T2>(t2, AL[P][1], k_fnma<T1>(t1, AL[P][2], AC[P][0]));
            b[1] = k_fma<T0>(t0, AL[P][2], k_fnma<T2>(t2, AL[P][0], AC[P][1]));
            b[2] = k_fma<T1>(t1, AL[P][0], k_fnma<T0>(t0, AL[P][1], AC[P][2]));
            b[0] = fma(W[P][1], u2, fma(-W[P][2], u1, b[0]));
            b[1] = fma(W[P][2], u0, fma(-W[P][0], u2, b[1]));
            b[2] = fma(W[P][0], u1, fma(-W[P][1], u0, b[2]));
            rb_rotate_in<M, I>(p, s[I], c[I], W[P], wn);                               // :129
            rb_rotate_in<M, I>(p, s[I], c[I], AL[P], aln);                             // :132
        }
        rb_rotate_in<M, I>(p, s[I], c[I], b, AC[I]);
        // alpha_i = E alpha + z ddq + (E w) x z dq   (:133, :137-138);  w_i = E w + z dq (:130)
        AL[I][0] = fma(wn[1], dqi, aln[0]);
        AL[I][1] = fma(-wn[0], dqi, aln[1]);
        AL[I][2] = HAS_DDQ ? aln[2] + ddq[I] : aln[2];
        W[I][0] = wn[0]; W[I][1] = wn[1]; W[I][2] = wn[2] + dqi;
        const RB_R (&w)[3] = W[I];
        const RB_R (&al)[3] = AL[I];
        const RB_R (&ac)[3] = AC[I];
        // wrench of link i about its origin (:140)
        const RB_R m = KV(I, RB_F_M, 0);
        const RB_R h0 = KV(I, RB_F_H, 0), h1 = KV(I, RB_F_H, 1), h2 = KV(I, RB_F_H, 2);
        const RB_R Ixx = KV(I, RB_F_I, 0), Ixy = KV(I, RB_F_I, 1), Ixz = KV(I, RB_F_I, 2);
        const RB_R Iyy = KV(I, RB_F_I, 3), Iyz = KV(I, RB_F_I, 4), Izz = KV(I, RB_F_I, 5);
        const RB_R e0 = fma(w[1], h2, -(w[2] * h1));
        const RB_R e1 = fma(w[2], h0, -(w[0] * h2));
        const RB_R e2 = fma(w[0], h1, -(w[1] * h0));
        fl[I][0] = fma(w[1], e2, fma(-w[2], e1, fma(al[1], h2, fma(-al[2], h1, m * ac[0]))));
        fl[I][1] = fma(w[2], e0, fma(-w[0], e2, fma(al[2], h0, fma(-al[0], h2, m * ac[1]))));
        fl[I][2] = fma(w[0], e1, fma(-w[1], e0, fma(al[0], h1, fma(-al[1], h0, m * ac[2]))));
        const RB_R L0 = fma(Ixz, w[2], fma(Ixy, w[1], Ixx * w[0]));
        const RB_R L1 = fma(Iyz, w[2], fma(Iyy, w[1], Ixy * w[0]));
        const RB_R L2 = fma(Izz, w[2], fma(Iyz, w[1], Ixz * w[0]));
        fr[I][0] = fma(h1, ac[2], fma(-h2, ac[1], fma(w[1], L2, fma(-w[2], L1, fma(Ixz, al[2], fma(Ixy, al[1], Ixx * al[0]))))));
        fr[I][1] = fma(h2, ac[0], fma(-h0, ac[2], fma(w[2], L0, fma(-w[0], L2, fma(Iyz, al[2], fma(Iyy, al[1], Ixy * al[0]))))));
        fr[I][2] = fma(h0, ac[1], fma(-h1, ac[0], fma(w[0], L1, fma(-w[1], L0, fma(Izz, al[2], fma(Iyz, al[1], Ixz * al[0]))))));
    });
    rb_for_down<N - 1>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        constexpr int P = M::template parent<I>();
        tau[I] = fr[I][2];                                                           // :144
        if constexpr (P >= 0) {
            RB_R tl[3], tr[3];
            rb_force<M, I>(p, s[I], c[I], fl[I], fr[I], tl, tr);                     // :147
            fl[P][0] += tl[0]; fl[P][1] += tl[1]; fl[P][2] += tl[2];                 // :148, into the parent
            fr[P][0] += tr[0]; fr[P][1] += tr[1]; fr[P][2] += tr[2];
        }
    });
}

// Composite of every sub-tree kept as (h, I_o) in the frame of its root link; children are folded into their parent
// with the same 10-parameter map as rb_crba_put.  put(RbIC<J>, RbIC<I>, v) is called for J = I and for every J that
// supports I; all other entries of H are zero and are not reported.
template <class M, class Put>
RB_DI void rb_crba_put_tree(const typename M::Param& p, const RB_R (&s)[M::N], const RB_R (&c)[M::N], Put&& put) {
    constexpr int N = M::N;
    RB_R hc[N][3], Ic[N][6];
    rb_for_up<0, N>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        hc[I][0] = KV(I, RB_F_H, 0); hc[I][1] = KV(I, RB_F_H, 1); hc[I][2] = KV(I, RB_F_H, 2);
        Ic[I][0] = KV(I, RB_F_I, 0); Ic[I][1] = KV(I, RB_F_I, 1); Ic[I][2] = KV(I, RB_F_I, 2);
        Ic[I][3] = KV(I, RB_F_I, 3); Ic[I][4] = KV(I, RB_F_I, 4); Ic[I][5] = KV(I, RB_F_I, 5);
    });
    rb_for_down<N - 1>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        constexpr int P = M::template parent<I>();
        const RB_R Ixx = Ic[I][0], Ixy = Ic[I][1], Ixz = Ic[I][2], Iyy = Ic[I][3], Iyz = Ic[I][4], Izz = Ic[I][5];
        put(RbIC<I>{}, RbIC<I>{}, Izz);                                              // :161
        RB_R Fl[3] = {-hc[I][1], hc[I][0], 0.0};                                     // :162  F = I^c * S_z
        RB_R Fr[3] = {Ixz, Iyz, Izz};
        rb_for_anc<M, I>([&](auto jc) {                                              // up the supporting branch
            constexpr int J = decltype(jc)::value;
            constexpr int PJ = M::template parent<J>();
            if constexpr (PJ >= 0) {
                RB_R ol[3], orr[3];
                rb_force<M, J>(p, s[J], c[J], Fl, Fr, ol, orr);                      // :165
                Fl[0] = ol[0]; Fl[1] = ol[1]; Fl[2] = ol[2];
                Fr[0] = orr[0]; Fr[1] = orr[1]; Fr[2] = orr[2];
                put(RbIC<PJ>{}, RbIC<I>{}, Fr[2]);                                   // :166
            }
        });
        if constexpr (P >= 0) {                                                      // :169-171, into the parent
            const RB_R si = s[I], ci = c[I];
            const RB_R g0 = fma(ci, hc[I][0], -(si * hc[I][1])), g1 = fma(si, hc[I][0], ci * hc[I][1]), g2 = hc[I][2];
            const RB_R cs = ci * si, s2 = cs + cs, c2 = fma(ci, ci, -(si * si));
            const RB_R hm = RB_R(0.5) * (Ixx - Iyy), hp = RB_R(0.5) * (Ixx + Iyy);
            const RB_R u_ = fma(hm, c2, -(Ixy * s2));
            const RB_R a = hp + u_, e = hp - u_, b = fma(hm, s2, Ixy * c2);
            const RB_R d = fma(ci, Ixz, -(si * Iyz)), f = fma(si, Ixz, ci * Iyz), g = Izz;
#define RR(r, k) KV(I, RB_F_R, 3 * (r) + (k))
#define RC(r, k) KC(I, RB_F_R, 3 * (r) + (k))
            const RB_R q0 = k_dot3<RC(0, 0), RC(0, 1), RC(0, 2)>(RR(0, 0), RR(0, 1), RR(0, 2), g0, g1, g2);
            const RB_R q1 = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), g0, g1, g2);
            const RB_R q2 = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), g0, g1, g2);
            const RB_R P00 = k_dot3<RC(0, 0), RC(0, 1), RC(0, 2)>(RR(0, 0), RR(0, 1), RR(0, 2), a, b, d);
            const RB_R P01 = k_dot3<RC(0, 0), RC(0, 1), RC(0, 2)>(RR(0, 0), RR(0, 1), RR(0, 2), b, e, f);
            const RB_R P02 = k_dot3<RC(0, 0), RC(0, 1), RC(0, 2)>(RR(0, 0), RR(0, 1), RR(0, 2), d, f, g);
            const RB_R P10 = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), a, b, d);
            const RB_R P11 = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), b, e, f);
            const RB_R P12 = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), d, f, g);
            const RB_R P20 = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), a, b, d);
            const RB_R P21 = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), b, e, f);
            const RB_R P22 = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), d, f, g);
            const RB_R Jxx = k_dot3<RC(0, 0), RC(0, 1), RC(0, 2)>(RR(0, 0), RR(0, 1), RR(0, 2), P00, P01, P02);
            const RB_R Jxy = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), P00, P01, P02);
            const RB_R Jxz = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), P00, P01, P02);
            const RB_R Jyy = k_dot3<RC(1, 0), RC(1, 1), RC(1, 2)>(RR(1, 0), RR(1, 1), RR(1, 2), P10, P11, P12);
            const RB_R Jyz = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), P10, P11, P12);
            const RB_R Jzz = k_dot3<RC(2, 0), RC(2, 1), RC(2, 2)>(RR(2, 0), RR(2, 1), RR(2, 2), P20, P21, P22);
#undef RR
#undef RC
            const RB_R mc = KV(I, RB_F_M, 1);                                        // mass of the sub-tree rooted at I
            const RB_R t0 = KV(I, RB_F_T, 0), t1 = KV(I, RB_F_T, 1), t2 = KV(I, RB_F_T, 2);
            constexpr int T0 = KC(I, RB_F_T, 0), T1 = KC(I, RB_F_T, 1), T2 = KC(I, RB_F_T, 2);
            const RB_R hmc = RB_R(0.5) * mc;
            const RB_R u0 = k_fma<T0>(t0, hmc, q0), u1 = k_fma<T1>(t1, hmc, q1), u2 = k_fma<T2>(t2, hmc, q2);
            const RB_R tu0 = k_mul<T0>(t0, u0), tu1 = k_mul<T1>(t1, u1), tu2 = k_mul<T2>(t2, u2);
            Ic[P][0] += fma(RB_R(2), tu1 + tu2, Jxx);
            Ic[P][3] += fma(RB_R(2), tu0 + tu2, Jyy);
            Ic[P][5] += fma(RB_R(2), tu0 + tu1, Jzz);
            Ic[P][1] += k_fnma<T1>(t1, u0, k_fnma<T0>(t0, u1, Jxy));
            Ic[P][2] += k_fnma<T2>(t2, u0, k_fnma<T0>(t0, u2, Jxz));
            Ic[P][4] += k_fnma<T2>(t2, u1, k_fnma<T1>(t1, u2, Jyz));
            hc[P][0] += k_fma<T0>(t0, mc, q0);
            hc[P][1] += k_fma<T1>(t1, mc, q1);
            hc[P][2] += k_fma<T2>(t2, mc, q2);
        }
    });
}

// Tip = last link; its pose is composed along the supporting branch only.
template <class M>
RB_DI void rb_fwd_kin_tree(const typename M::Param& p, const RB_R (&s)[M::N], const RB_R (&c)[M::N], RB_R (&pos)[3]) {
    constexpr int N = M::N;
    pos[0] = 0.0; pos[1] = 0.0; pos[2] = 0.0;
    rb_for_anc<M, N - 1>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        const RB_R y0 = fma(c[I], pos[0], -(s[I] * pos[1])), y1 = fma(s[I], pos[0], c[I] * pos[1]), y2 = pos[2];
        pos[0] = k_dot3_acc<KC(I, RB_F_R, 0), KC(I, RB_F_R, 1), KC(I, RB_F_R, 2)>(KV(I, RB_F_T, 0), KV(I, RB_F_R, 0), KV(I, RB_F_R, 1), KV(I, RB_F_R, 2), y0, y1, y2);
        pos[1] = k_dot3_acc<KC(I, RB_F_R, 3), KC(I, RB_F_R, 4), KC(I, RB_F_R, 5)>(KV(I, RB_F_T, 1), KV(I, RB_F_R, 3), KV(I, RB_F_R, 4), KV(I, RB_F_R, 5), y0, y1, y2);
        pos[2] = k_dot3_acc<KC(I, RB_F_R, 6), KC(I, RB_F_R, 7), KC(I, RB_F_R, 8)>(KV(I, RB_F_T, 2), KV(I, RB_F_R, 6), KV(I, RB_F_R, 7), KV(I, RB_F_R, 8), y0, y1, y2);
    });
}

// Tip-frame Jacobian: columns of the joints that support the tip; the others are zero.
template <class M>
RB_DI void rb_jac_tree(const typename M::Param& p, const RB_R (&s)[M::N], const RB_R (&c)[M::N], RB_R (&J)[M::N][6]) {
    constexpr int N = M::N;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) J[i][k] = RB_R(0);
    RB_R A[3][3] = {{M::template tip<0>(p), M::template tip<1>(p), M::template tip<2>(p)},
                    {M::template tip<3>(p), M::template tip<4>(p), M::template tip<5>(p)},
                    {M::template tip<6>(p), M::template tip<7>(p), M::template tip<8>(p)}};
    RB_R r[3] = {0.0, 0.0, 0.0};
    rb_for_anc<M, N - 1>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            J[I][k] = fma(A[1][k], r[0], -(A[0][k] * r[1]));
            J[I][3 + k] = A[2][k];
        }
        if constexpr (M::template parent<I>() >= 0) {
            RB_R B[3][3], y[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                B[0][k] = fma(c[I], A[0][k], -(s[I] * A[1][k]));
                B[1][k] = fma(s[I], A[0][k], c[I] * A[1][k]);
                B[2][k] = A[2][k];
            }
            y[0] = fma(c[I], r[0], -(s[I] * r[1])); y[1] = fma(s[I], r[0], c[I] * r[1]); y[2] = r[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                A[0][k] = k_dot3<KC(I, RB_F_R, 0), KC(I, RB_F_R, 1), KC(I, RB_F_R, 2)>(KV(I, RB_F_R, 0), KV(I, RB_F_R, 1), KV(I, RB_F_R, 2), B[0][k], B[1][k], B[2][k]);
                A[1][k] = k_dot3<KC(I, RB_F_R, 3), KC(I, RB_F_R, 4), KC(I, RB_F_R, 5)>(KV(I, RB_F_R, 3), KV(I, RB_F_R, 4), KV(I, RB_F_R, 5), B[0][k], B[1][k], B[2][k]);
                A[2][k] = k_dot3<KC(I, RB_F_R, 6), KC(I, RB_F_R, 7), KC(I, RB_F_R, 8)>(KV(I, RB_F_R, 6), KV(I, RB_F_R, 7), KV(I, RB_F_R, 8), B[0][k], B[1][k], B[2][k]);
            }
            r[0] = k_dot3_acc<KC(I, RB_F_R, 0), KC(I, RB_F_R, 1), KC(I, RB_F_R, 2)>(KV(I, RB_F_T, 0), KV(I, RB_F_R, 0), KV(I, RB_F_R, 1), KV(I, RB_F_R, 2), y[0], y[1], y[2]);
            r[1] = k_dot3_acc<KC(I, RB_F_R, 3), KC(I, RB_F_R, 4), KC(I, RB_F_R, 5)>(KV(I, RB_F_T, 1), KV(I, RB_F_R, 3), KV(I, RB_F_R, 4), KV(I, RB_F_R, 5), y[0], y[1], y[2]);
            r[2] = k_dot3_acc<KC(I, RB_F_R, 6), KC(I, RB_F_R, 7), KC(I, RB_F_R, 8)>(KV(I, RB_F_T, 2), KV(I, RB_F_R, 6), KV(I, RB_F_R, 7), KV(I, RB_F_R, 8), y[0], y[1], y[2]);
        }
    });
}
