/*
 * rb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See rb_oracle.h.
 *
 * Reference-shaped restatement: quaternion isometries, one state at a time, same
 * operation order as the cited Rust lines.  PARITY UNPINNED (see header).
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps the arithmetic as written (rustc does not contract to FMA).
 */
#include "rb_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ 3-vectors */
static inline void v3_cross(const double a[3], const double b[3], double o[3]) {
    /* nalgebra Vector3::cross */
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void v3_sub(const double a[3], const double b[3], double o[3]) {
    o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
static inline void v3_add(const double a[3], const double b[3], double o[3]) {
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
}
/* nalgebra Vector3::cross_matrix, row-major */
static inline void v3_cross_matrix(const double v[3], double M[9]) {
    M[0] = 0.0;   M[1] = -v[2]; M[2] = v[1];
    M[3] = v[2];  M[4] = 0.0;   M[5] = -v[0];
    M[6] = -v[1]; M[7] = v[0];  M[8] = 0.0;
}
static inline void m3_mul(const double A[9], const double B[9], double C[9]) {
    double T[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            T[3 * r + c] = A[3 * r + 0] * B[0 + c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
    memcpy(C, T, sizeof T);
}
static inline void m3_transpose(const double A[9], double T[9]) {
    double B[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(T, B, sizeof B);
}
static inline void m3_mulv(const double A[9], const double v[3], double o[3]) {
    double x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    double y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    double z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}

/* ------------------------------------------------------------------ quaternions (nalgebra) */
/* UnitQuaternion::from_scaled_axis(v) = Quaternion::from_imag(v/2).exp()  (nalgebra 0.33
 * geometry/quaternion_construction.rs; Quaternion::exp_eps with eps = f64::EPSILON):
 * identity when |v/2|^2 <= eps^2, else (cos n, sin(n)/n * v/2) with n = |v/2|.
 * Call site: joint.rs:48-50 (joint_transform) and Isometry3::new at joint.rs:57-64. */
rbo_quat rbo_quat_from_scaled_axis(const double v[3]) {
    double h[3] = {v[0] / 2.0, v[1] / 2.0, v[2] / 2.0};
    double nn = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    const double eps = 2.220446049250313e-16;
    rbo_quat q;
    if (nn <= eps * eps) {
        q.i = 0.0; q.j = 0.0; q.k = 0.0; q.w = 1.0;
        return q;
    }
    double n = sqrt(nn);
    double s = sin(n) / n;
    q.i = h[0] * s; q.j = h[1] * s; q.k = h[2] * s; q.w = cos(n);
    return q;
}

/* Quaternion * Quaternion (Hamilton product), nalgebra geometry/quaternion_ops.rs */
rbo_quat rbo_quat_mul(rbo_quat a, rbo_quat b) {
    rbo_quat r;
    r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
    r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
    r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
    r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
    return r;
}
static inline rbo_quat quat_conj(rbo_quat q) {
    /* UnitQuaternion::inverse == conjugate */
    rbo_quat r = {-q.i, -q.j, -q.k, q.w};
    return r;
}
/* UnitQuaternion * Vector3: t = 2 (qv x v); v' = t*w + qv x t + v */
void rbo_quat_rotate(rbo_quat q, const double v[3], double out[3]) {
    double qv[3] = {q.i, q.j, q.k};
    double t[3], c[3];
    v3_cross(qv, v, t);
    t[0] *= 2.0; t[1] *= 2.0; t[2] *= 2.0;
    v3_cross(qv, t, c);
    out[0] = t[0] * q.w + c[0] + v[0];
    out[1] = t[1] * q.w + c[1] + v[1];
    out[2] = t[2] * q.w + c[2] + v[2];
}
/* UnitQuaternion::to_rotation_matrix (row-major out) */
void rbo_quat_to_matrix(rbo_quat q, double R[9]) {
    double i = q.i, j = q.j, k = q.k, w = q.w;
    double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    double ij = i * j * 2.0, wk = w * k * 2.0, wj = w * j * 2.0;
    double ik = i * k * 2.0, jk = j * k * 2.0, wi = w * i * 2.0;
    R[0] = ww + ii - jj - kk; R[1] = ij - wk;           R[2] = wj + ik;
    R[3] = wk + ij;           R[4] = ww - ii + jj - kk; R[5] = jk - wi;
    R[6] = ik - wj;           R[7] = wi + jk;           R[8] = ww - ii - jj + kk;
}

/* ------------------------------------------------------------------ isometries (nalgebra) */
/* Isometry3::inverse: rot' = rot^-1, t' = rot' * (-t) */
rbo_iso rbo_iso_inverse(rbo_iso a) {
    rbo_iso r;
    r.rot = quat_conj(a.rot);
    double nt[3] = {-a.t[0], -a.t[1], -a.t[2]};
    rbo_quat_rotate(r.rot, nt, r.t);
    return r;
}
/* Isometry3 * Isometry3: t = t1 + R1 t2, R = R1 R2 */
rbo_iso rbo_iso_mul(rbo_iso a, rbo_iso b) {
    rbo_iso r;
    double shift[3];
    rbo_quat_rotate(a.rot, b.t, shift);
    v3_add(a.t, shift, r.t);
    r.rot = rbo_quat_mul(a.rot, b.rot);
    return r;
}
static inline rbo_iso iso_identity(void) {
    rbo_iso r = {{0.0, 0.0, 0.0, 1.0}, {0.0, 0.0, 0.0}};
    return r;
}

/* ------------------------------------------------------------------ spatial.rs */
/* spatial.rs:110-116  SpatialVelocity::transform (called through Transform * &SpatialVelocity, :20-26) */
rbo_sv rbo_motion_transform(const rbo_sv* v, const rbo_iso* tr) {
    rbo_quat rot = quat_conj(tr->rot);
    double c[3], d[3];
    rbo_sv o;
    v3_cross(tr->t, v->rot, c);
    v3_sub(v->lin, c, d);
    rbo_quat_rotate(rot, d, o.lin);
    rbo_quat_rotate(rot, v->rot, o.rot);
    return o;
}
/* spatial.rs:242-248  SpatialForce::transform (through Transform * &SpatialForce, :211-218) */
rbo_sv rbo_force_transform(const rbo_sv* f, const rbo_iso* tr) {
    rbo_quat rot = quat_conj(tr->rot);
    double c[3], d[3];
    rbo_sv o;
    rbo_quat_rotate(rot, f->lin, o.lin);
    v3_cross(tr->t, f->lin, c);
    v3_sub(f->rot, c, d);
    rbo_quat_rotate(rot, d, o.rot);
    return o;
}
/* spatial.rs:129-134 */
rbo_sv rbo_cross_star(const rbo_sv* v, const rbo_sv* f) {
    rbo_sv o;
    double a[3], b[3];
    v3_cross(v->rot, f->lin, o.lin);
    v3_cross(v->rot, f->rot, a);
    v3_cross(v->lin, f->lin, b);
    v3_add(a, b, o.rot);
    return o;
}
/* spatial.rs:194-201 / :203-209 */
static inline rbo_sv sv_add(const rbo_sv* a, const rbo_sv* b) {
    rbo_sv o;
    v3_add(a->lin, b->lin, o.lin);
    v3_add(a->rot, b->rot, o.rot);
    return o;
}

/* spatial.rs:32-47  6x6 [rot;lin]-ordered motion matrix (tests only), row-major */
void rbo_plucker_motion(const rbo_iso* T, double X[36]) {
    double E[9], rc[9], Er[9];
    rbo_quat_to_matrix(T->rot, E);
    v3_cross_matrix(T->t, rc);
    m3_mul(E, rc, Er);
    memset(X, 0, 36 * sizeof(double));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            X[6 * r + c] = E[3 * r + c];
            X[6 * (r + 3) + c] = -Er[3 * r + c];
            X[6 * (r + 3) + c + 3] = E[3 * r + c];
        }
}
/* spatial.rs:53-66 */
void rbo_plucker_force(const rbo_iso* T, double X[36]) {
    double E[9], rc[9], Er[9];
    rbo_quat_to_matrix(T->rot, E);
    v3_cross_matrix(T->t, rc);
    m3_mul(E, rc, Er);
    memset(X, 0, 36 * sizeof(double));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            X[6 * r + c] = E[3 * r + c];
            X[6 * r + c + 3] = -Er[3 * r + c];
            X[6 * (r + 3) + c + 3] = E[3 * r + c];
        }
}

/* ------------------------------------------------------------------ inertia.rs */
/* inertia.rs:21-35 */
rbo_inertia rbo_inertia_from_com(double mass, const double com[3], const double inertia_com[9]) {
    rbo_inertia I;
    double C[9], mC[9], Ct[9], P[9];
    I.mass = mass;
    memcpy(I.com, com, 3 * sizeof(double));
    memcpy(I.inertia_com, inertia_com, 9 * sizeof(double));
    v3_cross_matrix(com, C);
    for (int k = 0; k < 9; ++k) mC[k] = mass * C[k];
    m3_transpose(C, Ct);
    m3_mul(mC, Ct, P);
    for (int k = 0; k < 9; ++k) I.inertia[k] = inertia_com[k] + P[k];
    return I;
}
/* inertia.rs:37-51 */
static rbo_inertia inertia_from_origin(double mass, const double com[3], const double inertia[9]) {
    rbo_inertia I;
    double C[9], mC[9], Ct[9], P[9];
    I.mass = mass;
    memcpy(I.com, com, 3 * sizeof(double));
    memcpy(I.inertia, inertia, 9 * sizeof(double));
    v3_cross_matrix(com, C);
    for (int k = 0; k < 9; ++k) mC[k] = mass * C[k];
    m3_transpose(C, Ct);
    m3_mul(mC, Ct, P);
    for (int k = 0; k < 9; ++k) I.inertia_com[k] = inertia[k] - P[k];
    return I;
}
/* inertia.rs:81-89 */
rbo_inertia rbo_inertia_transform(const rbo_inertia* I, const rbo_iso* tr) {
    double R[9], Rt[9], RI[9], RIRt[9], rc[3], com[3];
    rbo_quat_to_matrix(tr->rot, R);
    rbo_quat_rotate(tr->rot, I->com, rc);   /* Isometry3 * Point3 = R p + t */
    v3_add(rc, tr->t, com);
    m3_mul(R, I->inertia_com, RI);
    m3_transpose(R, Rt);
    m3_mul(RI, Rt, RIRt);
    return rbo_inertia_from_com(I->mass, com, RIRt);
}
/* inertia.rs:96-105 */
rbo_inertia rbo_inertia_add(const rbo_inertia* a, const rbo_inertia* b) {
    double com[3], S[9];
    double msum = a->mass + b->mass;
    for (int k = 0; k < 3; ++k) com[k] = (a->mass * a->com[k] + b->mass * b->com[k]) / msum;
    for (int k = 0; k < 9; ++k) S[k] = a->inertia[k] + b->inertia[k];
    return inertia_from_origin(msum, com, S);
}
/* inertia.rs:107-117 */
rbo_sv rbo_inertia_mul(const rbo_inertia* I, const rbo_sv* a) {
    rbo_sv f;
    double cr[3], cl[3], Iw[3];
    v3_cross(I->com, a->rot, cr);
    v3_cross(I->com, a->lin, cl);
    m3_mulv(I->inertia, a->rot, Iw);
    for (int k = 0; k < 3; ++k) {
        f.lin[k] = I->mass * a->lin[k] - I->mass * cr[k];
        f.rot[k] = Iw[k] + I->mass * cl[k];
    }
    return f;
}

/* ------------------------------------------------------------------ joint.rs */
/* Rotation3::from_euler_angles(roll, pitch, yaw) = Rz(yaw) Ry(pitch) Rx(roll), nalgebra
 * geometry/rotation_specialization.rs; then Rotation3::scaled_axis() = axis() * angle():
 * angle = acos(clamp((trace-1)/2)), axis = normalize((R21-R12, R02-R20, R10-R01)) or None. */
static void euler_to_scaled_axis(double r, double p, double y, double out[3]) {
    double sr = sin(r), cr = cos(r), sp = sin(p), cp = cos(p), sy = sin(y), cy = cos(y);
    double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                   sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                   -sp,     cp * sr,                cp * cr};
    double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double nrm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const double eps = 2.220446049250313e-16;
    if (!(nrm > eps)) { out[0] = out[1] = out[2] = 0.0; return; }
    double c = (R[0] + R[4] + R[8] - 1.0) / 2.0;
    if (c < -1.0) c = -1.0;
    if (c > 1.0) c = 1.0;
    double ang = acos(c);
    out[0] = ax[0] / nrm * ang; out[1] = ax[1] / nrm * ang; out[2] = ax[2] / nrm * ang;
}

/* joint.rs:53-68 */
int rbo_multibody_init(rbo_multibody* mb, int n, const double* axis, const double* xyz,
                       const double* rpy, const double* mass, const double* com,
                       const double* inertia6) {
    if (!mb || n < 1 || n > RBO_MAX_N) return -1;
    mb->n = n;
    for (int i = 0; i < n; ++i) {
        rbo_joint* jt = &mb->jt[i];
        const double* a = axis + 3 * i;
        double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);   /* UnitVector3::new_normalize :56 */
        jt->axis[0] = a[0] / an; jt->axis[1] = a[1] / an; jt->axis[2] = a[2] / an;
        double sa[3];
        euler_to_scaled_axis(rpy[3 * i], rpy[3 * i + 1], rpy[3 * i + 2], sa);   /* :59-63 */
        jt->parent.rot = rbo_quat_from_scaled_axis(sa);                        /* Isometry3::new :57 */
        memcpy(jt->parent.t, xyz + 3 * i, 3 * sizeof(double));
        const double* s = inertia6 + 6 * i;
        double Ic[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
        jt->body = rbo_inertia_from_com(mass[i], com + 3 * i, Ic);             /* :66 */
    }
    return 0;
}

/* joint.rs:48-50 joint_transform, :36-38 parent_to_child; multibody.rs:41-49 */
void rbo_get_transforms(const rbo_multibody* mb, const double* q, rbo_iso* tr) {
    for (int i = 0; i < mb->n; ++i) {
        const rbo_joint* jt = &mb->jt[i];
        double sa[3] = {jt->axis[0] * q[i], jt->axis[1] * q[i], jt->axis[2] * q[i]};
        rbo_quat jq = rbo_quat_from_scaled_axis(sa);
        tr[i].rot = rbo_quat_mul(jt->parent.rot, jq);   /* Isometry3 * UnitQuaternion */
        memcpy(tr[i].t, jt->parent.t, 3 * sizeof(double));
    }
}

/* ------------------------------------------------------------------ multibody.rs */
/* multibody.rs:111-153 */
void rbo_rnea_tr(const rbo_multibody* mb, const rbo_iso* tr, const double* dq, const double* ddq, double* tau) {
    const int n = mb->n;
    rbo_sv f[RBO_MAX_N];
    rbo_sv v = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};           /* :116 */
    rbo_sv a = {{0.0, 0.0, 9.81}, {0.0, 0.0, 0.0}};          /* :117-120 */
    for (int i = 0; i < n; ++i) {
        const rbo_joint* jt = &mb->jt[i];
        v = rbo_motion_transform(&v, &tr[i]);                /* :129 */
        v.rot[2] += dq[i];                                   /* :130 */
        a = rbo_motion_transform(&a, &tr[i]);                /* :132 */
        a.rot[2] += ddq[i];                                  /* :133 */
        a.lin[0] += v.lin[1] * dq[i];                        /* :135-138 */
        a.lin[1] += -v.lin[0] * dq[i];
        a.rot[0] += v.rot[1] * dq[i];
        a.rot[1] += -v.rot[0] * dq[i];
        rbo_sv Ia = rbo_inertia_mul(&jt->body, &a);          /* :140 */
        rbo_sv Iv = rbo_inertia_mul(&jt->body, &v);
        rbo_sv vx = rbo_cross_star(&v, &Iv);
        f[i] = sv_add(&Ia, &vx);
    }
    for (int i = n - 1; i >= 0; --i) {
        tau[i] = f[i].rot[2];                                /* :144 */
        if (i > 0) {
            rbo_iso inv = rbo_iso_inverse(tr[i]);            /* :147 */
            rbo_sv ft = rbo_force_transform(&f[i], &inv);
            f[i - 1] = sv_add(&f[i - 1], &ft);               /* :148 */
        }
    }
}

/* multibody.rs:155-174 */
void rbo_crba_tr(const rbo_multibody* mb, const rbo_iso* tr, double* H) {
    const int n = mb->n;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) H[r + (size_t)n * c] = (r == c) ? 1.0 : 0.0;   /* :156 */
    rbo_inertia I = mb->jt[n - 1].body;                                              /* :157 */
    const rbo_sv S = {{0.0, 0.0, 0.0}, {0.0, 0.0, 1.0}};                             /* multibody.rs:29 */
    for (int i = n - 1; i >= 0; --i) {
        H[i + (size_t)n * i] = I.inertia[8];                                          /* :161 get_rotz */
        rbo_sv F = rbo_inertia_mul(&I, &S);                                           /* :162 */
        for (int j = i - 1; j >= 0; --j) {
            rbo_iso inv = rbo_iso_inverse(tr[j + 1]);                                 /* :165 */
            F = rbo_force_transform(&F, &inv);
            H[j + (size_t)n * i] = F.rot[2];                                          /* :166 */
        }
        if (i > 0) {
            rbo_inertia It = rbo_inertia_transform(&I, &tr[i]);                       /* :170 */
            I = rbo_inertia_add(&mb->jt[i - 1].body, &It);
        }
    }
}

/* multibody.rs:87-93 */
void rbo_fwd_kin_tr(const rbo_multibody* mb, const rbo_iso* tr, rbo_iso* out) {
    rbo_iso acc = iso_identity();
    for (int i = mb->n - 1; i >= 0; --i) acc = rbo_iso_mul(tr[i], acc);
    *out = acc;
}

/* multibody.rs:95-108 */
void rbo_jac_tr(const rbo_multibody* mb, const rbo_iso* tr, double* J) {
    const int n = mb->n;
    const rbo_sv S = {{0.0, 0.0, 0.0}, {0.0, 0.0, 1.0}};
    rbo_iso acc = iso_identity();
    for (int i = n - 1; i >= 0; --i) {
        rbo_sv vi = rbo_motion_transform(&S, &acc);          /* :100 */
        for (int k = 0; k < 3; ++k) {
            J[6 * i + k] = vi.lin[k];                        /* :102 */
            J[6 * i + 3 + k] = vi.rot[k];                    /* :103 */
        }
        acc = rbo_iso_mul(tr[i], acc);                       /* :105 */
    }
}

/* ------------------------------------------------------------------ rigidbody_bindings/src/lib.rs shape */
void rbo_rnea(const rbo_multibody* mb, const double* q, const double* dq, const double* ddq, double* tau) {
    rbo_iso tr[RBO_MAX_N];
    rbo_get_transforms(mb, q, tr);          /* lib.rs:24 */
    rbo_rnea_tr(mb, tr, dq, ddq, tau);      /* lib.rs:28 */
}
void rbo_crba(const rbo_multibody* mb, const double* q, double* H) {
    rbo_iso tr[RBO_MAX_N];
    rbo_get_transforms(mb, q, tr);          /* lib.rs:39 */
    rbo_crba_tr(mb, tr, H);                 /* lib.rs:41 */
}
void rbo_fwd_kin(const rbo_multibody* mb, const double* q, double* xyz3) {
    rbo_iso tr[RBO_MAX_N], out;
    rbo_get_transforms(mb, q, tr);          /* lib.rs:52 */
    rbo_fwd_kin_tr(mb, tr, &out);
    memcpy(xyz3, out.t, 3 * sizeof(double));   /* lib.rs:55 */
}
void rbo_jac(const rbo_multibody* mb, const double* q, double* J) {
    rbo_iso tr[RBO_MAX_N];
    rbo_get_transforms(mb, q, tr);          /* lib.rs:66 */
    rbo_jac_tr(mb, tr, J);
}

/* ------------------------------------------------------------------ forward dynamics (new) */
int rbo_forward_dynamics(const rbo_multibody* mb, const double* q, const double* dq, const double* tau, double* qdd) {
    const int n = mb->n;
    rbo_iso tr[RBO_MAX_N];
    double zero[RBO_MAX_N] = {0.0}, c[RBO_MAX_N], y[RBO_MAX_N];
    static _Thread_local double H[RBO_MAX_N * RBO_MAX_N], L[RBO_MAX_N * RBO_MAX_N];
    rbo_get_transforms(mb, q, tr);
    rbo_rnea_tr(mb, tr, dq, zero, c);
    rbo_crba_tr(mb, tr, H);
    /* symmetrise: the reference leaves the strict lower triangle at 0 (multibody.rs:156) */
    for (int cidx = 0; cidx < n; ++cidx)
        for (int r = cidx + 1; r < n; ++r) H[r + (size_t)n * cidx] = H[cidx + (size_t)n * r];
    /* Cholesky H = L L^T (lower), column by column */
    for (int j = 0; j < n; ++j) {
        double d = H[j + (size_t)n * j];
        for (int k = 0; k < j; ++k) d -= L[j + (size_t)n * k] * L[j + (size_t)n * k];
        if (!(d > 0.0)) return -1;
        double ljj = sqrt(d);
        L[j + (size_t)n * j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double s = H[i + (size_t)n * j];
            for (int k = 0; k < j; ++k) s -= L[i + (size_t)n * k] * L[j + (size_t)n * k];
            L[i + (size_t)n * j] = s / ljj;
        }
    }
    for (int i = 0; i < n; ++i) {           /* L y = tau - c */
        double s = tau[i] - c[i];
        for (int k = 0; k < i; ++k) s -= L[i + (size_t)n * k] * y[k];
        y[i] = s / L[i + (size_t)n * i];
    }
    for (int i = n - 1; i >= 0; --i) {      /* L^T qdd = y */
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[k + (size_t)n * i] * qdd[k];
        qdd[i] = s / L[i + (size_t)n * i];
    }
    return 0;
}

int rbo_rollout(const rbo_multibody* mb, const double* q0, const double* dq0, const double* tau,
                double dt, int horizon, double* q_traj, double* dq_traj) {
    const int n = mb->n;
    double q[RBO_MAX_N], dq[RBO_MAX_N], qdd[RBO_MAX_N];
    memcpy(q, q0, n * sizeof(double));
    memcpy(dq, dq0, n * sizeof(double));
    for (int t = 0; t < horizon; ++t) {
        if (rbo_forward_dynamics(mb, q, dq, tau + (size_t)t * n, qdd)) return -1;
        for (int i = 0; i < n; ++i) {
            dq[i] = dq[i] + dt * qdd[i];
            q[i] = q[i] + dt * dq[i];
        }
        memcpy(q_traj + (size_t)t * n, q, n * sizeof(double));
        memcpy(dq_traj + (size_t)t * n, dq, n * sizeof(double));
    }
    return 0;
}

/* ------------------------------------------------------------------ batch drivers */
int rbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline void gather(const double* x, size_t B, size_t s, int n, int soa, double* o) {
    if (soa) for (int i = 0; i < n; ++i) o[i] = x[(size_t)i * B + s];
    else     for (int i = 0; i < n; ++i) o[i] = x[s * (size_t)n + i];
}
static inline void scatter(double* x, size_t B, size_t s, int n, int soa, const double* o) {
    if (soa) for (int i = 0; i < n; ++i) x[(size_t)i * B + s] = o[i];
    else     for (int i = 0; i < n; ++i) x[s * (size_t)n + i] = o[i];
}

void rbo_rnea_batch(const rbo_multibody* mb, const double* q, const double* dq, const double* ddq,
                    double* tau, size_t B, int soa, int threads) {
    const int n = mb->n;
    if (threads < 1) threads = rbo_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long s = 0; s < (long long)B; ++s) {
        double a[RBO_MAX_N], b[RBO_MAX_N], c[RBO_MAX_N], t[RBO_MAX_N];
        gather(q, B, (size_t)s, n, soa, a);
        gather(dq, B, (size_t)s, n, soa, b);
        gather(ddq, B, (size_t)s, n, soa, c);
        rbo_rnea(mb, a, b, c, t);
        scatter(tau, B, (size_t)s, n, soa, t);
    }
}

int rbo_forward_dynamics_batch(const rbo_multibody* mb, const double* q, const double* dq, const double* tau,
                               double* qdd, size_t B, int soa, int threads) {
    const int n = mb->n;
    int bad = 0;
    if (threads < 1) threads = rbo_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads) reduction(| : bad)
    for (long long s = 0; s < (long long)B; ++s) {
        double a[RBO_MAX_N], b[RBO_MAX_N], c[RBO_MAX_N], t[RBO_MAX_N];
        gather(q, B, (size_t)s, n, soa, a);
        gather(dq, B, (size_t)s, n, soa, b);
        gather(tau, B, (size_t)s, n, soa, c);
        if (rbo_forward_dynamics(mb, a, b, c, t)) bad |= 1;
        scatter(qdd, B, (size_t)s, n, soa, t);
    }
    return bad ? -1 : 0;
}

/* H out: SoA -> [n*n][B] with row index k = r + n*c (column-major entry index); AoS -> [B][n*n] column-major */
void rbo_crba_batch(const rbo_multibody* mb, const double* q, double* H, size_t B, int soa, int threads) {
    const int n = mb->n;
    if (threads < 1) threads = rbo_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long s = 0; s < (long long)B; ++s) {
        double a[RBO_MAX_N];
        double Hs[RBO_MAX_N * RBO_MAX_N];
        gather(q, B, (size_t)s, n, soa, a);
        rbo_crba(mb, a, Hs);
        if (soa) for (int k = 0; k < n * n; ++k) H[(size_t)k * B + (size_t)s] = Hs[k];
        else     memcpy(H + (size_t)s * n * n, Hs, (size_t)n * n * sizeof(double));
    }
}

/* ------------------------------------------------------------------ sampler */
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
double rbo_sample(uint64_t seed, unsigned field, unsigned joint, uint64_t index, double lo, double hi) {
    uint64_t ctr = ((uint64_t)field << 58) | ((uint64_t)joint << 50) | (index & ((1ULL << 50) - 1));
    uint64_t z = mix64(seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL));
    double u = (double)(z >> 11) * 0x1.0p-53;
    return fma(hi - lo, u, lo);
}
void rbo_fill(double* out, uint64_t seed, unsigned field, int n, const double* lo, const double* hi,
              size_t first, size_t count, size_t ld, int soa) {
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < (long long)count; ++s)
        for (int i = 0; i < n; ++i) {
            double v = rbo_sample(seed, field, (unsigned)i, (uint64_t)(first + (size_t)s), lo[i], hi[i]);
            if (soa) out[(size_t)i * ld + (size_t)s] = v;
            else     out[(size_t)s * n + i] = v;
        }
}
