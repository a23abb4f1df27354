/*
 * rb_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the hot path of khaninger/rigidbody-rs, written in the
 * *shape of the reference*: unit-quaternion isometries (nalgebra Isometry3<f64>),
 * one state at a time, the same operation order as the cited Rust lines.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product (rigidbody_rs_b200/csrc) never does.
 *
 * PARITY UNPINNED: the reference cannot be compiled in this image (no cargo/rustc,
 * nightly-only crate, un-vendored nalgebra 0.33.2 / xurdf 0.2.5) and its own tests
 * pin no rnea/crba value.  What pins this oracle instead:
 *   - the reference's convention tests spatial.rs:283-382 restated in tests/;
 *   - an independent numpy restatement in matrix form (oracle/rb_oracle_np.py);
 *   - identities: sym(crba) ddq + rnea(q,dq,0) == rnea(q,dq,ddq), FD round trip,
 *     H symmetric positive definite, gravity torque == dU/dq, fwd_kin at q=0;
 *   - the SURVEY.md section 8c known-answer vectors (tests/golden/).
 * nalgebra 0.33.2 arithmetic (Cargo.lock:205-206) is restated from its published
 * algorithm; each function names the nalgebra entry point it mirrors.
 *
 * All file:line citations are relative to the reference checkout.
 */
#ifndef RB_ORACLE_H
#define RB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBO_MAX_N 64

/* nalgebra Quaternion coords are [i, j, k, w]; we keep named fields. */
typedef struct { double i, j, k, w; } rbo_quat;
/* rigidbody/src/lib.rs:16  Transform = Isometry3<f64> = rotation + translation */
typedef struct { rbo_quat rot; double t[3]; } rbo_iso;
/* rigidbody/src/spatial.rs:87-91 and :187-191 (same shape for velocity and force) */
typedef struct { double lin[3]; double rot[3]; } rbo_sv;
/* rigidbody/src/inertia.rs:12-18; 3x3 stored row-major */
typedef struct { double mass; double com[3]; double inertia_com[9]; double inertia[9]; } rbo_inertia;
/* rigidbody/src/joint.rs:26-31 */
typedef struct { double axis[3]; rbo_iso parent; rbo_inertia body; } rbo_joint;
/* rigidbody/src/multibody.rs:32, generalised from [RevoluteJoint;7] to n <= RBO_MAX_N
 * (serial chain, exactly as the reference: link i's parent is link i-1). */
typedef struct { int n; rbo_joint jt[RBO_MAX_N]; } rbo_multibody;

/* joint.rs:53-68 from_xurdf_joint for every surviving (joint, link) pair.
 * axis[3n], xyz[3n], rpy[3n], mass[n], com[3n], inertia6[6n] = ixx ixy ixz iyy iyz izz. */
int rbo_multibody_init(rbo_multibody* mb, int n, const double* axis, const double* xyz,
                       const double* rpy, const double* mass, const double* com,
                       const double* inertia6);

/* multibody.rs:41-49 / :83-85 */
void rbo_get_transforms(const rbo_multibody* mb, const double* q, rbo_iso* tr);
/* multibody.rs:111-153 */
void rbo_rnea_tr(const rbo_multibody* mb, const rbo_iso* tr, const double* dq, const double* ddq, double* tau);
/* multibody.rs:155-174; H is n x n column-major, identity-initialised, upper triangle written */
void rbo_crba_tr(const rbo_multibody* mb, const rbo_iso* tr, double* H);
/* multibody.rs:87-93; returns the base->tip isometry */
void rbo_fwd_kin_tr(const rbo_multibody* mb, const rbo_iso* tr, rbo_iso* out);
/* multibody.rs:95-108; J is 6 x n column-major, rows 0-2 lin, 3-5 rot */
void rbo_jac_tr(const rbo_multibody* mb, const rbo_iso* tr, double* J);

/* The FFI-shaped single-state calls (rigidbody_bindings/src/lib.rs:15-70): get_transforms then the op. */
void rbo_rnea(const rbo_multibody* mb, const double* q, const double* dq, const double* ddq, double* tau);
void rbo_crba(const rbo_multibody* mb, const double* q, double* H);
void rbo_fwd_kin(const rbo_multibody* mb, const double* q, double* xyz3);
void rbo_jac(const rbo_multibody* mb, const double* q, double* J);

/* Forward dynamics -- NOT in the reference: qdd = chol_solve(sym(crba(q)), tau - rnea(q,dq,0)).
 * Returns 0, or -1 if H is not positive definite. */
int rbo_forward_dynamics(const rbo_multibody* mb, const double* q, const double* dq, const double* tau, double* qdd);

/* Semi-implicit Euler rollout of one trajectory: per step qdd = FD(q,dq,tau_t);
 * dq += dt*qdd; q += dt*dq.  tau is [H][n]; q_traj/dq_traj are [H][n] (state after each step). */
int rbo_rollout(const rbo_multibody* mb, const double* q0, const double* dq0, const double* tau,
                double dt, int horizon, double* q_traj, double* dq_traj);

/* ---- batch drivers (OpenMP static chunks over states: the rayon par_chunks stand-in) ----
 * Layout: SoA joint-major [n][B] when soa != 0, AoS [B][n] otherwise. */
void rbo_rnea_batch(const rbo_multibody* mb, const double* q, const double* dq, const double* ddq,
                    double* tau, size_t B, int soa, int threads);
int rbo_forward_dynamics_batch(const rbo_multibody* mb, const double* q, const double* dq, const double* tau,
                               double* qdd, size_t B, int soa, int threads);
void rbo_crba_batch(const rbo_multibody* mb, const double* q, double* H, size_t B, int soa, int threads);
int rbo_max_threads(void);

/* ---- counter-based sampler shared (by definition, not by code) with the CUDA generator ----
 * u = splitmix64 finaliser of (seed + GOLDEN*(1 + (field<<58 | joint<<50 | index))), top 53 bits;
 * value = fma(hi-lo, u, lo). */
double rbo_sample(uint64_t seed, unsigned field, unsigned joint, uint64_t index, double lo, double hi);
void rbo_fill(double* out, uint64_t seed, unsigned field, int n, const double* lo, const double* hi,
              size_t first, size_t count, size_t ld, int soa);

/* ---- exposed building blocks so tests can restate spatial.rs:283-382 ---- */
rbo_quat rbo_quat_from_scaled_axis(const double v[3]);
rbo_quat rbo_quat_mul(rbo_quat a, rbo_quat b);
void rbo_quat_rotate(rbo_quat q, const double v[3], double out[3]);
void rbo_quat_to_matrix(rbo_quat q, double R[9]);
rbo_iso rbo_iso_inverse(rbo_iso a);
rbo_iso rbo_iso_mul(rbo_iso a, rbo_iso b);
rbo_sv rbo_motion_transform(const rbo_sv* v, const rbo_iso* tr);
rbo_sv rbo_force_transform(const rbo_sv* f, const rbo_iso* tr);
rbo_sv rbo_cross_star(const rbo_sv* v, const rbo_sv* f);
rbo_sv rbo_inertia_mul(const rbo_inertia* I, const rbo_sv* a);
rbo_inertia rbo_inertia_from_com(double mass, const double com[3], const double inertia_com[9]);
rbo_inertia rbo_inertia_transform(const rbo_inertia* I, const rbo_iso* tr);
rbo_inertia rbo_inertia_add(const rbo_inertia* a, const rbo_inertia* b);
void rbo_plucker_motion(const rbo_iso* T, double X[36]);      /* spatial.rs:32-47  transform_to_B_X_A */
void rbo_plucker_force(const rbo_iso* T, double X[36]);       /* spatial.rs:53-66  transform_to_B_X_A_star */

#ifdef __cplusplus
}
#endif
#endif
