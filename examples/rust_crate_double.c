/* rust_crate_double.c -- a C test double of rust/rigidbody_gpu_bindings.
 *
 * The Rust crate cannot be compiled in this repository's container (no cargo / rustc), so this C99 program performs
 * EXACTLY the crate's call sequence against the same C ABI, statement for statement:
 *
 *   ChainArrays::from_multibody(&mb)   ->  multibody_get_chain(): the same fields, in the same row-major order, that the
 *                                          crate reads from `for jt in mb.iter()` (multibody.rs:79-81; joint.rs:26-31;
 *                                          inertia.rs:12-17).  The reference-style handle stands in for the Rust
 *                                          `Multibody` (both are built by from_urdf, multibody.rs:65-77).
 *   ChainArrays::desc()                ->  the RbChainDesc literal below (parent = NULL, gravity (0, 0, 9.81)).
 *   GpuMultibody::new()                ->  multibody_gpu_new(&desc, device, &raw)
 *   PinnedBuffer::new()                ->  multibody_host_alloc / multibody_host_free
 *   rnea_batch / forward_dynamics_batch / rnea_fd_batch / crba_batch / fwd_kin_batch / jac_batch / rollout /
 *   rollout_cost / rnea_derivatives_batch / fd_derivatives_batch / sync / status / Drop
 *                                      ->  the multibody_*_batch calls with ld = 0, RbMem::Host, stream = NULL.
 *
 * It checks what the crate's users rely on: the flattened chain is recognised as the compiled-in FR3 model (so the
 * descriptor extraction is faithful to the last bit), batch results equal the single-state symbols of the reference ABI,
 * the one-call and two-call paths agree, the forward-dynamics round trip closes, and a NULL slice is an error, not UB.
 * Exit code 0 = all good, 1 = a check failed, 2 = no engine (no GPU: the message says so).
 *   gcc -std=c99 -Iinclude examples/rust_crate_double.c -Lrigidbody_rs_b200 -lrigidbody_b200 -lm -o rust_crate_double */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rigidbody.h"

#define N_MAX RB_MAX_JOINTS
#define CHECK(cond, what) do { if (!(cond)) { printf("FAILED: %s (%s)\n", what, multibody_last_error()); failed = 1; } } while (0)

/* rust: pub struct ChainArrays { axis, parent_rot, parent_trans, mass, com, inertia_com } */
typedef struct ChainArrays {
    int n;
    double axis[3 * N_MAX], parent_rot[9 * N_MAX], parent_trans[3 * N_MAX], mass[N_MAX], com[3 * N_MAX], inertia_com[9 * N_MAX];
} ChainArrays;

static double maxabs_diff(const double* a, const double* b, size_t n) {
    double m = 0.0;
    for (size_t i = 0; i < n; ++i) { const double d = fabs(a[i] - b[i]); if (d > m || d != d) m = d; }
    return m;
}

int main(int argc, char** argv) {
    const char* urdf = argc > 1 ? argv[1] : "assets/fr3.urdf";
    int failed = 0;
    /* let mb = Multibody::from_urdf(path) */
    Multibody* mb = multibody_new_from_urdf(urdf);
    if (!mb) { printf("load failed: %s\n", multibody_last_error()); return 2; }
    /* let arrays = ChainArrays::from_multibody(&mb); */
    static ChainArrays a;
    a.n = multibody_n_joints(mb);
    if (multibody_get_chain(mb, NULL, a.axis, a.parent_rot, a.parent_trans, a.mass, a.com, a.inertia_com) != RB_OK) return 1;
    /* let desc = arrays.desc(); */
    RbChainDesc desc;
    memset(&desc, 0, sizeof desc);
    desc.n_joints = a.n; desc.parent = NULL; desc.axis = a.axis; desc.parent_rot = a.parent_rot; desc.parent_trans = a.parent_trans;
    desc.mass = a.mass; desc.com = a.com; desc.inertia_com = a.inertia_com;
    desc.gravity[0] = 0.0; desc.gravity[1] = 0.0; desc.gravity[2] = 9.81;
    /* GpuMultibody::new(&mb, 0) */
    RbGpu* g = NULL;
    if (multibody_gpu_new(&desc, 0, &g) != RB_OK) { printf("engine: %s\n", multibody_last_error()); multibody_free(mb); return 2; }
    const int n = multibody_gpu_n_joints(g);
    printf("n = %d, kernel family: %s\n", n, multibody_gpu_kernel_variant(g));
    if (n == 7 && strstr(urdf, "fr3")) CHECK(strcmp(multibody_gpu_kernel_variant(g), "fr3-specialised") == 0, "flattened FR3 is recognised bit for bit");

    /* PinnedBuffer::new(): host batches in pinned memory, SoA [n][B] */
    const size_t B = 3001;
    double *q, *dq, *ddq, *tau, *both, *back;
    void* p[6];
    const size_t sizes[6] = {n * B, n * B, n * B, n * B, 2 * n * B, n * B};
    for (int k = 0; k < 6; ++k) if (multibody_host_alloc(&p[k], sizes[k] * sizeof(double)) != RB_OK) return 2;
    q = p[0]; dq = p[1]; ddq = p[2]; tau = p[3]; both = p[4]; back = p[5];
    unsigned long long s = 0x9E3779B97F4A7C15ULL;
    for (size_t k = 0; k < (size_t)n * B; ++k) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL; q[k] = ((double)(s >> 11) / 9007199254740992.0 - 0.5) * 4.0;
        s = s * 6364136223846793005ULL + 1442695040888963407ULL; dq[k] = ((double)(s >> 11) / 9007199254740992.0 - 0.5) * 2.0;
        s = s * 6364136223846793005ULL + 1442695040888963407ULL; ddq[k] = ((double)(s >> 11) / 9007199254740992.0 - 0.5) * 10.0;
    }
    /* gm.rnea_batch(&q, &dq, &ddq, &mut tau, B, RbLayout::Soa) ... */
    CHECK(multibody_rnea_batch(g, q, dq, ddq, tau, B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, NULL) == RB_OK, "rnea_batch");
    CHECK(multibody_forward_dynamics_batch(g, q, dq, tau, back, B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, NULL) == RB_OK, "forward_dynamics_batch");
    CHECK(multibody_rnea_fd_batch(g, q, dq, ddq, tau, both, B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, NULL) == RB_OK, "rnea_fd_batch");
    double scale = 1.0;
    for (size_t k = 0; k < (size_t)n * B; ++k) if (fabs(tau[k]) > scale) scale = fabs(tau[k]);
    CHECK(maxabs_diff(both, tau, (size_t)n * B) < 1e-12 * scale, "one call: tau equals the inverse-dynamics call to rounding");
    CHECK(maxabs_diff(both + (size_t)n * B, back, (size_t)n * B) == 0.0, "one call: qdd equals the forward-dynamics call bit for bit");
    CHECK(maxabs_diff(back, ddq, (size_t)n * B) < 1e-9, "FD(rnea(ddq)) == ddq");
    /* the single-state symbols of the reference ABI (rigidbody_bindings/src/lib.rs:15-30) against state 0 of the batch */
    if (n == 7) {
        double q1[7], dq1[7], ddq1[7];
        for (int i = 0; i < 7; ++i) { q1[i] = q[i * B]; dq1[i] = dq[i * B]; ddq1[i] = ddq[i * B]; }
        double* t1 = multibody_rnea(mb, q1, dq1, ddq1);
        CHECK(t1 != NULL, "multibody_rnea");
        if (t1) { for (int i = 0; i < 7; ++i) CHECK(t1[i] == tau[i * B], "batch[0] equals the single-state call"); multibody_free_result(t1); }
    }
    /* crba / fwd_kin / jac / derivatives: shapes and status only (values are the parity suite's business) */
    {
        const size_t Bs = 17;
        double* H = malloc(sizeof(double) * n * n * Bs); double* xyz = malloc(sizeof(double) * 3 * Bs); double* J = malloc(sizeof(double) * 6 * n * Bs);
        double* D = malloc(sizeof(double) * 3 * n * n * Bs);
        double *qa = malloc(sizeof(double) * n * Bs), *dqa = malloc(sizeof(double) * n * Bs), *xa = malloc(sizeof(double) * n * Bs);
        for (size_t st = 0; st < Bs; ++st) for (int i = 0; i < n; ++i) { qa[st * n + i] = q[i * B + st]; dqa[st * n + i] = dq[i * B + st]; xa[st * n + i] = ddq[i * B + st]; }
        CHECK(multibody_crba_batch(g, qa, H, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "crba_batch (AoS)");
        CHECK(multibody_fwd_kin_batch(g, qa, xyz, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "fwd_kin_batch");
        CHECK(multibody_jac_batch(g, qa, J, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "jac_batch");
        CHECK(multibody_rnea_derivatives_batch(g, qa, dqa, xa, D, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "rnea_derivatives_batch");
        if (n <= 12) CHECK(multibody_fd_derivatives_batch(g, qa, dqa, xa, D, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "fd_derivatives_batch");
        CHECK(H[1] == 0.0 && H[n] != 0.0, "crba keeps the reference's convention: strict lower triangle 0, column-major");
        /* rollout + fused cost over Bs trajectories, horizon 8 */
        const int Hh = 8;
        double* tt = malloc(sizeof(double) * Hh * n * Bs); double* qt = malloc(sizeof(double) * Hh * n * Bs); double* dqt = malloc(sizeof(double) * Hh * n * Bs);
        double* cost = malloc(sizeof(double) * Bs);
        for (size_t k = 0; k < (size_t)Hh * n * Bs; ++k) tt[k] = 0.1 * (double)(k % 13) - 0.6;
        CHECK(multibody_rollout(g, qa, dqa, tt, 1e-3, Hh, qt, dqt, NULL, NULL, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "rollout");
        double w[N_MAX]; for (int i = 0; i < n; ++i) w[i] = 1.0;
        RbQuadCost qc; memset(&qc, 0, sizeof qc); qc.w_q = w; qc.w_tau = w;
        CHECK(multibody_rollout_cost(g, qa, dqa, tt, 1e-3, Hh, &qc, cost, NULL, NULL, Bs, 0, RB_LAYOUT_AOS, RB_MEM_HOST, NULL) == RB_OK, "rollout_cost");
        CHECK(cost[0] > 0.0 && cost[0] == cost[0], "rollout_cost returns a finite positive cost");
        free(H); free(xyz); free(J); free(D); free(qa); free(dqa); free(xa); free(tt); free(qt); free(dqt); free(cost);
    }
    /* errors are statuses with a message, never UB or an abort (the reference unwraps / panics, lib.rs:22) */
    CHECK(multibody_rnea_batch(g, NULL, dq, ddq, tau, B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, NULL) == RB_ERR_NULL, "NULL slice -> RB_ERR_NULL");
    CHECK(multibody_rnea_batch(g, q, dq, ddq, tau, B, B - 1, RB_LAYOUT_SOA, RB_MEM_HOST, NULL) == RB_ERR_ARG, "ld < n_states -> RB_ERR_ARG");
    CHECK(multibody_gpu_sync(g) == RB_OK && multibody_gpu_status(g) == RB_OK, "sync / status");
    CHECK(multibody_gpu_launch_count(g) > 0, "launch_count");
    /* impl Drop */
    for (int k = 0; k < 6; ++k) multibody_host_free(p[k]);
    multibody_gpu_free(g);
    multibody_free(mb);
    printf(failed ? "rust_crate_double: FAILED\n" : "rust_crate_double: all checks passed\n");
    return failed;
}
