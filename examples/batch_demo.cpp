// batch_demo.cpp -- a C++ caller of the C ABI (include/rigidbody.h), the way rigidbody_bindings/main.cpp calls the
// reference's (main.cpp:66-98): load the FR3, evaluate the single-state reference symbols, then a batch through the new
// entry points, and check that entry 0 of the batch equals the single-state result.
//
//   g++ -std=c++17 -Iinclude examples/batch_demo.cpp -Lrigidbody_rs_b200 -lrigidbody_b200 -Wl,-rpath,$PWD/rigidbody_rs_b200 -o batch_demo
//   ./batch_demo assets/fr3.urdf
#include <cmath>
#include <cstdio>
#include <vector>

#include "rigidbody.h"

int main(int argc, char** argv) {
    const char* urdf = argc > 1 ? argv[1] : "assets/fr3.urdf";
    // ---- the reference's own surface (Part 1): one state, malloc'ed results
    Multibody* mb = multibody_new_from_urdf(urdf);
    if (!mb) { std::printf("load failed: %s\n", multibody_last_error()); return 2; }
    const int n = multibody_n_joints(mb);
    const double q1[7] = {0, 0, 1, 0, 1, 0, 0}, dq1[7] = {0, 0, 0, 0, 1, 0, 0}, ddq1[7] = {1, 0, 0, 0, 0, 1, 0};   // main.cpp:103-105
    double* tau1 = multibody_rnea(mb, q1, dq1, ddq1);
    if (!tau1) { std::printf("rnea failed: %s\n", multibody_last_error()); return 2; }
    std::printf("n = %d   tau(single) =", n);
    for (int i = 0; i < n; ++i) std::printf(" %.12g", tau1[i]);
    std::printf("\n");

    // ---- the batched engine (Part 2): caller-allocated, status-returning
    RbGpu* g = nullptr;
    if (multibody_gpu_new_from_urdf(urdf, 0, &g) != RB_OK) { std::printf("engine: %s\n", multibody_last_error()); return 2; }
    std::printf("kernel family: %s\n", multibody_gpu_kernel_variant(g));
    const size_t B = 1000;
    std::vector<double> q(n * B), dq(n * B), ddq(n * B), tau(n * B), qdd(n * B), out(2 * n * B);
    for (size_t s = 0; s < B; ++s)
        for (int i = 0; i < n; ++i) {                      // joint-major SoA: x[i * B + s]; state 0 = the single state above
            q[i * B + s] = q1[i] + 1e-3 * s; dq[i * B + s] = dq1[i]; ddq[i * B + s] = ddq1[i];
        }
    if (multibody_rnea_batch(g, q.data(), dq.data(), ddq.data(), tau.data(), B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, nullptr) != RB_OK ||
        multibody_forward_dynamics_batch(g, q.data(), dq.data(), tau.data(), qdd.data(), B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, nullptr) != RB_OK ||
        multibody_rnea_fd_batch(g, q.data(), dq.data(), ddq.data(), tau.data(), out.data(), B, 0, RB_LAYOUT_SOA, RB_MEM_HOST, nullptr) != RB_OK) {
        std::printf("batch call failed: %s\n", multibody_last_error());
        return 2;
    }
    double e_single = 0.0, e_round = 0.0, e_onecall = 0.0;
    for (int i = 0; i < n; ++i) e_single = std::fmax(e_single, std::fabs(tau[i * B] - tau1[i]));
    for (size_t k = 0; k < (size_t)n * B; ++k) {
        e_round = std::fmax(e_round, std::fabs(qdd[k] - ddq[k]));                       // FD(q, dq, RNEA(q, dq, ddq)) = ddq
        e_onecall = std::fmax(e_onecall, std::fabs(out[k] - tau[k]));                   // first block of the one-call result = tau
    }
    std::printf("batch[0] vs single-state: %.2e   FD round trip: %.2e   one-call vs two-call: %.2e\n", e_single, e_round, e_onecall);
    multibody_free_result(tau1);
    multibody_gpu_free(g);
    multibody_free(mb);
    // the one-call entry point runs the fused kernel: tau = bias + H ddq agrees with the recursion to rounding
    return (e_single == 0.0 && e_round < 1e-9 && e_onecall < 1e-11) ? 0 : 1;
}
