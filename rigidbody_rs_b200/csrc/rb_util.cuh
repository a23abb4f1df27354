// rb_util.cuh -- utility kernels of the engine (defined in rb_kernels_n.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rb_model.h"

struct RbFillRange { double lo[RB_MAX_N]; double hi[RB_MAX_N]; };
#define RB_PEAK_INNER 64

cudaError_t rb_launch_fill(double* out, uint64_t seed, uint32_t field, int n, const RbFillRange& rg,
                           size_t first, size_t count, size_t ld, cudaStream_t st);
// `per_state` doubles per state: [B][per_state] <-> [per_state][ld]
cudaError_t rb_launch_aos_to_soa(const double* aos, double* soa, int per_state, size_t B, size_t ld, cudaStream_t st);
cudaError_t rb_launch_soa_to_aos(const double* soa, double* aos, int per_state, size_t B, size_t ld, cudaStream_t st);
cudaError_t rb_launch_fp64_peak(double* out, int blocks, int threads, int iters, cudaStream_t st);
// Batched LDL^T solve in shared-memory tiles (rb_kernels_n.cu): H packed upper [n(n+1)/2][hpk_states], rhs/x in qdd.
cudaError_t rb_launch_ldlt_tiles(int n, const double* hpk, size_t hpk_states, double* qdd, size_t cnt, size_t ld,
                                 int* status, cudaStream_t st);
// Forward dynamics of a chain of n <= 32 joints, one warp per state / one lane per joint (rb_kernels_warp.cu);
// `model` = device rows in the rb_model.h layout (n x 24 doubles, then g[3]).
cudaError_t rb_launch_warp_fd(const double* model, int n, const double* q, const double* dq, const double* tau,
                              double* qdd, size_t B, size_t ld, int* status, cudaStream_t st);
#ifndef RB_WARP_FD
#define RB_WARP_FD 1
#endif
// Analytical derivatives (rb_kernels_deriv.cu) of serial chains of n <= RB_DERIV_MAX_N joints; `flat_model` = HOST rows in
// the rb_model.h layout (passed on as a kernel parameter).  out: [2 n^2][ld] resp. [3 n^2][ld], see include/rigidbody.h.
#define RB_DERIV_MAX_N 12        // forward-dynamics derivatives
#define RB_RNEA_DERIV_MAX_N 32   // inverse-dynamics derivatives
cudaError_t rb_launch_rnea_deriv(int n, const double* flat_model, const double* q, const double* dq, const double* ddq,
                                 double* out, size_t B, size_t ld, cudaStream_t st);
cudaError_t rb_launch_fd_deriv(int n, const double* flat_model, const double* q, const double* dq, const double* tau,
                               double* out, size_t B, size_t ld, int* status, cudaStream_t st);
