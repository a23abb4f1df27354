// rb_kernels_n.cu -- run-time-n kernel family ("generic-n") + the small utility kernels of the engine
// (sampler fill, AoS<->SoA transposes, FP64 peak probe).
//
// generic-n: persistent grid-stride kernels; the model rows sit in global memory (uniform loads, cached in
// L1/constant path); per-thread per-link state goes to a strided scratch [slot][thread] in HBM.
#include "rb_kernels.cuh"
#include "rb_dyn_n.cuh"
#include "rb_util.cuh"

__global__ void __launch_bounds__(RB_BLOCK)
rbn_rnea_kernel(RbNParam P, const double* __restrict__ q, const double* __restrict__ dq, const double* __restrict__ ddq,
                double* __restrict__ tau, size_t B, size_t ld) {
    const size_t tid = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    const int n = P.n;
    const RbJointK* jt = reinterpret_cast<const RbJointK*>(P.model);
    const double* g = P.model + (size_t)n * 24;
    RbScratch sc{P.scratch + tid, P.threads};
    for (size_t s = tid; s < B; s += nthr) {
        for (int i = 0; i < n; ++i) {
            double sn, cs;
            sincos(__ldcs(q + (size_t)i * ld + s), &sn, &cs);
            sc[i] = sn; sc[n + i] = cs;
        }
        RbScratch out{tau + s, ld};
        rbn_rnea(jt, g, n, sc, dq + s, ddq + s, ld, out);
    }
}

// ---- forward dynamics for long chains, two kernels per chunk of states ----------------------------------------
// A 32-joint mass matrix is 528 doubles: no thread can hold it, and factorising it in a per-thread global scratch
// costs O(n^3) HBM accesses per state (the first version of this family did that: 0.027 G states/s for n = 32).  Instead:
//   1. rbn_fd_prepare_kernel (thread per state, persistent): sin/cos, bias forces -> rhs = tau - rnea(q,dq,0)
//      written into qdd, CRBA -> packed upper triangle of H streamed ONCE to HBM, coalesced ([k][state]);
//   2. rbn_ldlt_tile_kernel: a warp pulls a tile of TS states (all n(n+1)/2 rows, 256-byte runs) into shared
//      memory, lane = state, and runs the dot-product (Crout) LDL^T + both triangular solves entirely on chip:
//      2 shared loads per FMA, no bank conflicts ([k][lane] layout), H read from HBM exactly once.
// HBM traffic per state: 2 * n(n+1)/2 * 8 B for H (8.4 KB at n = 32) + the 32 n algorithmic bytes.
__global__ void __launch_bounds__(RB_BLOCK)
rbn_fd_prepare_kernel(RbNParam P, const double* __restrict__ q, const double* __restrict__ dq, const double* __restrict__ tau,
                      double* __restrict__ qdd, size_t B, size_t ld) {
    const size_t tid = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    const int n = P.n;
    const RbJointK* jt = reinterpret_cast<const RbJointK*>(P.model);
    const double* g = P.model + (size_t)n * 24;
    RbScratch sc{P.scratch + tid, P.threads};
    for (size_t s = tid; s < B; s += nthr) {          // B <= P.hpk_states: s doubles as the chunk-local index
        for (int i = 0; i < n; ++i) {
            double sn, cs;
            sincos(__ldcs(q + (size_t)i * ld + s), &sn, &cs);
            sc[i] = sn; sc[n + i] = cs;
        }
        RbScratch x{qdd + s, ld};
        rbn_rnea(jt, g, n, sc, dq + s, nullptr, ld, x);                       // bias
        for (int i = 0; i < n; ++i) x[i] = __ldcs(tau + (size_t)i * ld + s) - x[i];
        double* hp = P.hpk + s;
        const size_t hs = P.hpk_states;
        rbn_crba(jt, n, sc, [&](int r, int c, double v) {
            __stcs(hp + (size_t)(r * n - r * (r - 1) / 2 + (c - r)) * hs, v);
        });
    }
}

// One block = RB_TILE_WARPS warps sharing one tile of TS states; lane = state, warps split the columns.
// Shared memory: S[np][TS] (packed upper, row j = entries (j, j..n-1)), M[n][TS], X[n][TS], DI[n][TS].
// The tile arrives through cp.async (8 bytes per lane per row, every row of the tile in flight at once): a plain
// load loop in one warp left ~528 dependent-latency round trips per tile and ran 20x slower.
#ifndef RB_TILE_WARPS
#define RB_TILE_WARPS 16
#endif
// True iff every pivot of this lane's state was positive (DI holds 1/d_j; NaN and non-positive fail).
__device__ __forceinline__ bool ok_all_lanes(bool, const double* DI, int n, int TS, int lane) {
    bool ok = true;
    for (int j = 0; j < n; ++j) ok = ok && (DI[j * TS + lane] > 0.0) && (DI[j * TS + lane] < 1.0e300);
    return ok;
}
template <int TS>
__global__ void __launch_bounds__(32 * RB_TILE_WARPS)
rbn_ldlt_tile_kernel(int n, const double* __restrict__ hpk, size_t hpk_states, double* __restrict__ qdd, size_t B, size_t ld,
                     int* __restrict__ status) {
    extern __shared__ double rb_tile[];
    const int np = n * (n + 1) / 2;
    double* S = rb_tile;
    double* Mv = S + (size_t)np * TS;
    double* X = Mv + (size_t)n * TS;
    double* DI = X + (size_t)n * TS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t tiles = (B + TS - 1) / TS;
    auto row = [&](int j) { return j * n - j * (j - 1) / 2 - j; };           // S(j, i) at (row(j) + i) * TS + lane
    bool all_ok = true;
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const size_t s = tile * TS + lane;
        const bool live = lane < TS && s < B;
        if (live) {
            for (int k = w; k < np; k += RB_TILE_WARPS) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(S + (size_t)k * TS + lane);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(hpk + (size_t)k * hpk_states + s) : "memory");
            }
            for (int i = w; i < n; i += RB_TILE_WARPS) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(X + i * TS + lane);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(qdd + (size_t)i * ld + s) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        bool ok = true;
        // factorise: S(j, i) <- A(j, i) - sum_{k<j} (S(k, j) / d_k) S(k, i),  d_j = S(j, j).
        // The right-hand side rides along as column n (same recurrence), which is the forward substitution:
        // X_j <- b_j - sum_{k<j} (S(k, j) / d_k) X_k = (D U x)_j.
        for (int j = 0; j < n; ++j) {
            if (live)
                for (int k = w; k < j; k += RB_TILE_WARPS) Mv[k * TS + lane] = S[(size_t)(row(k) + j) * TS + lane] * DI[k * TS + lane];
            __syncthreads();
            if (live) {
                const int rj = row(j);
                // warp w owns columns j + w, j + w + W, ... of row j; column n is the right-hand side
                for (int i = j + w; i <= n; i += 2 * RB_TILE_WARPS) {
                    const int i1 = i + RB_TILE_WARPS;
                    const bool two = i1 <= n;
                    double* p0 = i < n ? S + (size_t)(rj + i) * TS + lane : X + j * TS + lane;
                    double* p1 = two ? (i1 < n ? S + (size_t)(rj + i1) * TS + lane : X + j * TS + lane) : p0;
                    double a0 = *p0, a1 = *p1;
                    int rk = 0;                                      // row(k), advanced incrementally
#pragma unroll 4
                    for (int k = 0; k < j; ++k) {
                        const double m = Mv[k * TS + lane];
                        const double s0 = i < n ? S[(size_t)(rk + i) * TS + lane] : X[k * TS + lane];
                        const double s1 = i1 < n ? S[(size_t)(rk + (two ? i1 : i)) * TS + lane] : X[k * TS + lane];
                        a0 = fma(-m, s0, a0);
                        a1 = fma(-m, s1, a1);
                        rk += n - k - 1;
                    }
                    *p0 = a0;
                    if (two) *p1 = a1;
                    if (i == j) { ok = ok && (a0 > 0.0); DI[j * TS + lane] = 1.0 / a0; }
                }
            }
            __syncthreads();
        }
        // U x = z with z_k = X_k / d_k, column-oriented so the warps split the rows:
        //   for k = n-1..0: x_k = X_k / d_k (final); every i < k: X_i -= S(i, k) x_k
        const bool okv = live ? ok_all_lanes(ok, DI, n, TS, lane) : true;
        for (int k = n - 1; k >= 0; --k) {
            if (live) {
                const double xk = X[k * TS + lane] * DI[k * TS + lane];
                for (int i = w; i < k; i += RB_TILE_WARPS)
                    X[i * TS + lane] = fma(-S[(size_t)(row(i) + k) * TS + lane], xk, X[i * TS + lane]);
                if (w == (k & (RB_TILE_WARPS - 1)))
                    __stcs(qdd + (size_t)k * ld + s, okv ? xk : __longlong_as_double(0x7ff8000000000000LL));
            }
            __syncthreads();
        }
        all_ok = all_ok && okv;
        __syncthreads();                                             // tile buffers are reused by the next iteration
    }
    if (!all_ok) atomicOr(status, RB_STATUS_NOT_SPD);
}

__global__ void __launch_bounds__(RB_BLOCK)
rbn_crba_kernel(RbNParam P, const double* __restrict__ q, double* __restrict__ Hout, size_t B, size_t ld) {
    const size_t tid = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    const int n = P.n;
    const RbJointK* jt = reinterpret_cast<const RbJointK*>(P.model);
    RbScratch sc{P.scratch + tid, P.threads};
    for (size_t s = tid; s < B; s += nthr) {
        for (int i = 0; i < n; ++i) {
            double sn, cs;
            sincos(__ldcs(q + (size_t)i * ld + s), &sn, &cs);
            sc[i] = sn; sc[n + i] = cs;
        }
        for (int c = 0; c < n; ++c)
            for (int r = c + 1; r < n; ++r) __stcs(Hout + (size_t)(r + n * c) * ld + s, 0.0);
        rbn_crba(jt, n, sc, [&](int r, int c, double v) { __stcs(Hout + (size_t)(r + n * c) * ld + s, v); });
    }
}

__global__ void __launch_bounds__(RB_BLOCK)
rbn_fk_jac_kernel(RbNParam P, const double* __restrict__ q, double* __restrict__ xyz, double* __restrict__ Jout,
                  size_t B, size_t ld) {
    const size_t tid = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    const int n = P.n;
    const RbJointK* jt = reinterpret_cast<const RbJointK*>(P.model);
    for (size_t s = tid; s < B; s += nthr) {
        double A[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
        double r[3] = {0.0, 0.0, 0.0};
        for (int i = n - 1; i >= 0; --i) {
            const RbJointK& j = jt[i];
            if (Jout) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    __stcs(Jout + (size_t)(k + 6 * i) * ld + s, fma(A[1][k], r[0], -(A[0][k] * r[1])));
                    __stcs(Jout + (size_t)(3 + k + 6 * i) * ld + s, A[2][k]);
                }
            }
            double sn, cs;
            sincos(__ldcs(q + (size_t)i * ld + s), &sn, &cs);
            double Bm[3][3], y[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Bm[0][k] = fma(cs, A[0][k], -(sn * A[1][k]));
                Bm[1][k] = fma(sn, A[0][k], cs * A[1][k]);
                Bm[2][k] = A[2][k];
            }
            y[0] = fma(cs, r[0], -(sn * r[1])); y[1] = fma(sn, r[0], cs * r[1]); y[2] = r[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    A[a][k] = fma(j.R[3 * a + 2], Bm[2][k], fma(j.R[3 * a + 1], Bm[1][k], j.R[3 * a] * Bm[0][k]));
                r[a] = fma(j.R[3 * a + 2], y[2], fma(j.R[3 * a + 1], y[1], fma(j.R[3 * a], y[0], j.t[a])));
            }
        }
        if (xyz) {
            __stcs(xyz + s, r[0]); __stcs(xyz + ld + s, r[1]); __stcs(xyz + 2 * ld + s, r[2]);
        }
    }
}

// scratch slots: [0,2n) sincos, [2n,8n) f, [8n, 8n+n*n) H, n rhs/x, n dinv, 2n carried (q, dq).  The per-step solve
// still factorises H in the per-thread global scratch (slow for long chains; rollouts of long chains are not a
// BASELINE.json configuration).
__global__ void __launch_bounds__(RB_BLOCK)
rbn_rollout_kernel(RbNParam P, const double* __restrict__ q0, const double* __restrict__ dq0, const double* __restrict__ tau,
                   double dt, int horizon, double* __restrict__ q_traj, double* __restrict__ dq_traj,
                   double* __restrict__ q_fin, double* __restrict__ dq_fin, size_t B, size_t ld, int* __restrict__ status) {
    const size_t tid = (size_t)blockIdx.x * RB_BLOCK + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * RB_BLOCK;
    const int n = P.n;
    const RbJointK* jt = reinterpret_cast<const RbJointK*>(P.model);
    const double* g = P.model + (size_t)n * 24;
    RbScratch sc{P.scratch + tid, P.threads};
    RbScratch Hs{P.scratch + tid + (size_t)(8 * n) * P.threads, P.threads};
    RbScratch x{P.scratch + tid + (size_t)(8 * n + n * n) * P.threads, P.threads};
    RbScratch dinv{P.scratch + tid + (size_t)(9 * n + n * n) * P.threads, P.threads};
    RbScratch qs{P.scratch + tid + (size_t)(10 * n + n * n) * P.threads, P.threads};
    RbScratch dqs{P.scratch + tid + (size_t)(11 * n + n * n) * P.threads, P.threads};
    const size_t step = (size_t)n * ld;
    bool all_ok = true;
    for (size_t s = tid; s < B; s += nthr) {
        for (int i = 0; i < n; ++i) { qs[i] = q0[(size_t)i * ld + s]; dqs[i] = dq0[(size_t)i * ld + s]; }
        for (int t = 0; t < horizon; ++t) {
            for (int i = 0; i < n; ++i) {
                double sn, cs;
                sincos(qs[i], &sn, &cs);
                sc[i] = sn; sc[n + i] = cs;
            }
            rbn_rnea(jt, g, n, sc, dqs.base, nullptr, dqs.stride, x);
            for (int i = 0; i < n; ++i) x[i] = __ldcs(tau + (size_t)t * step + (size_t)i * ld + s) - x[i];
            rbn_crba(jt, n, sc, [&](int r, int c, double v) { Hs[r * n + c] = v; });
            all_ok = rbn_ldlt_solve(n, Hs, x, dinv) && all_ok;
            for (int i = 0; i < n; ++i) {
                const double dqn = fma(dt, x[i], dqs[i]);
                const double qn = fma(dt, dqn, qs[i]);
                dqs[i] = dqn; qs[i] = qn;
                if (q_traj) __stcs(q_traj + (size_t)t * step + (size_t)i * ld + s, qn);
                if (dq_traj) __stcs(dq_traj + (size_t)t * step + (size_t)i * ld + s, dqn);
            }
        }
        for (int i = 0; i < n; ++i) {
            if (q_fin) q_fin[(size_t)i * ld + s] = qs[i];
            if (dq_fin) dq_fin[(size_t)i * ld + s] = dqs[i];
        }
    }
    if (!all_ok) atomicOr(status, RB_STATUS_NOT_SPD);
}

namespace {
unsigned ngrid(const RbNParam* P, size_t B) {
    size_t want = (B + RB_BLOCK - 1) / RB_BLOCK;
    size_t cap = P->threads / RB_BLOCK;
    return (unsigned)(want < cap ? want : cap);
}
cudaError_t n_rnea(const void* param, const double* q, const double* dq, const double* ddq, double* tau, size_t B, size_t ld, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    if (B == 0) return cudaSuccess;
    rbn_rnea_kernel<<<ngrid(P, B), RB_BLOCK, 0, st>>>(*P, q, dq, ddq, tau, B, ld);
    return cudaGetLastError();
}
template <int TS>
cudaError_t launch_tile(int n, const double* hpk, size_t hpk_states, double* qdd, size_t cnt, size_t ld, int* status,
                        size_t smem, int dev, int sms, cudaStream_t st) {
    static size_t configured[64] = {0};     // per device: largest dynamic shared-memory size opted into
    if (dev >= 0 && dev < 64 && smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(rbn_ldlt_tile_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    const size_t tiles = (cnt + TS - 1) / TS;
    size_t per_sm = (size_t)(220 * 1024) / (smem + 1024);
    per_sm = per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm);
    const size_t cap = (size_t)sms * per_sm;
    rbn_ldlt_tile_kernel<TS><<<(unsigned)(tiles < cap ? tiles : cap), 32 * RB_TILE_WARPS, smem, st>>>(n, hpk, hpk_states, qdd, cnt, ld, status);
    return cudaGetLastError();
}
}  // namespace

// Solves H x = rhs for `cnt` states: H packed upper in hpk ([n(n+1)/2][hpk_states]), rhs in qdd (overwritten by x).
cudaError_t rb_launch_ldlt_tiles(int n, const double* hpk, size_t hpk_states, double* qdd, size_t cnt, size_t ld,
                                 int* status, cudaStream_t st) {
    if (cnt == 0) return cudaSuccess;
    const size_t rows = (size_t)n * (n + 1) / 2 + 3 * (size_t)n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // widest tile whose shared-memory footprint fits one SM
    if (rows * 32 * sizeof(double) <= 200 * 1024) return launch_tile<32>(n, hpk, hpk_states, qdd, cnt, ld, status, rows * 32 * sizeof(double), dev, sms, st);
    if (rows * 16 * sizeof(double) <= 200 * 1024) return launch_tile<16>(n, hpk, hpk_states, qdd, cnt, ld, status, rows * 16 * sizeof(double), dev, sms, st);
    return launch_tile<8>(n, hpk, hpk_states, qdd, cnt, ld, status, rows * 8 * sizeof(double), dev, sms, st);
}

namespace {
cudaError_t n_fd(const void* param, const double* q, const double* dq, const double* tau, double* qdd, size_t B, size_t ld, int* status, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    for (size_t off = 0; off < B; off += P->hpk_states) {
        const size_t cnt = (B - off < P->hpk_states) ? (B - off) : P->hpk_states;
        rbn_fd_prepare_kernel<<<ngrid(P, cnt), RB_BLOCK, 0, st>>>(*P, q + off, dq + off, tau + off, qdd + off, cnt, ld);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = rb_launch_ldlt_tiles(P->n, P->hpk, P->hpk_states, qdd + off, cnt, ld, status, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
cudaError_t n_crba(const void* param, const double* q, double* H, size_t B, size_t ld, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    if (B == 0) return cudaSuccess;
    rbn_crba_kernel<<<ngrid(P, B), RB_BLOCK, 0, st>>>(*P, q, H, B, ld);
    return cudaGetLastError();
}
cudaError_t n_fk(const void* param, const double* q, double* xyz, size_t B, size_t ld, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    if (B == 0) return cudaSuccess;
    rbn_fk_jac_kernel<<<ngrid(P, B), RB_BLOCK, 0, st>>>(*P, q, xyz, nullptr, B, ld);
    return cudaGetLastError();
}
cudaError_t n_jac(const void* param, const double* q, double* J, size_t B, size_t ld, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    if (B == 0) return cudaSuccess;
    rbn_fk_jac_kernel<<<ngrid(P, B), RB_BLOCK, 0, st>>>(*P, q, nullptr, J, B, ld);
    return cudaGetLastError();
}
cudaError_t n_rollout(const void* param, const double* q0, const double* dq0, const double* tau, double dt, int horizon,
                      double* q_traj, double* dq_traj, double* q_fin, double* dq_fin, size_t B, size_t ld, int* status, cudaStream_t st) {
    const RbNParam* P = (const RbNParam*)param;
    if (B == 0) return cudaSuccess;
    rbn_rollout_kernel<<<ngrid(P, B), RB_BLOCK, 0, st>>>(*P, q0, dq0, tau, dt, horizon, q_traj, dq_traj, q_fin, dq_fin, B, ld, status);
    return cudaGetLastError();
}
}  // namespace

const RbOps* rb_ops_generic_n() {
    static const RbOps ops = {"generic-n", 0, sizeof(RbNParam), &n_rnea, &n_fd, nullptr, nullptr, &n_crba, &n_fk, &n_jac, &n_rollout};
    return &ops;
}

// ------------------------------------------------------------------ utility kernels (declared in rb_util.cuh)
__global__ void rb_fill_kernel(double* __restrict__ out, uint64_t seed, uint32_t field, int n, RbFillRange rg,
                               size_t first, size_t count, size_t ld) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    for (int i = 0; i < n; ++i) {
        const uint64_t ctr = ((uint64_t)field << 58) | ((uint64_t)i << 50) | ((first + s) & ((1ULL << 50) - 1));
        uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        const double u = (double)(z >> 11) * 0x1.0p-53;
        out[(size_t)i * ld + s] = fma(rg.hi[i] - rg.lo[i], u, rg.lo[i]);
    }
}

// [B][n] (AoS) <-> [n][ld] (SoA): `spb` states per block pass through shared memory so both the global reads
// and the global writes are contiguous runs.
__global__ void rb_aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, int n, size_t B, size_t ld, int spb) {
    extern __shared__ double tile[];               // [spb][n] as in memory
    const size_t s0 = (size_t)blockIdx.x * spb;
    const size_t cnt = (B - s0 < (size_t)spb) ? (B - s0) : (size_t)spb;
    const size_t tot = cnt * (size_t)n;
    for (size_t k = threadIdx.x; k < tot; k += blockDim.x) tile[k] = aos[s0 * n + k];
    __syncthreads();
    for (size_t k = threadIdx.x; k < tot; k += blockDim.x) {
        const size_t i = k / cnt, s = k % cnt;
        soa[i * ld + s0 + s] = tile[s * n + i];
    }
}
__global__ void rb_soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, int n, size_t B, size_t ld, int spb) {
    extern __shared__ double tile[];
    const size_t s0 = (size_t)blockIdx.x * spb;
    const size_t cnt = (B - s0 < (size_t)spb) ? (B - s0) : (size_t)spb;
    const size_t tot = cnt * (size_t)n;
    for (size_t k = threadIdx.x; k < tot; k += blockDim.x) {
        const size_t i = k / cnt, s = k % cnt;
        tile[s * n + i] = soa[i * ld + s0 + s];
    }
    __syncthreads();
    for (size_t k = threadIdx.x; k < tot; k += blockDim.x) aos[s0 * n + k] = tile[k];
}

// Register-only DFMA chains: 8 independent accumulators per thread, RB_PEAK_INNER fused multiply-adds each
// per outer iteration.  2 flops per DFMA.
__global__ void __launch_bounds__(256) rb_fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < RB_PEAK_INNER; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456) out[0] = r;   // keeps the chains live without a store in the common case
}

cudaError_t rb_launch_fill(double* out, uint64_t seed, uint32_t field, int n, const RbFillRange& rg,
                           size_t first, size_t count, size_t ld, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    rb_fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(out, seed, field, n, rg, first, count, ld);
    return cudaGetLastError();
}
static int rb_tr_spb(int n) { int spb = 4096 / n; return spb < 1 ? 1 : (spb > 128 ? 128 : spb); }
cudaError_t rb_launch_aos_to_soa(const double* aos, double* soa, int n, size_t B, size_t ld, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    const int spb = rb_tr_spb(n);
    rb_aos_to_soa_kernel<<<(unsigned)((B + spb - 1) / spb), 256, (size_t)spb * n * sizeof(double), st>>>(aos, soa, n, B, ld, spb);
    return cudaGetLastError();
}
cudaError_t rb_launch_soa_to_aos(const double* soa, double* aos, int n, size_t B, size_t ld, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    const int spb = rb_tr_spb(n);
    rb_soa_to_aos_kernel<<<(unsigned)((B + spb - 1) / spb), 256, (size_t)spb * n * sizeof(double), st>>>(soa, aos, n, B, ld, spb);
    return cudaGetLastError();
}
cudaError_t rb_launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t st) {
    rb_fp64_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.999999, 1e-7);
    return cudaGetLastError();
}
