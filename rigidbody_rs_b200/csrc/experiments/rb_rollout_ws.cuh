// experiments/rb_rollout_ws.cuh -- warp-specialised rollout (two warps per 32 trajectories), a measured dead end.
//
// Built only with -DRB_ROLLOUT_WS=1; then RIGIDBODY_B200_ROLLOUT=ws selects it at run time.  Bit-identical to
// rb_rollout_kernel, correct (the GPU parity tests pass with it), and slower at every size
// (profiles/r2_kbench_rollout_ws.jsonl): 0.487 vs 0.408 ms at 65 536 trajectories, 0.161 vs 0.139 ms at 8 192.
// Why: the step is S (sin/cos) -> {bias recursion || CRBA + factorisation} -> solve + integrate.  Splitting it over two
// warps removes the instruction-level parallelism a single thread gets for free from interleaving the two independent
// middle parts, each half is latency-bound on its own dependent chain, and S and the solve stay serial, so the chain per
// step only drops from ~4 100 to ~3 500 cycles in the best case while half of all warp-time is spent at the two named
// barriers (ncu: 49 % of warp samples in `barrier`, FP64 pipe 61 % against 71 %).  Lessons kept in the product:
// the branch-free batched sin/cos (rb_sincos_batched) and the factorisation interleaved with the bias recursion
// (rb_forward_dynamics_interleaved) both came out of profiling this kernel.
#pragma once

// One trajectory's forward-dynamics step is a ~1900-instruction dependent chain; a thread per trajectory needs 255
// registers (8 warps per SM, blocks of 128 trajectories: 512 blocks over 148 SMs = 13 % imbalance) and, when the
// trajectories are split over several GPUs, leaves a lone warp per scheduler that no amount of SM coverage speeds up
// (profiles/r1_kbench_rollout.jsonl: 8 192 and 16 384 trajectories both take 0.16 ms).  Here TWO warps serve 32
// trajectories, lane = trajectory in both:
//   warp A ("bias"):   sin/cos of q, the bias recursion rnea(q, dq, 0), then -- once L arrives -- the two triangular
//                      solves, the semi-implicit Euler update, the trajectory stores and the running cost;
//   warp B ("matrix"): crba(q) from A's sin/cos and the LDL^T factorisation, which need neither dq nor tau; it also
//                      fetches the step's torques from HBM into shared memory while it waits for the sin/cos.
// The hand-over goes through shared memory ([value][lane]: conflict-free) with two named barriers per step.  The
// dependent chain per step drops from ~1 130 to ~670 FP64 instructions and the operations are those of
// rb_rollout_kernel in the same order: results are bit-identical (test_rollout_kernels_agree_bitwise).
// A block is 128 threads = 2 groups = 64 trajectories: warp w of a block runs on SM sub-partition w % 4 (measured:
// 64-thread blocks leave two of the four FP64 pipes idle, profiles/r2_kbench_fused_and_ws_rollout_v1.jsonl), so the
// two groups place their A and B warps crosswise and alternate blocks flip the placement: every sub-partition sees both
// roles.  1 024 blocks of 64 trajectories balance over 148 SMs to 1 %.
#ifndef RB_RO2_MINB
#define RB_RO2_MINB 4       // 128-thread blocks per SM -> 128 registers
#endif
#ifndef RB_RO2_LAZY_SINCOS
#define RB_RO2_LAZY_SINCOS 1
#endif
// Column `lane` of a [value][32] shared-memory table, indexable like an array.
template <class T> struct RbSmemColumn {
    const T* base;
    RB_DI T operator[](int i) const { return base[i * 32]; }
};
#define RB_RO2_MAX_N 12     // static shared memory: 2 x (3n + n(n+1)/2) x 32 doubles per block
RB_DI void rb_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <class M>
__global__ void __launch_bounds__(128, M::kSpecialised && M::N <= 8 ? RB_RO2_MINB : 2)
rb_rollout_ws_kernel(const __grid_constant__ typename M::Param p, const RB_R* __restrict__ q0, const RB_R* __restrict__ dq0,
                     const RB_R* __restrict__ tau, RB_R dt, int horizon, RB_R* __restrict__ q_traj,
                     RB_R* __restrict__ dq_traj, RB_R* __restrict__ q_fin, RB_R* __restrict__ dq_fin,
                     size_t B, size_t ld, int* __restrict__ status, const RB_R* __restrict__ cost_w, RB_R* __restrict__ cost) {
    constexpr int N = M::N, NL = N * (N - 1) / 2;
    __shared__ RB_R sh_sc_[2][2 * N][32];    // sin (rows 0..N-1) and cos (N..2N-1) of this step's q
    __shared__ RB_R sh_ld_[2][NL + N][32];   // L (strict upper of the factorised H, row-major) then 1/d
    __shared__ RB_R sh_u_[2][N][32];         // this step's torques
    __shared__ int sh_ok_[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = warp >> 1;
    // A on warps 0 and 3, B on 1 and 2; flipped in every other block of an SM's share (blocks b, b + #SMs, ...)
    unsigned nsm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const bool flip = ((blockIdx.x / nsm) & 1u) != 0;
#ifdef RB_RO2_ONLY      // register-pressure probe: compile one role only (never a product build)
    const bool role_a = RB_RO2_ONLY == 1;
#else
    const bool role_a = ((warp == 0 || warp == 3) != flip);
#endif
    RB_R (&sh_sc)[2 * N][32] = sh_sc_[grp];
    RB_R (&sh_ld)[NL + N][32] = sh_ld_[grp];
    RB_R (&sh_u)[N][32] = sh_u_[grp];
    int (&sh_ok)[32] = sh_ok_[grp];
    const int bar1 = 1 + 2 * grp, bar2 = 2 + 2 * grp;
    const size_t s = ((size_t)blockIdx.x * 2 + grp) * 32 + lane;
    const bool live = s < B;                 // padding lanes run the arithmetic on zeros and store nothing
    const size_t step = (size_t)N * ld;
    if (role_a) {
        // ---------------------------------------------------------------- warp A
        RB_R q[N], dq[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { q[i] = RB_R(0); dq[i] = RB_R(0); }
        if (live) { rb_load<N>(q0, ld, s, q); rb_load<N>(dq0, ld, s, dq); }
        RB_R J = RB_R(0);
        for (int t = 0; t < horizon; ++t) {
            RB_R x[N];
            {
                RB_R sn[N], cs[N];
                rb_sincos_all<N>(q, sn, cs);
#pragma unroll
                for (int i = 0; i < N; ++i) { sh_sc[i][lane] = sn[i]; sh_sc[N + i][lane] = cs[i]; }
                rb_bar_sync(bar1, 64);                            // sin/cos published
                // Read them back: arithmetic is free to move across a barrier, and the whole bias recursion was scheduled
                // BEFORE this one (the matrix warp then started a full recursion late: ncu showed half of all warp samples
                // waiting at barriers).  A load after the barrier is a dependency the scheduler cannot break.
#pragma unroll
                for (int i = 0; i < N; ++i) { sn[i] = sh_sc[i][lane]; cs[i] = sh_sc[N + i][lane]; }
                rb_rnea<M, false>(p, sn, cs, dq, dq /*unused*/, x);  // bias
            }
            rb_bar_sync(bar2, 64);                                // L, 1/d and the torques published by warp B
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = sh_u[i][lane] - x[i];
            {
                RB_R dinv[N];
#pragma unroll
                for (int i = 0; i < N; ++i) dinv[i] = sh_ld[NL + i][lane];
                rb_ldlt_apply_fn<N>([&](auto jc, auto ic) {
                    constexpr int Jr = decltype(jc)::value, Ic = decltype(ic)::value;       // row Jr < column Ic
                    return sh_ld[Jr * N - Jr * (Jr + 1) / 2 + (Ic - Jr - 1)][lane];
                }, dinv, x);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                dq[i] = fma(dt, x[i], dq[i]);
                q[i] = fma(dt, dq[i], q[i]);
            }
            if (live) {
                if (q_traj) rb_store<N>(q_traj + (size_t)t * step, ld, s, q);
                if (dq_traj) rb_store<N>(dq_traj + (size_t)t * step, ld, s, dq);
            }
            if (cost) {                                           // same accumulation order as rb_rollout_kernel
                RB_R c = RB_R(0);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const RB_R e = q[i] - __ldg(cost_w + RB_CW_QREF * RB_MAX_N + i);
                    const RB_R u = sh_u[i][lane];                 // still this step's: B rewrites it after the next bar1
                    c = fma(__ldg(cost_w + RB_CW_Q * RB_MAX_N + i) * e, e, c);
                    c = fma(__ldg(cost_w + RB_CW_DQ * RB_MAX_N + i) * dq[i], dq[i], c);
                    c = fma(__ldg(cost_w + RB_CW_TAU * RB_MAX_N + i) * u, u, c);
                }
                J = fma(dt, c, J);
            }
        }
        rb_bar_sync(bar1, 64);                                    // warp B's verdict on positive definiteness
        if (live) {
            if (cost) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const RB_R e = q[i] - __ldg(cost_w + RB_CW_QREF * RB_MAX_N + i);
                    J = fma(__ldg(cost_w + RB_CW_QF * RB_MAX_N + i) * e, e, J);
                    J = fma(__ldg(cost_w + RB_CW_DQF * RB_MAX_N + i) * dq[i], dq[i], J);
                }
                __stcs(cost + s, sh_ok[lane] ? J : rb_nan<RB_R>());
            }
            if (q_fin) rb_store<N>(q_fin, ld, s, q);
            if (dq_fin) rb_store<N>(dq_fin, ld, s, dq);
        }
    } else {
        // ---------------------------------------------------------------- warp B
        bool ok = true;
        RB_R u[N];                                  // torques of the coming step, fetched one step ahead (HBM latency
#pragma unroll                                      // would otherwise sit on the critical path of every step)
        for (int i = 0; i < N; ++i) u[i] = RB_R(0);
        if (live && horizon > 0) rb_load<N>(tau, ld, s, u);
        for (int t = 0; t < horizon; ++t) {
            RB_R H[N][N], dinv[N];
            rb_bar_sync(bar1, 64);
#pragma unroll
            for (int i = 0; i < N; ++i) sh_u[i][lane] = u[i];
            if (live && t + 1 < horizon) rb_load<N>(tau + (size_t)(t + 1) * step, ld, s, u);
#if RB_RO2_LAZY_SINCOS
            // sin/cos stay in shared memory and are read where a transform needs them (24 fewer live registers)
            const RbSmemColumn<RB_R> sn{&sh_sc[0][lane]}, cs{&sh_sc[N][lane]};
#else
            RB_R sn[N], cs[N];
#pragma unroll
            for (int i = 0; i < N; ++i) { sn[i] = sh_sc[i][lane]; cs[i] = sh_sc[N + i][lane]; }
#endif
            rb_crba<M>(p, sn, cs, H);
            ok = rb_ldlt_factor<N>(H, dinv) && ok;
            rb_for_up<0, N>([&](auto jc) {
                constexpr int Jr = decltype(jc)::value;
                rb_for_up<Jr + 1, N>([&](auto ic) {
                    constexpr int Ic = decltype(ic)::value;
                    sh_ld[Jr * N - Jr * (Jr + 1) / 2 + (Ic - Jr - 1)][lane] = H[Jr][Ic];
                });
                sh_ld[NL + Jr][lane] = dinv[Jr];
            });
            rb_bar_sync(bar2, 64);
        }
        sh_ok[lane] = ok ? 1 : 0;
        rb_bar_sync(bar1, 64);
        if (!ok && live) atomicOr(status, RB_STATUS_NOT_SPD);
    }
}

