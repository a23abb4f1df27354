"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): fp64, per state max_i|x_i - ref_i| <= 1e-10 * max(1, ||ref||_inf).
Small/medium sizes compare against the oracle directly and against the frozen golden vectors; the full
BASELINE.json sizes use size-independent properties (FD round trip, the CRBA/RNEA identity, a strided oracle
sample) because the oracle would need minutes there.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import CHAIN32, FR3, TOL, load_golden, state_err

pytestmark = pytest.mark.gpu


def _states(orc, B, seed=0x5EED0001, first=0):
    lim = orc.model
    q = orc.fill(seed, 0, lim.lower, lim.upper, first, B)
    dq = orc.fill(seed, 1, -lim.velocity, lim.velocity, first, B)
    ddq = orc.fill(seed, 2, -10.0, 10.0, first, B)
    tau = orc.fill(seed + 1, 3, -lim.effort, lim.effort, first, B)
    return q, dq, ddq, tau


def _variants(rb, urdf):
    """Every kernel family able to serve the chain (selected via RIGIDBODY_B200_VARIANT)."""
    out = []
    for v in ("auto", "jit-specialised", "generic-7", "generic-n"):
        os.environ["RIGIDBODY_B200_VARIANT"] = v
        try:
            out.append(rb.Multibody.from_urdf(urdf))
        except rb.RigidBodyError:
            pass
        finally:
            os.environ.pop("RIGIDBODY_B200_VARIANT", None)
    return out


def test_kernel_families_selected(rb, mb_fr3, mb_chain32):
    assert mb_fr3.kernel_variant == "fr3-specialised"
    assert mb_chain32.kernel_variant == "chain32-specialised"
    assert [m.kernel_variant for m in _variants(rb, CHAIN32)] == ["chain32-specialised", "generic-n"]
    names = [m.kernel_variant for m in _variants(rb, FR3)]
    assert names == ["fr3-specialised", "jit-specialised", "generic-7", "generic-n"]


def test_golden_vectors_fr3_all_families(rb):
    g = load_golden("fr3_oracle.json")
    q, dq, ddq, tau = (np.array(g[k]) for k in ("q", "dq", "ddq", "tau_in"))
    for mb in _variants(rb, FR3):
        assert state_err(mb.rnea(q, dq, ddq, layout="aos"), np.array(g["rnea"]), 1).max() < TOL, mb.kernel_variant
        assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), np.array(g["fd"]), 1).max() < TOL
        assert state_err(mb.crba(q[:8], layout="aos"), np.array(g["crba_colmajor"]), 1).max() < TOL
        assert np.abs(mb.fwd_kin(q[:8], layout="aos") - np.array(g["fwd_kin"])).max() < TOL
        assert np.abs(mb.jac(q[:8], layout="aos") - np.array(g["jac_colmajor"])).max() < TOL


def test_golden_vectors_chain32(rb):
    g = load_golden("chain32_oracle.json")
    q, dq, ddq, tau = (np.array(g[k]) for k in ("q", "dq", "ddq", "tau_in"))
    for mb in _variants(rb, CHAIN32):
        assert state_err(mb.rnea(q, dq, ddq, layout="aos"), np.array(g["rnea"]), 1).max() < TOL, mb.kernel_variant
        assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), np.array(g["fd"]), 1).max() < TOL, mb.kernel_variant
        assert state_err(mb.crba(q, layout="aos"), np.array(g["crba_colmajor"])[:4], 1).max() < TOL, mb.kernel_variant


def test_survey_kats_and_single_state_calls(rb, mb_fr3):
    """One state through the batched entry points = the reference's single-state FFI calls (lib.rs:15-70)."""
    k = load_golden("survey_kat.json")
    for key in ("rnea_zero", "main_cpp", "generic"):
        c = k[key]
        np.testing.assert_allclose(mb_fr3.rnea(c["q"], c["dq"], c["ddq"]), c["tau"], rtol=0, atol=2e-12)
    c = k["main_cpp"]
    H = mb_fr3.crba(c["q"])
    np.testing.assert_allclose(np.diag(H), c["crba_diag"], rtol=1e-12)
    np.testing.assert_allclose([H[0, 1], H[1, 3], H[5, 6]], [c["H01"], c["H13"], c["H56"]], rtol=1e-11)
    assert np.all(np.tril(H, -1) == 0.0)
    np.testing.assert_allclose(mb_fr3.fwd_kin(np.zeros(7)), k["fwd_kin_zero"], atol=1e-15)


def test_reference_symbols_single_state(rb, oracle_fr3):
    """Part 1 of the header: multibody_new_from_urdf / multibody_rnea / crba / fwd_kin / jac / free."""
    from rigidbody_rs_b200 import _lib
    lib = _lib.lib
    mb = lib.multibody_new_from_urdf(FR3.encode())
    assert mb
    arr = lambda v: (C.c_double * len(v))(*v)
    q, dq, ddq = [0, 0, 1, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 0, 1, 0]      # main.cpp:103-105
    p = lib.multibody_rnea(mb, arr(q), arr(dq), arr(ddq))
    assert p, lib.multibody_last_error()
    np.testing.assert_allclose(p[:7], oracle_fr3.rnea(q, dq, ddq), rtol=0, atol=1e-12)
    lib.multibody_free_result(p)
    p = lib.multibody_crba(mb, arr(q))
    H = np.array(p[:49]).reshape(7, 7).T                                                   # H[i+7*j], main.cpp:91-95
    np.testing.assert_allclose(H, oracle_fr3.crba(q), rtol=0, atol=1e-12)
    lib.multibody_free_result(p)
    p = lib.multibody_fwd_kin(mb, arr(q))
    np.testing.assert_allclose(p[:3], oracle_fr3.fwd_kin(q), rtol=0, atol=1e-13)
    lib.multibody_free_result(p)
    p = lib.multibody_jac(mb, arr(q))
    J = np.array(p[:42]).reshape(7, 6).T                                                   # J[6*i+j], main.cpp:76-79
    np.testing.assert_allclose(J, oracle_fr3.jac(q), rtol=0, atol=1e-13)
    lib.multibody_free_result(p)
    lib.multibody_free(mb)


@pytest.mark.parametrize("B", [1, 31, 128, 129, 5000])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_rnea_fd_host_batches(mb_fr3, oracle_fr3, B, layout):
    q, dq, ddq, tau = _states(oracle_fr3, B)
    ax = 0
    want_t = oracle_fr3.rnea_batch(q, dq, ddq)
    want_a = oracle_fr3.forward_dynamics_batch(q, dq, tau)
    if layout == "aos":
        q, dq, ddq, tau, want_t, want_a = (np.ascontiguousarray(x.T) for x in (q, dq, ddq, tau, want_t, want_a))
        ax = 1
    assert state_err(mb_fr3.rnea(q, dq, ddq, layout=layout), want_t, ax).max() < TOL
    assert state_err(mb_fr3.forward_dynamics(q, dq, tau, layout=layout), want_a, ax).max() < TOL


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_rnea_fd_one_call_equals_two_calls(rb, mb_fr3, oracle_fr3, layout):
    """multibody_rnea_fd_batch: q and dq staged once; host and device, both layouts, chunked sizes.  The fused kernel
    (rb_rnea_fd_fused) shares sin/cos, the bias recursion and the mass matrix: qdd is bit-identical to the separate
    forward-dynamics call, tau = bias + sym(H) ddq differs from the recursion by rounding only (held to 1e-12 here,
    1e-10 against the oracle).  With RIGIDBODY_B200_FUSED=0 the call issues the two launches: everything bit-identical."""
    import torch
    B = 3000
    q, dq, ddq, tau = _states(oracle_fr3, B)
    want_t, want_a = oracle_fr3.rnea_batch(q, dq, ddq), oracle_fr3.forward_dynamics_batch(q, dq, tau)
    ax = 0
    if layout == "aos":
        q, dq, ddq, tau, want_t, want_a = (np.ascontiguousarray(x.T) for x in (q, dq, ddq, tau, want_t, want_a))
        ax = 1
    n = 7
    split = lambda x: (x[:n], x[n:]) if layout == "soa" else (x[:, :n], x[:, n:])
    t, a = mb_fr3.rnea(q, dq, ddq, layout=layout), mb_fr3.forward_dynamics(q, dq, tau, layout=layout)
    both = mb_fr3.rnea_fd(q, dq, ddq, tau, layout=layout)
    bt, ba = split(both)
    assert np.array_equal(ba, a)
    assert state_err(bt, t, ax).max() < 1e-12
    assert state_err(bt, want_t, ax).max() < TOL and state_err(ba, want_a, ax).max() < TOL
    dev = torch.device("cuda:0")
    tq, tdq, tddq, ttau = (torch.from_numpy(x).to(dev) for x in (q, dq, ddq, tau))
    got = mb_fr3.rnea_fd(tq, tdq, tddq, ttau, layout=layout)
    mb_fr3.sync()
    assert np.array_equal(got.cpu().numpy(), both)
    sl = (lambda x: x[0]) if layout == "aos" else (lambda x: x[:, 0])
    t1, a1 = mb_fr3.rnea_fd(sl(q), sl(dq), sl(ddq), sl(tau))
    assert np.array_equal(t1, sl(bt)) and np.array_equal(a1, sl(ba))
    os.environ["RIGIDBODY_B200_FUSED"] = "0"
    try:
        two = rb.Multibody.from_urdf(FR3)
    finally:
        os.environ.pop("RIGIDBODY_B200_FUSED", None)
    assert np.array_equal(two.rnea_fd(q, dq, ddq, tau, layout=layout), np.concatenate([t, a], axis=ax))


def test_fused_rnea_fd_all_families(rb, oracle_fr3):
    """The fused inverse + forward dynamics pass in every family that has it (ahead-of-time, run-time compiled,
    run-time constants), ragged size; families without it (generic-n) issue the two launches."""
    B = 1000
    q, dq, ddq, tau = _states(oracle_fr3, B, seed=0x5EED0007)
    want_t, want_a = oracle_fr3.rnea_batch(q, dq, ddq), oracle_fr3.forward_dynamics_batch(q, dq, tau)
    for mb in _variants(rb, FR3):
        both = mb.rnea_fd(q, dq, ddq, tau)
        assert state_err(both[:7], want_t, 0).max() < TOL, mb.kernel_variant
        assert state_err(both[7:], want_a, 0).max() < TOL, mb.kernel_variant
        assert np.array_equal(both[7:], mb.forward_dynamics(q, dq, tau)), mb.kernel_variant


def test_empty_batch_and_bad_shapes(mb_fr3):
    z = np.zeros((7, 0))
    assert mb_fr3.rnea(z, z, z).shape == (7, 0)
    with pytest.raises(ValueError):
        mb_fr3.rnea(np.zeros((6, 4)), np.zeros((6, 4)), np.zeros((6, 4)))
    with pytest.raises(ValueError):
        mb_fr3.rnea(np.zeros(6), np.zeros(6), np.zeros(6))


def test_ragged_leading_dimension_device(rb, mb_fr3, oracle_fr3):
    """SoA with ld > n_states through the raw C ABI on device pointers (views into a wider allocation)."""
    import torch
    from rigidbody_rs_b200 import _lib
    B, ld = 1000, 1536
    q, dq, ddq, _ = _states(oracle_fr3, B)
    dev = torch.device("cuda:0")
    bufs = [torch.full((7, ld), float("nan"), dtype=torch.float64, device=dev) for _ in range(4)]
    for b, x in zip(bufs, (q, dq, ddq)):
        b[:, :B] = torch.from_numpy(x).to(dev)
    torch.cuda.synchronize()
    rc = _lib.lib.multibody_rnea_batch(mb_fr3._h, *[C.c_void_p(b.data_ptr()) for b in bufs], B, ld,
                                       _lib.RB_LAYOUT_SOA, _lib.RB_MEM_DEVICE, None)
    assert rc == 0, _lib.lib.multibody_last_error()
    mb_fr3.sync()
    out = bufs[3].cpu().numpy()
    assert state_err(out[:, :B], oracle_fr3.rnea_batch(q, dq, ddq), 0).max() < TOL
    assert np.isnan(out[:, B:]).all()                       # padding untouched
    # ld < n_states is rejected
    rc = _lib.lib.multibody_rnea_batch(mb_fr3._h, *[C.c_void_p(b.data_ptr()) for b in bufs], B, B - 1,
                                       _lib.RB_LAYOUT_SOA, _lib.RB_MEM_DEVICE, None)
    assert rc == _lib.RB_ERR_ARG


def test_ragged_leading_dimension_rollout_and_kinematics(rb, oracle_fr3):
    """Memory safety of the remaining kernels without compute-sanitizer (closed on this pool): ld > n with NaN-poisoned
    padding through the rollout, the fused rollout cost, the fused inverse + forward dynamics pass, crba, jac and fwd_kin,
    in every kernel family and at sizes that leave ragged warps and blocks -- results equal the compact call bit for bit,
    every padding element still NaN."""
    import torch
    from rigidbody_rs_b200 import _lib
    lib = _lib.lib
    dev = torch.device("cuda:0")
    nan = float("nan")
    H, dt = 5, 1e-3
    lim = oracle_fr3.model
    for mb in _variants(rb, FR3):
        for B, ld in ((1, 8), (33, 64), (129, 200)):
            q, dq, ddq, tin = _states(oracle_fr3, B, seed=0x5EED0021)
            tau = np.stack([oracle_fr3.fill(0x5EED0021, 4 + t, -lim.effort, lim.effort, t * B, B) for t in range(H)])

            def wide(x):                                   # [..., n, B] -> NaN-padded [..., n, ld] on the device
                w = torch.full(x.shape[:-1] + (ld,), nan, dtype=torch.float64, device=dev)
                w[..., :B] = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
                return w
            p = lambda t: C.c_void_p(t.data_ptr())
            wq, wdq, wddq, wtin, wtau = wide(q), wide(dq), wide(ddq), wide(tin), wide(tau)
            blank = lambda *shape: torch.full(shape, nan, dtype=torch.float64, device=dev)
            # rollout: trajectory + final state
            qt, dqt, qf, dqf = blank(H, 7, ld), blank(H, 7, ld), blank(7, ld), blank(7, ld)
            assert lib.multibody_rollout(mb._h, p(wq), p(wdq), p(wtau), dt, H, p(qt), p(dqt), p(qf), p(dqf), B, ld, 0, 1, None) == 0
            # fused cost
            w = np.linspace(0.5, 2.0, 7)
            qc = _lib.RbQuadCost()
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
            qc.w_q, qc.w_tau, qc.w_q_final = dp(w), dp(w), dp(w)
            cost = blank(ld)
            assert lib.multibody_rollout_cost(mb._h, p(wq), p(wdq), p(wtau), dt, H, C.byref(qc), p(cost), None, None, B, ld, 0, 1, None) == 0
            # fused inverse + forward dynamics, crba, jac, fwd_kin
            both, Hm, J, xyz = blank(14, ld), blank(49, ld), blank(42, ld), blank(3, ld)
            assert lib.multibody_rnea_fd_batch(mb._h, p(wq), p(wdq), p(wddq), p(wtin), p(both), B, ld, 0, 1, None) == 0
            assert lib.multibody_crba_batch(mb._h, p(wq), p(Hm), B, ld, 0, 1, None) == 0
            assert lib.multibody_jac_batch(mb._h, p(wq), p(J), B, ld, 0, 1, None) == 0
            assert lib.multibody_fwd_kin_batch(mb._h, p(wq), p(xyz), B, ld, 0, 1, None) == 0
            mb.sync()
            ref_qt, ref_dqt, ref_qf, ref_dqf = mb.rollout(q, dq, tau, dt, final=True)
            ref_cost = mb.rollout_cost(q, dq, tau, dt, w_q=w, w_tau=w, w_q_final=w)
            for got, want in ((qt, ref_qt), (dqt, ref_dqt), (qf, ref_qf), (dqf, ref_dqf), (cost, ref_cost), (both, mb.rnea_fd(q, dq, ddq, tin)),
                              (Hm, mb.crba(q)), (J, mb.jac(q)), (xyz, mb.fwd_kin(q))):
                g = got.cpu().numpy()
                np.testing.assert_array_equal(g[..., :B], np.asarray(want).reshape(g[..., :B].shape), err_msg=f"{mb.kernel_variant} B={B}")
                assert np.isnan(g[..., B:]).all(), (mb.kernel_variant, B)


def test_ragged_leading_dimension_long_chain_and_derivatives(rb, mb_fr3, mb_chain32, oracle_chain32):
    """ld > n_states with NaN-poisoned padding through the lane-per-joint FD kernel (odd tail: half- and quarter-warp state slots, staging
    groups) and the derivative kernels: results right, padding untouched."""
    import torch
    from rigidbody_rs_b200 import _lib
    from oracle.rb_oracle_np import ChainNP
    dev = torch.device("cuda:0")
    nan = float("nan")
    for B in (1, 2, 3, 5, 127):
        ld = B + 9
        rng = np.random.default_rng(B)
        q, dq, tau = rng.uniform(-3, 3, (32, B)), rng.uniform(-1, 1, (32, B)), rng.uniform(-20, 20, (32, B))
        bufs = [torch.full((32, ld), nan, dtype=torch.float64, device=dev) for _ in range(4)]
        for b, x in zip(bufs, (q, dq, tau)):
            b[:, :B] = torch.from_numpy(x).to(dev)
        torch.cuda.synchronize()
        rc = _lib.lib.multibody_forward_dynamics_batch(mb_chain32._h, *[C.c_void_p(b.data_ptr()) for b in bufs], B, ld,
                                                       _lib.RB_LAYOUT_SOA, _lib.RB_MEM_DEVICE, None)
        assert rc == 0, _lib.lib.multibody_last_error()
        mb_chain32.sync()
        out = bufs[3].cpu().numpy()
        assert state_err(out[:, :B], oracle_chain32.forward_dynamics_batch(q, dq, tau), 0).max() < TOL
        assert np.isnan(out[:, B:]).all()
    B, ld = 37, 64
    rng = np.random.default_rng(1)
    q, dq, ddq = rng.uniform(-2, 2, (3, 7, B))
    ins = [torch.full((7, ld), nan, dtype=torch.float64, device=dev) for _ in range(3)]
    for b, x in zip(ins, (q, dq, ddq)):
        b[:, :B] = torch.from_numpy(x).to(dev)
    out = torch.full((147, ld), nan, dtype=torch.float64, device=dev)
    rc = _lib.lib.multibody_fd_derivatives_batch(mb_fr3._h, *[C.c_void_p(b.data_ptr()) for b in ins], C.c_void_p(out.data_ptr()), B, ld,
                                                 _lib.RB_LAYOUT_SOA, _lib.RB_MEM_DEVICE, None)
    assert rc == 0, _lib.lib.multibody_last_error()
    mb_fr3.sync()
    o = out.cpu().numpy()
    assert np.isnan(o[:, B:]).all() and np.isfinite(o[:, :B]).all()
    want = mb_fr3.fd_derivatives(np.ascontiguousarray(q), np.ascontiguousarray(dq), np.ascontiguousarray(ddq))
    np.testing.assert_array_equal(o[:, :B], want)


def test_device_tensors_all_ops_medium_batch(rb, oracle_fr3):
    import torch
    B = 200_000
    q, dq, ddq, tau = _states(oracle_fr3, B)
    dev = torch.device("cuda:0")
    tq, tdq, tddq, ttau = (torch.from_numpy(x).to(dev) for x in (q, dq, ddq, tau))
    want_t = oracle_fr3.rnea_batch(q, dq, ddq)
    want_a = oracle_fr3.forward_dynamics_batch(q, dq, tau)
    want_H = oracle_fr3.crba_batch(q[:, :20000])
    for mb in _variants(rb, FR3):
        assert state_err(mb.rnea(tq, tdq, tddq).cpu().numpy(), want_t, 0).max() < TOL, mb.kernel_variant
        assert state_err(mb.forward_dynamics(tq, tdq, ttau).cpu().numpy(), want_a, 0).max() < TOL, mb.kernel_variant
        assert state_err(mb.crba(tq[:, :20000].contiguous()).cpu().numpy(), want_H, 0).max() < TOL
        fk = mb.fwd_kin(tq[:, :2000].contiguous()).cpu().numpy()
        jc = mb.jac(tq[:, :2000].contiguous()).cpu().numpy()
        for s in range(0, 2000, 97):
            assert np.abs(fk[:, s] - oracle_fr3.fwd_kin(q[:, s])).max() < TOL
            assert np.abs(jc[:, s].reshape(7, 6).T - oracle_fr3.jac(q[:, s])).max() < TOL


def test_aos_device_batches_equal_soa_bitwise(rb, oracle_fr3):
    """AoS [B][7] device tensors go through the in-kernel shared-memory staging (no transposes) for the unrolled
    families and must give exactly the SoA kernel's numbers, including a ragged last block."""
    import torch
    B = 100_003
    q, dq, ddq, tau = _states(oracle_fr3, B)
    dev = torch.device("cuda:0")
    soa = [torch.from_numpy(x).to(dev) for x in (q, dq, ddq, tau)]
    aos = [t.t().contiguous() for t in soa]
    for mb in _variants(rb, FR3):
        l0 = mb.launch_count
        t_aos = mb.rnea(aos[0], aos[1], aos[2], layout="aos")
        launches = mb.launch_count - l0
        t_soa = mb.rnea(soa[0], soa[1], soa[2])
        a_aos = mb.forward_dynamics(aos[0], aos[1], aos[3], layout="aos")
        a_soa = mb.forward_dynamics(soa[0], soa[1], soa[3])
        mb.sync()
        assert torch.equal(t_aos.t(), t_soa) and torch.equal(a_aos.t(), a_soa), mb.kernel_variant
        assert launches == (5 if mb.kernel_variant == "generic-n" else 1), (mb.kernel_variant, launches)
    assert state_err(t_soa.cpu().numpy(), oracle_fr3.rnea_batch(q, dq, ddq), 0).max() < TOL


def test_device_sampler_bit_identical_to_oracle(mb_fr3, oracle_fr3):
    import torch
    lim = mb_fr3.limits()
    out = torch.empty((7, 4096), dtype=torch.float64, device="cuda:0")
    mb_fr3.fill(out, 0x5EED0001, 0, lim["lower"], lim["upper"], first_index=123456)
    mb_fr3.sync()
    want = oracle_fr3.fill(0x5EED0001, 0, lim["lower"], lim["upper"], 123456, 4096)
    np.testing.assert_array_equal(out.cpu().numpy(), want)              # bit-exact: integer mixing + one fma


def test_rollout_matches_oracle(rb, oracle_fr3):
    B, H, dt = 64, 64, 1e-3
    q, dq, _, _ = _states(oracle_fr3, B, seed=0x5EED0003)
    lim = oracle_fr3.model
    tau = np.stack([oracle_fr3.fill(0x5EED0003, 4 + t % 32, -lim.effort, lim.effort, t * B, B) for t in range(H)])
    oq, odq = oracle_fr3.rollout_batch(q, dq, tau, dt)
    for mb in _variants(rb, FR3):
        qt, dqt, qf, dqf = mb.rollout(q, dq, tau, dt, final=True)
        assert state_err(qt, oq, 1).max() < TOL and state_err(dqt, odq, 1).max() < TOL, mb.kernel_variant
        np.testing.assert_array_equal(qf, qt[-1]); np.testing.assert_array_equal(dqf, dqt[-1])
        # AoS front end gives the same numbers
        qa, dqa = mb.rollout(np.ascontiguousarray(q.T), np.ascontiguousarray(dq.T),
                             np.ascontiguousarray(tau.transpose(0, 2, 1)), dt, layout="aos")
        np.testing.assert_array_equal(qa.transpose(0, 2, 1), qt)


@pytest.mark.parametrize("B", [1, 33, 100, 4096])
def test_rollout_launch_modes_agree_bitwise(rb, oracle_fr3, B):
    """$RIGIDBODY_B200_ROLLOUT=ws selects the two-warp kernel of experiments/rb_rollout_ws.cuh in builds that have it
    (-DRB_ROLLOUT_WS=1; ignored otherwise): bit-identical trajectories and costs, ragged tails included."""
    H, dt = 24, 1e-3
    q, dq, _, _ = _states(oracle_fr3, B, seed=0x5EED0003)
    lim = oracle_fr3.model
    tau = np.stack([oracle_fr3.fill(0x5EED0003, 4 + t % 32, -lim.effort, lim.effort, t * B, B) for t in range(H)])
    w = np.linspace(0.5, 2.0, 7)
    for mb in _variants(rb, FR3):
        if mb.kernel_variant == "generic-n":
            continue
        a = mb.rollout(q, dq, tau, dt, final=True)
        ca = mb.rollout_cost(q, dq, tau, dt, w_q=w, w_dq=0.1 * w, w_tau=1e-3 * w, w_q_final=3 * w)
        for mode in ("ws",):
            os.environ["RIGIDBODY_B200_ROLLOUT"] = mode
            try:
                b = mb.rollout(q, dq, tau, dt, final=True)
                cb = mb.rollout_cost(q, dq, tau, dt, w_q=w, w_dq=0.1 * w, w_tau=1e-3 * w, w_q_final=3 * w)
            finally:
                os.environ.pop("RIGIDBODY_B200_ROLLOUT", None)
            for x, y in zip(a, b):
                np.testing.assert_array_equal(x, y, err_msg=mb.kernel_variant + " " + mode)
            np.testing.assert_array_equal(ca, cb, err_msg=mb.kernel_variant + " " + mode)


def test_rollout_is_forward_dynamics_plus_fma_euler_bitwise(mb_fr3, oracle_fr3):
    """The rollout kernel performs the arithmetic of the forward-dynamics kernel: stepping with
    multibody_forward_dynamics_batch and an exactly rounded fma-Euler update reproduces it bit for bit."""
    import mpmath as mp
    mp.mp.prec = 200
    B, H, dt = 33, 5, 1e-3
    q, dq, _, _ = _states(oracle_fr3, B, seed=0x5EED0003)
    lim = oracle_fr3.model
    tau = np.stack([oracle_fr3.fill(0x5EED0003, 4 + t % 32, -lim.effort, lim.effort, t * B, B) for t in range(H)])
    qt, dqt = mb_fr3.rollout(q, dq, tau, dt)
    fma = np.vectorize(lambda a, b, c: float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c)))
    qs, dqs = q.copy(), dq.copy()
    for t in range(H):
        qdd = mb_fr3.forward_dynamics(qs, dqs, tau[t])
        dqs = fma(dt, qdd, dqs)
        qs = fma(dt, dqs, qs)
        np.testing.assert_array_equal(dqt[t], dqs)
        np.testing.assert_array_equal(qt[t], qs)


def test_rollout_chain32(rb, mb_chain32, oracle_chain32):
    """Long-chain rollouts: one forward-dynamics launch (lane-per-joint kernel) + one integration launch per step."""
    B, H, dt = 70, 12, 1e-3
    rng = np.random.default_rng(5)
    q, dq = rng.uniform(-3, 3, (32, B)), rng.uniform(-1, 1, (32, B))
    tau = rng.uniform(-20, 20, (H, 32, B))
    oq, odq = oracle_chain32.rollout_batch(q, dq, tau, dt)
    qt, dqt, qf, dqf = mb_chain32.rollout(q, dq, tau, dt, final=True)
    assert state_err(qt, oq, 1).max() < TOL and state_err(dqt, odq, 1).max() < TOL
    np.testing.assert_array_equal(qf, qt[-1]); np.testing.assert_array_equal(dqf, dqt[-1])
    w = np.linspace(0.5, 2.0, 32)
    c = mb_chain32.rollout_cost(q, dq, tau, dt, w_q=w, w_dq=0.1 * w, w_tau=1e-3 * w, w_q_final=3 * w)
    want = ((w[None, :, None] * qt ** 2).sum(1) + (0.1 * w[None, :, None] * dqt ** 2).sum(1) + (1e-3 * w[None, :, None] * tau ** 2).sum(1)).sum(0) * dt \
        + (3 * w[:, None] * qt[-1] ** 2).sum(0)
    np.testing.assert_allclose(c, want, rtol=1e-12)


@pytest.mark.parametrize("n,seed", [(1, 91), (2, 92), (6, 93)])
def test_runtime_n_family_on_short_chains(rb, n, seed):
    """What a chain gets when the run-time compiler is unavailable: the run-time-n kernels, including the lane-per-joint
    forward dynamics with most lanes idle."""
    from test_host import _random_chain
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, seed)
    os.environ["RIGIDBODY_B200_VARIANT"] = "generic-n"
    try:
        mb = rb.Multibody.from_descriptor(R, t, m, c, Ic)
    finally:
        os.environ.pop("RIGIDBODY_B200_VARIANT", None)
    assert mb.kernel_variant == "generic-n"
    ch = ChainNP.from_arrays(R, t, m, c, Ic)
    rng = np.random.default_rng(seed)
    B = 1027
    q, dq, ddq = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n))
    tau = ch.rnea(q, dq, ddq)
    assert state_err(mb.rnea(q, dq, ddq, layout="aos"), tau, 1).max() < TOL
    assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), ddq, 1).max() < TOL


def test_stepwise_rollout_in_chunks(rb, mb_fr3):
    """The run-time-n rollout carries its state in engine scratch; more trajectories than fit are processed in chunks."""
    import torch
    os.environ["RIGIDBODY_B200_VARIANT"] = "generic-n"
    try:
        mb = rb.Multibody.from_urdf(FR3)
    finally:
        os.environ.pop("RIGIDBODY_B200_VARIANT", None)
    B, H = 1_000_003, 2                                   # scratch holds ~720 k 7-joint trajectories at once
    dev = torch.device("cuda:0")
    lim = mb.limits()
    q = torch.empty((7, B), dtype=torch.float64, device=dev); dq = torch.empty_like(q)
    tau = torch.empty((H, 7, B), dtype=torch.float64, device=dev)
    mb.fill(q, 7, 0, lim["lower"], lim["upper"]); mb.fill(dq, 7, 1, -lim["velocity"], lim["velocity"])
    for t in range(H):
        mb.fill(tau[t], 7, 2 + t, -lim["effort"], lim["effort"])
    got = mb.rollout(q, dq, tau, 1e-3, final=True)
    want = mb_fr3.rollout(q, dq, tau, 1e-3, final=True)
    mb.sync(); mb_fr3.sync()
    for g, w in zip(got, want):
        assert float((g - w).abs().max()) < 1e-10


def test_rollout_cost_matches_trajectory_cost(rb, oracle_fr3):
    """The fused rollout+cost kernel returns exactly the quadratic cost of the trajectory the oracle integrates."""
    import torch
    B, H, dt = 96, 32, 2e-3
    q, dq, _, _ = _states(oracle_fr3, B, seed=0x5EED0003)
    lim = oracle_fr3.model
    tau = np.stack([oracle_fr3.fill(0x5EED0003, 4 + t % 32, -lim.effort, lim.effort, t * B, B) for t in range(H)])
    oq, odq = oracle_fr3.rollout_batch(q, dq, tau, dt)
    rng = np.random.default_rng(3)
    q_ref = rng.uniform(-1, 1, 7); w_q, w_dq, w_tau = rng.uniform(0, 2, 7), rng.uniform(0, 1, 7), rng.uniform(0, 1e-3, 7)
    w_qf, w_dqf = rng.uniform(0, 50, 7), rng.uniform(0, 5, 7)
    run = ((w_q[None, :, None] * (oq - q_ref[None, :, None]) ** 2).sum(1) + (w_dq[None, :, None] * odq ** 2).sum(1)
           + (w_tau[None, :, None] * tau ** 2).sum(1)).sum(0) * dt
    want = run + (w_qf[:, None] * (oq[-1] - q_ref[:, None]) ** 2).sum(0) + (w_dqf[:, None] * odq[-1] ** 2).sum(0)
    kw = dict(q_ref=q_ref, w_q=w_q, w_dq=w_dq, w_tau=w_tau, w_q_final=w_qf, w_dq_final=w_dqf)
    for mb in _variants(rb, FR3):
        c_host, qf, dqf = mb.rollout_cost(q, dq, tau, dt, final=True, **kw)
        np.testing.assert_allclose(c_host, want, rtol=1e-10, err_msg=mb.kernel_variant)
        assert state_err(qf, oq[-1], 0).max() < TOL
        dev = torch.device("cuda:0")
        c_dev = mb.rollout_cost(torch.from_numpy(q).to(dev), torch.from_numpy(dq).to(dev), torch.from_numpy(tau).to(dev), dt, **kw)
        np.testing.assert_array_equal(c_dev.cpu().numpy(), c_host)
        c_aos = mb.rollout_cost(np.ascontiguousarray(q.T), np.ascontiguousarray(dq.T), np.ascontiguousarray(tau.transpose(0, 2, 1)),
                                dt, layout="aos", **kw)
        np.testing.assert_array_equal(c_aos, c_host)
    with pytest.raises(rb.RigidBodyError):
        mb.rollout_cost(q, dq, tau, dt, w_q=-np.ones(7))


def test_chain32_medium_batch(rb, mb_chain32, oracle_chain32):
    B = 4096
    o = oracle_chain32
    q = o.fill(0x5EED0005, 0, -np.pi, np.pi, 0, B); dq = o.fill(0x5EED0005, 1, -2.0, 2.0, 0, B)
    ddq = o.fill(0x5EED0005, 2, -10.0, 10.0, 0, B); tau = o.fill(0x5EED0005, 3, -50.0, 50.0, 0, B)
    assert state_err(mb_chain32.rnea(q, dq, ddq), o.rnea_batch(q, dq, ddq), 0).max() < TOL
    # cond(H) reaches ~3e4 on this chain; measured errors stay below 3e-12 (profiles/r2_error_budget.jsonl): the 1e-10 bar holds
    assert state_err(mb_chain32.forward_dynamics(q, dq, tau), o.forward_dynamics_batch(q, dq, tau), 0).max() < TOL
    t = mb_chain32.rnea(q, dq, ddq)
    assert state_err(mb_chain32.forward_dynamics(q, dq, t), ddq, 0).max() < TOL
    # the run-time-n family on the same chain, plus fwd_kin / jac / rollout which the long-chain family delegates to it
    gen = [m for m in _variants(rb, CHAIN32) if m.kernel_variant == "generic-n"][0]
    assert state_err(gen.rnea(q, dq, ddq), o.rnea_batch(q, dq, ddq), 0).max() < TOL
    assert state_err(gen.forward_dynamics(q, dq, tau), o.forward_dynamics_batch(q, dq, tau), 0).max() < TOL
    fk = mb_chain32.fwd_kin(q[:, :64]); jc = mb_chain32.jac(q[:, :64])
    for s in range(0, 64, 7):
        assert np.abs(fk[:, s] - o.fwd_kin(q[:, s])).max() < TOL
        assert np.abs(jc[:, s].reshape(32, 6).T - o.jac(q[:, s])).max() < TOL
    qt, dqt = mb_chain32.rollout(q[:, :8], dq[:, :8], np.repeat(tau[None, :, :8], 4, axis=0), 1e-3)
    oq, odq = o.rollout_batch(q[:, :8], dq[:, :8], np.repeat(tau[None, :, :8], 4, axis=0), 1e-3)
    assert state_err(qt, oq, 1).max() < TOL and state_err(dqt, odq, 1).max() < TOL


def test_long_chain_fd_is_independent_of_how_the_batch_is_cut(mb_chain32, oracle_chain32):
    """rbq_fd_kernel walks the batch in sweeps of (SMs x 8 warps) groups of 4 states; a warp whose group does not exist
    repeats the grid's last group and stores nothing, a state that does not exist is computed from zeros.  Whatever
    the cut -- one call, two calls at an odd boundary, a tail of 1..3 states, a batch smaller than one sweep -- every
    state's result is the same bits, and right."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    per_sweep = sms * 8 * 4                               # states one sweep of the grid covers
    B = 2 * per_sweep + 4 + 3                             # two full sweeps, then one full group and a 3-state tail
    o = oracle_chain32
    q = o.fill(0x5EED0005, 0, -np.pi, np.pi, 0, B); dq = o.fill(0x5EED0005, 1, -2.0, 2.0, 0, B)
    tau = o.fill(0x5EED0005, 3, -50.0, 50.0, 0, B)
    whole = mb_chain32.forward_dynamics(q, dq, tau)
    assert np.isfinite(whole).all()
    for cut in (1, 5, per_sweep - 1, per_sweep + 2, B - 3, B - 1):
        a = mb_chain32.forward_dynamics(np.ascontiguousarray(q[:, :cut]), np.ascontiguousarray(dq[:, :cut]), np.ascontiguousarray(tau[:, :cut]))
        b = mb_chain32.forward_dynamics(np.ascontiguousarray(q[:, cut:]), np.ascontiguousarray(dq[:, cut:]), np.ascontiguousarray(tau[:, cut:]))
        np.testing.assert_array_equal(np.concatenate([a, b], axis=1), whole, err_msg=f"cut at {cut}")
    idx = np.r_[0:64, per_sweep - 32:per_sweep + 32, B - 40:B]
    want = o.forward_dynamics_batch(np.ascontiguousarray(q[:, idx]), np.ascontiguousarray(dq[:, idx]), np.ascontiguousarray(tau[:, idx]))
    assert state_err(whole[:, idx], want, 0).max() < TOL


def test_jit_equals_ahead_of_time_build_bitwise(rb, oracle_fr3):
    """The FR3 kernels compiled at load time by NVRTC are the kernels nvcc compiled ahead of time: same bits out."""
    import torch
    q, dq, ddq, tau = _states(oracle_fr3, 50_000)
    dev = torch.device("cuda:0")
    tq, tdq, tddq, ttau = (torch.from_numpy(x).to(dev) for x in (q, dq, ddq, tau))
    fam = {m.kernel_variant: m for m in _variants(rb, FR3)}
    a, j = fam["fr3-specialised"], fam["jit-specialised"]
    assert "NVRTC" in j._note() or "cache" in j._note()
    for fn, x3 in (("rnea", tddq), ("forward_dynamics", ttau)):
        assert torch.equal(getattr(a, fn)(tq, tdq, x3), getattr(j, fn)(tq, tdq, x3)), fn
    assert torch.equal(a.crba(tq[:, :4096].contiguous()), j.crba(tq[:, :4096].contiguous()))
    assert torch.equal(a.jac(tq[:, :4096].contiguous()), j.jac(tq[:, :4096].contiguous()))
    ra = a.rollout(tq[:, :512].contiguous(), tdq[:, :512].contiguous(), ttau[:, :512].contiguous().unsqueeze(0).repeat(4, 1, 1), 1e-3)
    rj = j.rollout(tq[:, :512].contiguous(), tdq[:, :512].contiguous(), ttau[:, :512].contiguous().unsqueeze(0).repeat(4, 1, 1), 1e-3)
    assert torch.equal(ra[0], rj[0]) and torch.equal(ra[1], rj[1])


@pytest.mark.parametrize("n,seed", [(1, 5), (3, 6), (5, 7), (7, 8), (10, 9)])
def test_jit_random_chains_match_matrix_oracle(rb, n, seed):
    """Chains the library has never seen (general fixed rotations, 1..10 joints) get run-time specialised kernels;
    parity against the numpy matrix-form oracle (the C oracle takes URDF-style rpy only)."""
    from test_host import _random_chain
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, seed)
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic)
    assert mb.kernel_variant == "jit-specialised", mb._note()
    ch = ChainNP.from_arrays(R, t, m, c, Ic)
    rng = np.random.default_rng(seed)
    B = 2000
    q, dq, ddq, tau = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n)), rng.uniform(-20, 20, (B, n))
    assert state_err(mb.rnea(q, dq, ddq, layout="aos"), ch.rnea(q, dq, ddq), 1).max() < TOL
    assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), ch.forward_dynamics(q, dq, tau), 1).max() < TOL
    H = mb.crba(q[:64], layout="aos").reshape(64, n, n).transpose(0, 2, 1)
    assert np.abs(H - ch.crba(q[:64])).max() < TOL
    assert np.abs(mb.fwd_kin(q[:64], layout="aos") - ch.fwd_kin(q[:64])[1]).max() < TOL
    J = mb.jac(q[:64], layout="aos").reshape(64, n, 6).transpose(0, 2, 1)
    assert np.abs(J - ch.jac(q[:64])).max() < TOL


def _unpack(D, n, blocks):
    """[B, blocks*n*n] (aos packing of the C ABI: block b, entry r + n*c) -> list of [B, n, n] with [b, r, c]."""
    B = D.shape[0]
    return [D[:, b * n * n:(b + 1) * n * n].reshape(B, n, n).transpose(0, 2, 1) for b in range(blocks)]


def test_analytical_derivatives_fr3_all_families(rb, oracle_fr3):
    """d tau / d (q, dq) and d qdd / d (q, dq, tau) against complex-step differentiation of the
    twin's link-frame rnea.  Bar: |err| <= 1e-10 * max(1, max|ref|) per state, for every block."""
    from oracle.rb_oracle_np import ChainNP
    ch = ChainNP(oracle_fr3.model)
    B = 300
    q, dq, ddq, tau = (x.T.copy() for x in _states(oracle_fr3, B))
    Dq, Dv = ch.rnea_derivatives(q, dq, ddq)
    Aq, Av, Mi = ch.fd_derivatives(q, dq, tau)
    fams = []
    for mb in _variants(rb, FR3):                # the derivative kernels are model-agnostic: same code under every family
        fams.append(mb.kernel_variant)
        gq, gv = _unpack(mb.rnea_derivatives(q, dq, ddq, layout="aos"), 7, 2)
        for got, want in ((gq, Dq), (gv, Dv)):
            err = np.abs(got - want).reshape(B, -1).max(1) / np.maximum(1.0, np.abs(want).reshape(B, -1).max(1))
            assert err.max() < TOL, (mb.kernel_variant, err.max())
        fq, fv, fm = _unpack(mb.fd_derivatives(q, dq, tau, layout="aos"), 7, 3)
        for got, want in ((fq, Aq), (fv, Av), (fm, Mi)):
            err = np.abs(got - want).reshape(B, -1).max(1) / np.maximum(1.0, np.abs(want).reshape(B, -1).max(1))
            assert err.max() < TOL, (mb.kernel_variant, err.max())
    assert fams == ["fr3-specialised", "jit-specialised", "generic-7", "generic-n"]
    with pytest.raises(rb.RigidBodyError):       # forward-dynamics derivatives beyond 12 joints: unsupported, not wrong
        z = np.zeros((2, 32))
        rb.Multibody.from_urdf(CHAIN32).fd_derivatives(z, z, z, layout="aos")
    # device SoA tensors, and one state
    import torch
    mb = rb.Multibody.from_urdf(FR3)
    dev = torch.device("cuda:0")
    tq, tdq, tddq = (torch.from_numpy(np.ascontiguousarray(x.T)).to(dev) for x in (q, dq, ddq))
    D = mb.rnea_derivatives(tq, tdq, tddq)
    mb.sync()
    assert tuple(D.shape) == (98, B)
    gq, gv = _unpack(D.cpu().numpy().T.copy(), 7, 2)
    assert np.abs(gq - Dq).max() < TOL * max(1.0, np.abs(Dq).max()) and np.abs(gv - Dv).max() < TOL * max(1.0, np.abs(Dv).max())
    one_q, one_v = mb.rnea_derivatives(q[0], dq[0], ddq[0])
    assert np.abs(one_q - Dq[0]).max() < TOL * max(1.0, np.abs(Dq[0]).max()) and np.abs(one_v - Dv[0]).max() < TOL * max(1.0, np.abs(Dv[0]).max())
    # d tau / d ddq is the mass matrix: consistency of the three blocks with a finite step of the kernels themselves
    e = 1e-6 * np.random.default_rng(0).normal(size=q.shape)
    lin = np.einsum("brc,bc->br", Dq, e)
    # (a finite step: the bar is its O(|e|^2) truncation error, not a parity tolerance)
    assert np.abs(mb.rnea(q + e, dq, ddq, layout="aos") - mb.rnea(q, dq, ddq, layout="aos") - lin).max() < 1e-8


@pytest.mark.parametrize("n,seed", [(2, 31), (5, 32), (10, 33)])
def test_analytical_derivatives_random_chains(rb, n, seed):
    from test_host import _random_chain, AXES
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, seed)
    ax = (AXES + AXES)[:n]
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic, axis=ax)
    assert mb.kernel_variant == "jit-specialised"
    ch = ChainNP.from_arrays(R, t, m, c, Ic, axis=ax)
    rng = np.random.default_rng(seed)
    B = 200
    q, dq, ddq = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n))
    Dq, Dv = ch.rnea_derivatives(q, dq, ddq)
    gq, gv = _unpack(mb.rnea_derivatives(q, dq, ddq, layout="aos"), n, 2)
    assert np.abs(gq - Dq).max() < TOL * max(1.0, np.abs(Dq).max())
    assert np.abs(gv - Dv).max() < TOL * max(1.0, np.abs(Dv).max())
    tau = ch.rnea(q, dq, ddq)
    Aq, Av, Mi = ch.fd_derivatives(q, dq, tau)
    fq, fv, fm = _unpack(mb.fd_derivatives(q, dq, tau, layout="aos"), n, 3)
    for got, want in ((fq, Aq), (fv, Av), (fm, Mi)):
        assert np.abs(got - want).max() < TOL * max(1.0, np.abs(want).max())


def test_inverse_dynamics_derivatives_chain32(rb, mb_chain32, oracle_chain32):
    """The rolled derivative recursion keeps no per-joint state, so it serves long chains too (inverse dynamics only)."""
    from oracle.rb_oracle_np import ChainNP
    ch = ChainNP(oracle_chain32.model)
    rng = np.random.default_rng(8)
    B = 24
    q, dq, ddq = rng.uniform(-3, 3, (B, 32)), rng.uniform(-1, 1, (B, 32)), rng.uniform(-5, 5, (B, 32))
    Dq, Dv = ch.rnea_derivatives(q, dq, ddq)
    gq, gv = _unpack(mb_chain32.rnea_derivatives(q, dq, ddq, layout="aos"), 32, 2)
    assert np.abs(gq - Dq).max() < TOL * max(1.0, np.abs(Dq).max())
    assert np.abs(gv - Dv).max() < TOL * max(1.0, np.abs(Dv).max())


@pytest.mark.parametrize("n,seed", [(13, 1), (15, 7), (19, 6), (24, 2), (32, 3), (33, 4), (64, 5)])
def test_long_random_chains(rb, n, seed):
    """Chains beyond the register-resident limit: run-time-n kernels; forward dynamics by the warp-per-state kernel
    up to 32 joints (idle lanes padded) and by the shared-memory tile solver beyond."""
    from test_host import _random_chain
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, seed)
    t = t * (8.0 / n)                                    # keep the reach (and cond(H)) comparable across lengths
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic)
    # <= 18 joints: every kernel specialised at load time; 19..32: rnea / crba / fwd_kin / jac specialised ("jit-long"), forward
    # dynamics through the lane-per-joint kernel; beyond: loop-based kernels and the shared-memory tile solver
    assert mb.kernel_variant == ("jit-specialised" if n <= 18 else ("jit-long" if n <= 32 else "generic-n"))
    ch = ChainNP.from_arrays(R, t, m, c, Ic)
    rng = np.random.default_rng(seed)
    B = 333
    q, dq, ddq = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n))
    tau = ch.rnea(q, dq, ddq)
    assert state_err(mb.rnea(q, dq, ddq, layout="aos"), tau, 1).max() < TOL
    assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), ddq, 1).max() < TOL      # round trip
    qs, dqs = np.ascontiguousarray(q.T), np.ascontiguousarray(dq.T)
    got = mb.forward_dynamics(qs, dqs, np.ascontiguousarray(tau.T))                            # SoA, ragged tail of a group
    assert state_err(got, ddq.T, 0).max() < TOL
    H = mb.crba(q[:32], layout="aos").reshape(32, n, n).transpose(0, 2, 1)
    assert np.abs(H - ch.crba(q[:32])).max() < TOL * np.abs(H).max()
    assert np.abs(mb.fwd_kin(q[:32], layout="aos") - ch.fwd_kin(q[:32])[1]).max() < TOL
    J = mb.jac(q[:32], layout="aos").reshape(32, n, 6).transpose(0, 2, 1)
    assert np.abs(J - ch.jac(q[:32])).max() < TOL


@pytest.mark.parametrize("n,seed,variant", [(5, 1, "auto"), (9, 2, "auto"), (12, 5, "auto"), (5, 1, "generic-n"), (9, 2, "generic-n"),
                                            (20, 3, "auto"), (40, 4, "auto")])
def test_kinematic_trees(rb, n, seed, variant):
    """Branching trees (parent[i] < i): up to 18 joints the unrolled templates follow the compile-time parent table
    (run-time specialised), beyond that the run-time-n kernels with the parent-indexed recursions; the reference has only
    the serial case (multibody.rs:148,165), so the checker is the twin's tree form (pinned by energy identities in
    tests/test_host.py).  Oblique joint axes ride along."""
    from test_host import _random_chain, _random_tree
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, 70 + seed)
    par = _random_tree(n, seed)
    ax = np.random.default_rng(seed).normal(size=(n, 3))
    ch = ChainNP.from_arrays(R, t, m, c, Ic, axis=ax, parent=par)
    os.environ["RIGIDBODY_B200_VARIANT"] = variant
    try:
        mb = rb.Multibody.from_descriptor(R, t, m, c, Ic, axis=ax, parent=par)
    finally:
        os.environ.pop("RIGIDBODY_B200_VARIANT", None)
    assert mb.kernel_variant == ("jit-specialised" if n <= 18 and variant == "auto" else "generic-n") and "tree" in mb._note()
    rng = np.random.default_rng(seed)
    B = 257
    q, dq, ddq = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n))
    tau = ch.rnea(q, dq, ddq)
    assert state_err(mb.rnea(q, dq, ddq, layout="aos"), tau, 1).max() < TOL
    assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), ddq, 1).max() < TOL
    H = mb.crba(q[:32], layout="aos").reshape(32, n, n).transpose(0, 2, 1)
    Href = ch.crba(q[:32])
    assert np.abs(H - Href).max() < TOL * max(1.0, np.abs(Href).max())
    assert (Href == 0).sum() > 32 * n * (n - 1) // 2                  # branch-induced zeros above the diagonal too
    assert np.abs(mb.fwd_kin(q[:32], layout="aos") - ch.fwd_kin(q[:32])[1]).max() < TOL
    J = mb.jac(q[:32], layout="aos").reshape(32, n, 6).transpose(0, 2, 1)
    assert np.abs(J - ch.jac(q[:32])).max() < TOL
    # three semi-implicit Euler steps against the twin's forward dynamics
    qs, dqs = q[:16].copy(), dq[:16].copy()
    taus = rng.uniform(-5, 5, (3, 16, n))
    for k in range(3):
        a = ch.forward_dynamics(qs, dqs, taus[k])
        dqs = dqs + 1e-3 * a
        qs = qs + 1e-3 * dqs
    qf, dqf = mb.rollout(np.ascontiguousarray(q[:16].T), np.ascontiguousarray(dq[:16].T),
                         np.ascontiguousarray(taus.transpose(0, 2, 1)), 1e-3, trajectory=False, final=True)
    assert np.abs(qf - qs.T).max() < TOL and np.abs(dqf - dqs.T).max() < TOL
    with pytest.raises(rb.RigidBodyError):
        os.environ["RIGIDBODY_B200_VARIANT"] = "generic-7"
        try:
            rb.Multibody.from_descriptor(R, t, m, c, Ic, parent=par)
        finally:
            os.environ.pop("RIGIDBODY_B200_VARIANT", None)
    if n <= 18:                                        # analytical derivatives are for serial chains
        with pytest.raises(rb.RigidBodyError):
            mb.rnea_derivatives(q, dq, ddq, layout="aos")


@pytest.mark.parametrize("n", [7, 4])
def test_general_joint_axes_all_families(rb, n):
    """Joint axes other than +z (x, y, -z, unnormalised, oblique): the loader re-bases the frames, every kernel family
    stays z-only.  Checked against the twin's direct S = (axis; 0) formulation, tip-frame Jacobian included."""
    from test_host import _random_chain, AXES
    from oracle.rb_oracle_np import ChainNP
    R, t, m, c, Ic = _random_chain(n, 40 + n)
    ax = AXES[:n]
    ch = ChainNP.from_arrays(R, t, m, c, Ic, axis=ax)
    rng = np.random.default_rng(n)
    B = 1500
    q, dq, ddq, tau = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n)), rng.uniform(-20, 20, (B, n))
    want = ch.rnea(q, dq, ddq), ch.forward_dynamics(q, dq, tau), ch.crba(q[:64]), ch.fwd_kin(q[:64])[1], ch.jac(q[:64])
    seen = []
    for v in ("auto", "generic-7", "generic-n"):
        os.environ["RIGIDBODY_B200_VARIANT"] = v
        try:
            mb = rb.Multibody.from_descriptor(R, t, m, c, Ic, axis=ax)
        except rb.RigidBodyError:
            continue
        finally:
            os.environ.pop("RIGIDBODY_B200_VARIANT", None)
        seen.append(mb.kernel_variant)
        assert state_err(mb.rnea(q, dq, ddq, layout="aos"), want[0], 1).max() < TOL, mb.kernel_variant
        assert state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), want[1], 1).max() < TOL, mb.kernel_variant
        H = mb.crba(q[:64], layout="aos").reshape(64, n, n).transpose(0, 2, 1)
        assert np.abs(H - want[2]).max() < TOL, mb.kernel_variant
        assert np.abs(mb.fwd_kin(q[:64], layout="aos") - want[3]).max() < TOL, mb.kernel_variant
        J = mb.jac(q[:64], layout="aos").reshape(64, n, 6).transpose(0, 2, 1)
        assert np.abs(J - want[4]).max() < TOL, mb.kernel_variant
    assert seen == (["jit-specialised", "generic-7", "generic-n"] if n == 7 else ["jit-specialised", "generic-n"])


def test_fp32_mode_tolerances(rb, oracle_fr3):
    """Optional fp32 mode (include/rigidbody.h): float kernels against the fp64 oracle on the same (float-rounded)
    inputs.  Stated tolerance (BASELINE.json north_star), per state max_i|x_i - ref_i| <= 1e-4 * max(1, ||ref||_inf)."""
    import torch
    B = 200_000
    q, dq, ddq, tau = _states(oracle_fr3, B)
    f = [np.ascontiguousarray(x, dtype=np.float32) for x in (q, dq, ddq, tau)]
    want_t = oracle_fr3.rnea_batch(*(x.astype(np.float64) for x in f[:3]))
    want_a = oracle_fr3.forward_dynamics_batch(f[0].astype(np.float64), f[1].astype(np.float64), f[3].astype(np.float64))
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(x).to(dev) for x in f]
    for mb in _variants(rb, FR3):
        if mb.kernel_variant == "generic-n":
            with pytest.raises(rb.RigidBodyError):
                mb.rnea(t[0], t[1], t[2])
            continue
        got_t = mb.rnea(t[0], t[1], t[2])
        got_a = mb.forward_dynamics(t[0], t[1], t[3])
        assert got_t.dtype == torch.float32
        e_t = state_err(got_t.cpu().numpy().astype(np.float64), want_t, 0)
        e_a = state_err(got_a.cpu().numpy().astype(np.float64), want_a, 0)
        assert e_t.max() < 1e-4, (mb.kernel_variant, e_t.max())
        assert e_a.max() < 1e-4, (mb.kernel_variant, e_a.max())
        print(mb.kernel_variant, "fp32 errors: rnea", e_t.max(), "fd", e_a.max(), "fd 99.9th pct", np.quantile(e_a, 0.999))


def test_not_spd_is_reported(rb):
    """A physically impossible chain (negative rotational inertia) makes H indefinite: host calls return
    RB_ERR_NOT_SPD and NaN rows instead of garbage."""
    n = 2
    R = np.tile(np.eye(3), (n, 1, 1)); t = np.zeros((n, 3)); t[1] = [0.1, 0, 0]
    m = np.array([1.0, 1.0]); c = np.zeros((n, 3))
    Ic = np.tile(np.diag([0.1, 0.1, -5.0]), (n, 1, 1))
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic)
    z = np.zeros((2, 4))
    with pytest.raises(rb.NotPositiveDefinite):
        mb.forward_dynamics(z, z, z)
    # the same through the fused inverse + forward dynamics pass and both rollout kernels (the two-warp kernel learns it
    # from the matrix warp through shared memory)
    with pytest.raises(rb.NotPositiveDefinite):
        mb.rnea_fd(z, z, z, z)
    tau = np.zeros((3, 2, 4))
    for mode in (None, "ws"):
        if mode:
            os.environ["RIGIDBODY_B200_ROLLOUT"] = mode
        try:
            with pytest.raises(rb.NotPositiveDefinite):
                mb.rollout(z, z, tau, 1e-3)
            with pytest.raises(rb.NotPositiveDefinite):
                mb.rollout_cost(z, z, tau, 1e-3, w_q=np.ones(2))
            # a later, healthy call on the same engine is not blamed for it
            assert mb.rnea(z, z, z).shape == (2, 4)
        finally:
            os.environ.pop("RIGIDBODY_B200_ROLLOUT", None)


def test_nan_and_huge_angles_do_not_poison_neighbours(mb_fr3, oracle_fr3):
    """A NaN state yields NaN for that state only; |q| far outside the fast sincos range still matches the oracle."""
    q, dq, ddq, tau = _states(oracle_fr3, 256)
    q[:, 7] = np.nan
    q[:, 9] = [1.0e6, -3.0e7, 12345.678, -9.9e5, 2.5e8, -1.0e9, 7.7e6]
    t = mb_fr3.rnea(q, dq, ddq)
    assert np.isnan(t[:, 7]).all() and not np.isnan(np.delete(t, 7, axis=1)).any()
    keep = [i for i in range(256) if i != 7]
    assert state_err(t[:, keep], oracle_fr3.rnea_batch(q[:, keep], dq[:, keep], ddq[:, keep]), 0).max() < TOL
    assert state_err(t[:, [0, 1, 2, 100]], oracle_fr3.rnea_batch(q, dq, ddq)[:, [0, 1, 2, 100]], 0).max() < TOL


def test_concurrent_streams_and_threads(rb, mb_fr3, oracle_fr3):
    """Device-pointer calls are re-entrant: two host threads on two CUDA streams share one engine."""
    import threading
    import torch
    B = 300_000
    q, dq, ddq, tau = _states(oracle_fr3, B)
    dev = torch.device("cuda:0")
    tq, tdq, tddq, ttau = (torch.from_numpy(x).to(dev) for x in (q, dq, ddq, tau))
    torch.cuda.synchronize()
    res = {}

    def work(name, fn, a3):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            outs = [fn(tq, tdq, a3) for _ in range(8)]
        st.synchronize()
        res[name] = outs[-1].cpu().numpy()

    th = [threading.Thread(target=work, args=("rnea", mb_fr3.rnea, tddq)),
          threading.Thread(target=work, args=("fd", mb_fr3.forward_dynamics, ttau))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    idx = np.arange(0, B, 997)
    assert state_err(res["rnea"][:, idx], oracle_fr3.rnea_batch(q[:, idx], dq[:, idx], ddq[:, idx]), 0).max() < TOL
    assert state_err(res["fd"][:, idx], oracle_fr3.forward_dynamics_batch(q[:, idx], dq[:, idx], tau[:, idx]), 0).max() < TOL


def test_concurrent_streams_share_engine_scratch_safely(mb_chain32, oracle_chain32):
    """The long-chain and run-time-n families work in engine-owned scratch (H chunk buffer, per-thread scratch);
    calls issued from two threads on two streams must be ordered by the engine, not race."""
    import threading
    import torch
    o, B = oracle_chain32, 3000
    q = o.fill(0x5EED0005, 0, -np.pi, np.pi, 0, B); dq = o.fill(0x5EED0005, 1, -2.0, 2.0, 0, B)
    tau1 = o.fill(0x5EED0005, 3, -50.0, 50.0, 0, B); tau2 = o.fill(0x5EED0006, 3, -50.0, 50.0, 0, B)
    dev = torch.device("cuda:0")
    tq, tdq, t1, t2 = (torch.from_numpy(x).to(dev) for x in (q, dq, tau1, tau2))
    torch.cuda.synchronize()
    res = {}

    def work(name, tt):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            outs = [mb_chain32.forward_dynamics(tq, tdq, tt) for _ in range(6)]
        st.synchronize()
        res[name] = outs[-1].cpu().numpy()

    th = [threading.Thread(target=work, args=("a", t1)), threading.Thread(target=work, args=("b", t2))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert state_err(res["a"], o.forward_dynamics_batch(q, dq, tau1), 0).max() < TOL
    assert state_err(res["b"], o.forward_dynamics_batch(q, dq, tau2), 0).max() < TOL


def test_descriptor_upload_equals_urdf_load(rb, mb_fr3, oracle_fr3):
    """The RbChainDesc path (what the Rust side would call) selects the same specialised kernels and numbers."""
    from oracle.rb_oracle_np import ChainNP
    c = ChainNP(oracle_fr3.model)
    mod = oracle_fr3.model
    Ic = np.stack([[[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]] for s in mod.inertia6])
    mb = rb.Multibody.from_descriptor(c.Rp, mod.xyz, mod.mass, mod.com, Ic)
    assert mb.kernel_variant == "fr3-specialised"       # snapped rotations make the two loads bit-identical
    q, dq, ddq, _ = _states(oracle_fr3, 512)
    assert state_err(mb.rnea(q, dq, ddq), oracle_fr3.rnea_batch(q, dq, ddq), 0).max() < TOL


def test_full_size_properties_16M(mb_fr3, oracle_fr3):
    """BASELINE.json configs[1]/[2] at full size (2^24 states, device-resident, sampled on device):
    FD(q, dq, RNEA(q, dq, ddq)) == ddq for every state, plus a strided oracle sample of both outputs."""
    import torch
    B = 1 << 24
    dev = torch.device("cuda:0")
    lim = mb_fr3.limits()
    q = torch.empty((7, B), dtype=torch.float64, device=dev)
    dq = torch.empty_like(q); ddq = torch.empty_like(q)
    mb_fr3.fill(q, 0x5EED0001, 0, lim["lower"], lim["upper"])
    mb_fr3.fill(dq, 0x5EED0001, 1, -lim["velocity"], lim["velocity"])
    mb_fr3.fill(ddq, 0x5EED0001, 2, -10.0, 10.0)
    tau = mb_fr3.rnea(q, dq, ddq)
    back = mb_fr3.forward_dynamics(q, dq, tau)
    mb_fr3.sync()
    err = (back - ddq).abs().amax(0) / ddq.abs().amax(0).clamp_min(1.0)
    assert float(err.max()) < TOL, float(err.max())
    idx = torch.arange(0, B, 4099, device=dev)
    qs, dqs, ddqs = (x[:, idx].cpu().numpy() for x in (q, dq, ddq))
    assert state_err(tau[:, idx].cpu().numpy(), oracle_fr3.rnea_batch(qs, dqs, ddqs), 0).max() < TOL
    want = oracle_fr3.fill(0x5EED0001, 0, lim["lower"], lim["upper"], 0, 8)
    np.testing.assert_array_equal(q[:, :8].cpu().numpy(), want)
    # configs[2] proper: forward dynamics of the torques BASELINE samples (seed 0x5EED0002), a strided oracle sample of qdd;
    # and the fused inverse + forward dynamics pass on the same 2^24 states equals the two kernels (qdd bit for bit)
    tin = torch.empty_like(q)
    mb_fr3.fill(tin, 0x5EED0002, 3, -lim["effort"], lim["effort"])
    qdd = mb_fr3.forward_dynamics(q, dq, tin)
    both = mb_fr3.rnea_fd(q, dq, ddq, tin)
    mb_fr3.sync()
    tis = tin[:, idx].cpu().numpy()
    assert state_err(qdd[:, idx].cpu().numpy(), oracle_fr3.forward_dynamics_batch(qs, dqs, tis), 0).max() < TOL
    assert bool(torch.equal(both[7:], qdd))
    e = (both[:7] - tau).abs().amax(0) / tau.abs().amax(0).clamp_min(1.0)
    assert float(e.max()) < 1e-12, float(e.max())


def test_full_size_rollout_65536x64(mb_fr3, oracle_fr3):
    """BASELINE.json configs[3] at full size: 65 536 trajectories x 64 steps, dt = 1e-3, inputs sampled on the device as
    bench.py samples them; a strided sample of 1 024 whole trajectories against the oracle's rollout, the final state of
    every trajectory against the last row of its trajectory, and the fused cost against the cost of that trajectory."""
    import torch
    Bt, H, dt = 65536, 64, 1e-3
    dev = torch.device("cuda:0")
    lim = mb_fr3.limits()
    q = torch.empty((7, Bt), dtype=torch.float64, device=dev); dq = torch.empty_like(q)
    mb_fr3.fill(q, 0x5EED0003, 0, lim["lower"], lim["upper"])
    mb_fr3.fill(dq, 0x5EED0003, 1, -lim["velocity"], lim["velocity"])
    tau = torch.empty((H, 7, Bt), dtype=torch.float64, device=dev)
    for t in range(H):
        mb_fr3.fill(tau[t], 0x5EED0003, 4 + t % 32, -lim["effort"], lim["effort"], t * Bt)
    qt, dqt, qf, dqf = mb_fr3.rollout(q, dq, tau, dt, final=True)
    w = np.linspace(0.5, 2.0, 7)
    cost = mb_fr3.rollout_cost(q, dq, tau, dt, w_q=w, w_dq=0.1 * w, w_tau=1e-3 * w, w_q_final=3 * w)
    mb_fr3.sync()
    assert bool(torch.equal(qf, qt[-1])) and bool(torch.equal(dqf, dqt[-1]))
    assert bool(torch.isfinite(qt).all()) and bool(torch.isfinite(dqt).all())
    idx = torch.arange(0, Bt, 64, device=dev)                      # 1 024 trajectories
    oq, odq = oracle_fr3.rollout_batch(q[:, idx].cpu().numpy(), dq[:, idx].cpu().numpy(), tau[:, :, idx].cpu().numpy(), dt)
    assert state_err(qt[:, :, idx].cpu().numpy(), oq, 1).max() < TOL
    assert state_err(dqt[:, :, idx].cpu().numpy(), odq, 1).max() < TOL
    ts = tau[:, :, idx].cpu().numpy()
    want = ((w[None, :, None] * oq ** 2).sum(1) + (0.1 * w[None, :, None] * odq ** 2).sum(1) + (1e-3 * w[None, :, None] * ts ** 2).sum(1)).sum(0) * dt \
        + (3 * w[:, None] * oq[-1] ** 2).sum(0)
    np.testing.assert_allclose(cost[idx].cpu().numpy(), want, rtol=1e-10)


def test_full_size_chain32_4M(mb_chain32, oracle_chain32):
    """BASELINE.json configs[4] (the 32-joint chain) at 2^22 device-sampled states: the forward-dynamics round trip on
    every state, and a strided oracle sample of both tau and qdd held to the bound of conftest.fd_bound."""
    import torch
    from conftest import fd_bound, spd_cond
    B = 1 << 22
    dev = torch.device("cuda:0")
    n = 32
    pi = np.full(n, np.pi)
    q = torch.empty((n, B), dtype=torch.float64, device=dev)
    dq, ddq, tin = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    mb_chain32.fill(q, 0x5EED0005, 0, -pi, pi)
    mb_chain32.fill(dq, 0x5EED0005, 1, -2.0, 2.0)
    mb_chain32.fill(ddq, 0x5EED0005, 2, -10.0, 10.0)
    mb_chain32.fill(tin, 0x5EED0005, 3, -50.0, 50.0)
    tau = mb_chain32.rnea(q, dq, ddq)
    qdd = mb_chain32.forward_dynamics(q, dq, tin)
    back = mb_chain32.forward_dynamics(q, dq, tau)
    mb_chain32.sync()
    idx = torch.arange(0, B, 8191, device=dev)                     # 513 states
    qs, dqs, ddqs, tis = (x[:, idx].cpu().numpy() for x in (q, dq, ddq, tin))
    o = oracle_chain32
    assert state_err(tau[:, idx].cpu().numpy(), o.rnea_batch(qs, dqs, ddqs), 0).max() < TOL
    cond = spd_cond(o.crba_batch(qs), n)
    err = state_err(qdd[:, idx].cpu().numpy(), o.forward_dynamics_batch(qs, dqs, tis), 0)
    assert (err <= 2 * fd_bound(cond)).all(), (err.max(), cond.max())             # both sides carry the bound
    # round trip on every state: the bound with the largest condition number of the sample (x4: the sample is 1 in 8191)
    e = (back - ddq).abs().amax(0) / ddq.abs().amax(0).clamp_min(1.0)
    assert float(e.max()) <= 4 * float(fd_bound(cond.max())), (float(e.max()), cond.max())


def test_cpp_example_runs(rb, tmp_path):
    """examples/batch_demo.cpp through the C ABI from C++: single-state symbols, batched calls, one-call entry point."""
    import subprocess
    from test_host import _build_example
    r = subprocess.run([_build_example(tmp_path), FR3], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fr3-specialised" in r.stdout


def test_forward_dynamics_error_budget_against_mpmath_truth(rb, oracle_fr3, oracle_chain32):
    """Who owns the forward-dynamics error: a 40-digit mpmath solve (oracle/rb_oracle_np.py::fd_mp) is the truth, the
    oracle's LL^T and the CUDA paths (in-register LDL^T with hardware-seeded reciprocals; quarter-warp elimination with
    world-frame sums for the 32-joint chain) are both measured against it.  Every path sits within
    conftest.fd_bound(cond) = max(1e-10, 8 cond(H) eps) -- in fact below 1e-12 (profiles/r2_error_budget.jsonl)."""
    from conftest import fd_bound, spd_cond
    from oracle.rb_oracle_np import fd_mp
    for urdf, o, K, lim in ((FR3, oracle_fr3, 12, 80.0), (CHAIN32, oracle_chain32, 3, 50.0)):
        n = o.model.n
        rng = np.random.default_rng(17)
        q, dq, tau = rng.uniform(-np.pi, np.pi, (K, n)), rng.uniform(-2, 2, (K, n)), rng.uniform(-lim, lim, (K, n))
        truth = np.array([fd_mp(o.model, q[k], dq[k], tau[k]) for k in range(K)])
        cond = spd_cond(o.crba_batch(q, layout="aos").reshape(K, n, n).transpose(0, 2, 1))
        e_or = state_err(o.forward_dynamics_batch(q, dq, tau, layout="aos"), truth, 1)
        assert (e_or <= fd_bound(cond)).all() and e_or.max() < 1e-11
        for mb in _variants(rb, urdf):
            e = state_err(mb.forward_dynamics(q, dq, tau, layout="aos"), truth, 1)
            assert (e <= fd_bound(cond)).all() and e.max() < 1e-11, (mb.kernel_variant, e.max(), cond.max())


def test_rust_crate_double_runs(rb, tmp_path):
    """examples/rust_crate_double.c: the call sequence of rust/rigidbody_gpu_bindings (ChainArrays::from_multibody ->
    multibody_gpu_new -> the batch calls on pinned host buffers -> Drop) in C99 against the same ABI, for the FR3 and for
    the 32-joint chain (run-time-n / long-chain families behind the same calls)."""
    import subprocess
    from test_host import _build_rust_double
    exe = _build_rust_double(tmp_path)
    for urdf, family in ((FR3, "fr3-specialised"), (CHAIN32, "chain32-specialised")):
        r = subprocess.run([exe, urdf], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert family in r.stdout and "all checks passed" in r.stdout


# ------------------------------------------------------------------------------------------------ multi-device engine
def _two_devices():
    """[0, 1] on a multi-GPU box; on a one-GPU box the same device twice (test hook of multibody_gpu_new_multi): the
    dispatcher, the slicing and the worker threads are exercised either way."""
    import torch
    if torch.cuda.device_count() >= 2:
        return [0, 1]
    os.environ["RIGIDBODY_B200_ALLOW_DUPLICATE_DEVICES"] = "1"
    return [0, 0]


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_multi_device_host_batch_equals_single_device_bitwise(rb, mb_fr3, oracle_fr3, layout):
    """multibody_gpu_new_multi: ONE handle, ONE RB_MEM_HOST call, the batch cut into contiguous slices over 2 devices
    (the one-call shape of rigidbody_bindings/src/lib.rs:15-30).  Every result equals the single-device
    engine's bit for bit; ragged sizes, padded leading dimension, sizes below the device count."""
    devs = _two_devices()
    mm = rb.Multibody.from_urdf(FR3, devices=devs)
    assert mm.n_devices == 2 and mm.kernel_variant == "fr3-specialised" and mm.peer(1).device == devs[1]
    ax = 0 if layout == "soa" else 1
    for B in (1, 2, 3, 1001, 70001):
        q, dq, ddq, tau = _states(oracle_fr3, B, seed=0x5EED0011)
        if layout == "aos":
            q, dq, ddq, tau = (np.ascontiguousarray(x.T) for x in (q, dq, ddq, tau))
        for name, args in (("rnea", (q, dq, ddq)), ("forward_dynamics", (q, dq, tau)), ("rnea_fd", (q, dq, ddq, tau)),
                           ("crba", (q,)), ("jac", (q,)), ("fwd_kin", (q,)), ("rnea_derivatives", (q, dq, ddq)),
                           ("fd_derivatives", (q, dq, tau))):
            a = getattr(mm, name)(*args, layout=layout)
            b = getattr(mb_fr3, name)(*args, layout=layout)
            if B == 1 and isinstance(a, tuple):
                for x, y in zip(a, b):
                    np.testing.assert_array_equal(x, y, err_msg=f"{name} B={B}")
            else:
                np.testing.assert_array_equal(a, b, err_msg=f"{name} B={B}")
    # against the oracle, once
    q, dq, ddq, tau = _states(oracle_fr3, 4097, seed=0x5EED0012)
    assert state_err(mm.rnea(q, dq, ddq), oracle_fr3.rnea_batch(q, dq, ddq), 0).max() < TOL
    # a padded leading dimension through the raw C ABI: rows of 5000 doubles, 4097 states
    from rigidbody_rs_b200 import _lib
    ld = 5000
    bufs = [np.full((7, ld), np.nan) for _ in range(4)]
    for b_, x in zip(bufs, (q, dq, ddq)):
        b_[:, :4097] = x
    rc = _lib.lib.multibody_rnea_batch(mm._h, *[C.c_void_p(b_.ctypes.data) for b_ in bufs], 4097, ld, 0, 0, None)
    assert rc == 0, _lib.lib.multibody_last_error()
    np.testing.assert_array_equal(bufs[3][:, :4097], mb_fr3.rnea(q, dq, ddq))
    assert np.isnan(bufs[3][:, 4097:]).all()
    mm.close()


def test_multi_device_rollout_and_status(rb, mb_fr3, oracle_fr3):
    """Rollouts (SoA and AoS: the AoS trajectory arrays are strided per slice) and the fused cost over 2 devices equal
    the single-device results bit for bit; device pointers are refused with RB_ERR_UNSUPPORTED; a non-SPD state in ONE
    device's slice is reported by the call; launch counts add up."""
    import torch
    devs = _two_devices()
    mm = rb.Multibody.from_urdf(FR3, devices=devs)
    B, H, dt = 301, 12, 1e-3
    q, dq, _, _ = _states(oracle_fr3, B, seed=0x5EED0003)
    lim = oracle_fr3.model
    tau = np.stack([oracle_fr3.fill(0x5EED0003, 4 + t % 32, -lim.effort, lim.effort, t * B, B) for t in range(H)])
    w = np.linspace(0.5, 2.0, 7)
    l0 = mm.launch_count
    for x, y in zip(mm.rollout(q, dq, tau, dt, final=True), mb_fr3.rollout(q, dq, tau, dt, final=True)):
        np.testing.assert_array_equal(x, y)
    assert mm.launch_count - l0 == 2
    qa, dqa, ta = np.ascontiguousarray(q.T), np.ascontiguousarray(dq.T), np.ascontiguousarray(tau.transpose(0, 2, 1))
    for x, y in zip(mm.rollout(qa, dqa, ta, dt, layout="aos", final=True), mb_fr3.rollout(qa, dqa, ta, dt, layout="aos", final=True)):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(mm.rollout_cost(q, dq, tau, dt, w_q=w, w_tau=1e-3 * w, w_q_final=3 * w),
                                  mb_fr3.rollout_cost(q, dq, tau, dt, w_q=w, w_tau=1e-3 * w, w_q_final=3 * w))
    with pytest.raises(rb.RigidBodyError) as ei:
        mm.rnea(*(torch.zeros((7, 8), dtype=torch.float64, device="cuda:0") for _ in range(3)))
    assert ei.value.code == -6
    mm.close()
    # a chain whose mass matrix is indefinite: the error raised on a worker thread reaches the caller, and the next call is clean
    n = 2
    R = np.tile(np.eye(3), (n, 1, 1)); t = np.zeros((n, 3)); t[1] = [0.1, 0, 0]
    d = rb._lib.RbChainDesc()
    keep = [np.ascontiguousarray(x) for x in (R, t, np.ones(n), np.zeros((n, 3)), np.tile(np.diag([0.1, 0.1, -5.0]), (n, 1, 1)))]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    d.n_joints = n
    d.parent_rot, d.parent_trans, d.mass, d.com, d.inertia_com = (dp(x) for x in keep)
    d.gravity[:] = [0.0, 0.0, 9.81]
    h = C.c_void_p()
    devs = (C.c_int * 2)(*devs)
    assert rb._lib.lib.multibody_gpu_new_multi(C.byref(d), devs, 2, C.byref(h)) == 0, rb._lib.lib.multibody_last_error()
    bad = rb.Multibody(h)
    z = np.zeros((2, 64))
    with pytest.raises(rb.NotPositiveDefinite):
        bad.forward_dynamics(z, z, z)
    assert bad.rnea(z, z, z).shape == (2, 64)
    bad.close()


def test_multi_device_argument_checks(rb):
    from rigidbody_rs_b200 import _lib
    h = C.c_void_p()
    one = (C.c_int * 1)(0)
    assert _lib.lib.multibody_gpu_new_multi_from_urdf(FR3.encode(), one, 1, C.byref(h)) == 0
    g = rb.Multibody(h)
    assert g.n_devices == 1 and g.peer(0).device == 0            # n_dev = 1 is an ordinary engine
    g.close()
    twice = (C.c_int * 2)(0, 0)
    os.environ.pop("RIGIDBODY_B200_ALLOW_DUPLICATE_DEVICES", None)
    assert _lib.lib.multibody_gpu_new_multi_from_urdf(FR3.encode(), twice, 2, C.byref(h)) == _lib.RB_ERR_ARG
    assert _lib.lib.multibody_gpu_new_multi_from_urdf(FR3.encode(), None, 2, C.byref(h)) == _lib.RB_ERR_NULL
    far = (C.c_int * 2)(0, 63)
    assert _lib.lib.multibody_gpu_new_multi_from_urdf(FR3.encode(), far, 2, C.byref(h)) == _lib.RB_ERR_ARG
    up, down = rb.Multibody.from_urdf(FR3).copy_peak(64 << 20, 32 << 20)
    assert up > 1.0 and down > 0.5
