import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FR3 = os.path.join(ROOT, "assets", "fr3.urdf")
CHAIN32 = os.path.join(ROOT, "assets", "chain32.urdf")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# fp64 tolerance of BASELINE.json's north_star, in the per-state form of SURVEY.md section 8c:
#   max_i |x_i - ref_i| <= 1e-10 * max(1, ||ref||_inf)
TOL = 1e-10


# Forward dynamics solves H qdd = b: no fp64 algorithm can promise more than the forward-error bound of an SPD solve,
#   |x - x_true|_inf <= c * cond_2(H) * eps * |x|_inf,
# so where cond(H) * eps approaches 1e-10 (the 32-joint chain: cond up to ~5e4; random long chains: up to ~1e6) the bar
# is the larger of 1e-10 and that bound.  FD_C is the constant measured on the GPU against a 40-digit mpmath solve
# (tools/error_budget.py -> profiles/r2_error_budget.jsonl: largest c seen 2.6 for the CUDA paths, 1.9 for the oracle's
# LL^T), with head-room.  For the FR3 (cond <= ~1e3) the bound is below 1e-10 and TOL applies unchanged.
EPS = 2.0 ** -52
FD_C = 8.0


def fd_bound(cond):
    """Per-state bar for a forward-dynamics result in the state_err() normalisation."""
    return np.maximum(TOL, FD_C * np.asarray(cond, dtype=np.float64) * EPS)


def spd_cond(H, n=None):
    """2-norm condition numbers of a batch of mass matrices.  Accepts [B,n,n] (full or upper-triangular, the reference's
    convention) or the C ABI's SoA block [n*n, B] (entry r + n*c) together with n."""
    H = np.asarray(H, dtype=np.float64)
    if H.ndim == 2:
        H = H.reshape(n, n, -1).transpose(2, 1, 0)            # [B, r, c]
    lower_empty = np.abs(np.tril(H, -1)).max() == 0.0
    if lower_empty:
        H = H + np.triu(H, 1).transpose(0, 2, 1)
    return np.linalg.cond(H)


def fd_close(got, ref, cond, axis, sides=1):
    """All states within fd_bound(cond); `sides` = 2 when `ref` is itself an fp64 solve (oracle, twin)."""
    err = state_err(got, ref, axis)
    return bool((err <= sides * fd_bound(cond)).all()), float((err / fd_bound(cond)).max())


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def state_err(x, ref, axis):
    """Per-state error in the SURVEY 8c form; `axis` is the per-state (joint/entry) axis."""
    x, ref = np.asarray(x), np.asarray(ref)
    return np.abs(x - ref).max(axis) / np.maximum(1.0, np.abs(ref).max(axis))


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built():
    """Builds the product library and the oracle once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_fr3(built):
    from oracle.rb_oracle import Oracle
    return Oracle.from_urdf(FR3)


@pytest.fixture(scope="session")
def oracle_chain32(built):
    from oracle.rb_oracle import Oracle
    return Oracle.from_urdf(CHAIN32)


def _need_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device (there is no CPU fallback to test)")


@pytest.fixture(scope="session")
def rb(built):
    import rigidbody_rs_b200
    return rigidbody_rs_b200


@pytest.fixture(scope="session")
def mb_fr3(rb):
    _need_cuda()
    return rb.Multibody.from_urdf(FR3, device=0)


@pytest.fixture(scope="session")
def mb_chain32(rb):
    _need_cuda()
    return rb.Multibody.from_urdf(CHAIN32, device=0)
