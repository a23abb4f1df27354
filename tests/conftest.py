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
