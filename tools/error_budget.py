#!/usr/bin/env python
"""Who owns the forward-dynamics error?  For each chain: |gpu - truth| and |oracle - truth| against a 40-digit mpmath
evaluation (oracle/rb_oracle_np.py::fd_mp) on a few states, and |gpu - oracle| on many, all expressed as
c = err / (cond(H) * eps * max(1, |x|_inf)) -- the constant of the textbook forward-error bound of an SPD solve.
tests/conftest.py::FD_C is set from the largest c seen here.  Test infrastructure (uses oracle/).
usage: python tools/error_budget.py > profiles/r2_error_budget.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rigidbody_rs_b200 as rb
from oracle.rb_oracle import Oracle
from oracle.rb_oracle_np import ChainNP, fd_mp
from test_host import _random_chain

EPS = 2.0 ** -52


def conds(H):
    return np.linalg.cond(H)


def report(name, got, ref, cond, extra=None):
    scale = np.maximum(1.0, np.abs(ref).max(1))
    err = np.abs(got - ref).max(1) / scale
    c = err / (cond * EPS)
    d = {"case": name, "states": int(len(err)), "max_err": float(err.max()), "max_cond": float(cond.max()), "max_c": float(c.max()),
         "median_c": float(np.median(c))}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def urdf_case(name, urdf, variants, B, n_truth, qlim, taulim):
    o = Oracle.from_urdf(os.path.join(ROOT, urdf)); n = o.model.n
    rng = np.random.default_rng(11)
    q, dq, tau = rng.uniform(-qlim, qlim, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-taulim, taulim, (B, n))
    Hu = o.crba_batch(q, layout="aos").reshape(B, n, n).transpose(0, 2, 1)
    cond = conds(Hu + np.triu(Hu, 1).transpose(0, 2, 1))
    ref = o.forward_dynamics_batch(q, dq, tau, layout="aos")
    truth = np.array([fd_mp(o.model, q[k], dq[k], tau[k]) for k in range(n_truth)])
    report(f"{name}: oracle vs mpmath", ref[:n_truth], truth, cond[:n_truth])
    for v in variants:
        os.environ["RIGIDBODY_B200_VARIANT"] = v
        try:
            mb = rb.Multibody.from_urdf(os.path.join(ROOT, urdf))
        finally:
            os.environ.pop("RIGIDBODY_B200_VARIANT", None)
        got = mb.forward_dynamics(q, dq, tau, layout="aos")
        report(f"{name} [{mb.kernel_variant}]: gpu vs mpmath", got[:n_truth], truth, cond[:n_truth])
        report(f"{name} [{mb.kernel_variant}]: gpu vs oracle", got, ref, cond)
        ddq = rng.uniform(-10, 10, (B, n))
        t = mb.rnea(q, dq, ddq, layout="aos")
        report(f"{name} [{mb.kernel_variant}]: round trip FD(rnea(ddq)) vs ddq", mb.forward_dynamics(q, dq, t, layout="aos"), ddq, cond)


def random_case(n, seed, B=333, scale_t=True):
    R, t, m, c, Ic = _random_chain(n, seed)
    if scale_t:
        t = t * (8.0 / n)
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic)
    ch = ChainNP.from_arrays(R, t, m, c, Ic)
    rng = np.random.default_rng(seed)
    q, dq, ddq = rng.uniform(-3, 3, (B, n)), rng.uniform(-2, 2, (B, n)), rng.uniform(-10, 10, (B, n))
    cond = conds(ch.crba(q, symmetric=True))
    tau = ch.rnea(q, dq, ddq)
    report(f"random chain n={n} [{mb.kernel_variant}]: round trip", mb.forward_dynamics(q, dq, tau, layout="aos"), ddq, cond)
    tau2 = rng.uniform(-20, 20, (B, n))
    report(f"random chain n={n} [{mb.kernel_variant}]: gpu vs numpy twin", mb.forward_dynamics(q, dq, tau2, layout="aos"),
           ch.forward_dynamics(q, dq, tau2), cond)


urdf_case("fr3", "assets/fr3.urdf", ("auto", "jit-specialised", "generic-7", "generic-n"), 4096, 24, np.pi, 80.0)
urdf_case("chain32", "assets/chain32.urdf", ("auto", "generic-n"), 4096, 8, np.pi, 50.0)
for n, seed in ((3, 21), (6, 22), (10, 23), (13, 1), (15, 7), (19, 6), (24, 2), (32, 3), (33, 4), (64, 5)):
    random_case(n, seed)
