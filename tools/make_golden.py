#!/usr/bin/env python
"""Writes tests/golden/*.json: known-answer vectors for the hot path.

Two kinds of anchors:
  * `survey_kat`  -- the SURVEY.md section 8c values, produced by an independent numpy restatement of the cited
                     reference lines (NOT by this repo's oracle, NOT by the Rust crate: it cannot be built here).
  * `oracle`      -- outputs of oracle/rb_oracle.c (the reference-shaped C restatement) on seeded states, frozen
                     so that later edits to the oracle or the kernels cannot drift silently.
PARITY UNPINNED by the reference itself: its tests hold no rnea/crba values.
Run from the repo root:  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.rb_oracle import Oracle  # noqa: E402

SURVEY_KAT = {
    "source": "SURVEY.md section 8c (numpy restatement made during the survey; ~1e-12 smoke anchors)",
    "rnea_zero": {"q": [0.0] * 7, "dq": [0.0] * 7, "ddq": [0.0] * 7,
                  "tau": [0.0, -3.434431907689, 0.0, -3.257223811962, 0.0, 1.694216798552, 0.0]},
    "main_cpp": {"q": [0, 0, 1, 0, 1, 0, 0], "dq": [0, 0, 0, 0, 1, 0, 0], "ddq": [1, 0, 0, 0, 0, 1, 0],
                 "tau": [0.114975594093, 0.440352492965, 0.077638145604, -4.496336550466, 0.034046184773,
                         1.71942616226, -0.006413948423],
                 "crba_diag": [0.1143336658895, 2.603206629183, 0.07699621740009, 0.6017679340039,
                               0.03340425656912, 0.03175095254849, 0.004909651967361],
                 "H01": -0.1441771690957, "H13": -0.5932369920521, "H56": -5.483591064082e-4},
    "generic": {"q": [0.1, -0.2, 0.3, -1.5, 0.5, 1.6, -0.7], "dq": [0.5, -0.4, 0.3, -0.2, 0.1, 0.6, -0.7],
                "ddq": [1, -2, 3, -4, 5, -6, 7],
                "tau": [4.08436515612732, -17.9732901023693, 3.23976939321532, 15.5232578036260,
                        0.973033907775699, 1.01502347286127, -1.00904116038e-4]},
    "fwd_kin_zero": [0.088, 0.0, 1.033],
}


def frozen(urdf, count, seed):
    o = Oracle.from_urdf(os.path.join(ROOT, urdf))
    n = o.n
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, (count, n)); dq = rng.uniform(-2, 2, (count, n))
    ddq = rng.uniform(-10, 10, (count, n)); tau = rng.uniform(-50, 50, (count, n))
    q[0] = 0; dq[0] = 0; ddq[0] = 0            # the reference's bench_rnea input (multibody.rs:206-208)
    L = lambda a: [[float(x) for x in row] for row in np.atleast_2d(a)]
    return {
        "urdf": urdf, "n": n, "q": L(q), "dq": L(dq), "ddq": L(ddq), "tau_in": L(tau),
        "rnea": L(o.rnea_batch(q, dq, ddq, layout="aos")),
        "fd": L(o.forward_dynamics_batch(q, dq, tau, layout="aos")),
        "crba_colmajor": L(o.crba_batch(q[:8], layout="aos")),
        "fwd_kin": L(np.stack([o.fwd_kin(x) for x in q[:8]])),
        "jac_colmajor": L(np.stack([o.jac(x).T.reshape(-1) for x in q[:8]])),
    }


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "survey_kat.json"), "w") as f:
        json.dump(SURVEY_KAT, f, indent=1)
    with open(os.path.join(out, "fr3_oracle.json"), "w") as f:
        json.dump(frozen("assets/fr3.urdf", 32, 20261018), f)
    with open(os.path.join(out, "chain32_oracle.json"), "w") as f:
        json.dump(frozen("assets/chain32.urdf", 4, 20261019), f)
    print("wrote tests/golden/{survey_kat,fr3_oracle,chain32_oracle}.json")
