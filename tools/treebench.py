#!/usr/bin/env python
"""Times RNEA / forward dynamics of a quadruped-like kinematic tree (4 legs x 3 joints hanging off the base) through
the run-time specialised kernels and through the run-time-n family.  usage: python tools/treebench.py [--states N]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rigidbody_rs_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--states", type=int, default=1 << 22)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--serial", type=int, default=0, help="time a serial chain of this many joints (same links) instead of the tree")
a = ap.parse_args()
n, B = (a.serial or 12), a.states
parent = np.array([-1, 0, 1, -1, 3, 4, -1, 6, 7, -1, 9, 10], dtype=np.int32) if not a.serial else np.arange(n, dtype=np.int32) - 1
rng = np.random.default_rng(0)
from scipy.spatial.transform import Rotation
R = np.stack([Rotation.from_euler("xyz", [0.0, 0.0, 0.0] if i % 3 == 0 else [np.pi / 2 if i % 3 == 1 else 0.0, 0.0, 0.0]).as_matrix() for i in range(n)])
t = np.array([[0.2 * (1 if (i // 3) % 2 == 0 else -1), 0.1 * (1 if i // 6 == 0 else -1), 0.0] if i % 3 == 0 else [0.0, 0.0, -0.2] for i in range(n)])
m = np.array(([0.7, 1.0, 0.2] * 22)[:n])
c = np.tile(np.array([0.0, 0.01, -0.08]), (n, 1))
Ic = np.stack([np.diag([0.004, 0.004, 0.001]) * mi for mi in m])
for variant in ("auto", "generic-n"):
    os.environ["RIGIDBODY_B200_VARIANT"] = variant
    mb = rb.Multibody.from_descriptor(R, t, m, c, Ic, parent=parent)
    os.environ.pop("RIGIDBODY_B200_VARIANT")
    dev = torch.device("cuda:0")
    q = torch.empty((n, B), dtype=torch.float64, device=dev); dq, x3, out = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    mb.fill(q, 1, 0, -2.0, 2.0); mb.fill(dq, 1, 1, -2.0, 2.0); mb.fill(x3, 1, 2, -10.0, 10.0)
    res = {"chain": "quadruped-like tree, 12 joints" if not a.serial else f"serial chain, {n} joints", "variant": mb.kernel_variant, "states": B}
    for op, fn in (("rnea", lambda: mb.rnea(q, dq, x3, out=out)), ("fd", lambda: mb.forward_dynamics(q, dq, x3, out=out))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[op] = {"ms_med": round(ts[len(ts) // 2], 4), "Geval_s": round(B / ts[len(ts) // 2] / 1e6, 3)}
    print(json.dumps(res), flush=True)
