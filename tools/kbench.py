#!/usr/bin/env python
"""Kernel-only timing of the FR3 hot path for tuning experiments (CUDA events, device-resident inputs).
usage: [RIGIDBODY_B200_LIB=path.so] python tools/kbench.py [--states N] [--iters K] [--tag name]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rigidbody_rs_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--states", type=int, default=1 << 24)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--tag", default="")
ap.add_argument("--urdf", default="assets/fr3.urdf")
ap.add_argument("--ops", default="rnea,fd")
ap.add_argument("--layout", default="soa", choices=["soa", "aos"])
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
ap.add_argument("--traj", type=int, default=65536, help="trajectories of the rollout op")
a = ap.parse_args()
mb = rb.Multibody.from_urdf(a.urdf)
n, B = mb.n, a.states
lim = mb.limits()
dev = torch.device("cuda:0")
q = torch.empty((n, B), dtype=torch.float64, device=dev)
dq, x3, out = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
mb.fill(q, 1, 0, lim["lower"], lim["upper"]); mb.fill(dq, 1, 1, -lim["velocity"], lim["velocity"]); mb.fill(x3, 1, 2, -10.0, 10.0)
res = {"tag": a.tag, "variant": mb.kernel_variant, "states": B, "layout": a.layout, "dtype": a.dtype}
H = 64
for op in a.ops.split(","):
    units = B
    if op == "rollout":
        Bt = min(B, a.traj)
        tau = torch.empty((H, n, Bt), dtype=torch.float64, device=dev)
        for t in range(H):
            mb.fill(tau[t], 3, 4 + t % 32, -lim["effort"], lim["effort"], first_index=t * Bt)
        q0, dq0 = q[:, :Bt].contiguous(), dq[:, :Bt].contiguous()
        fn = lambda: mb.rollout(q0, dq0, tau, 1e-3)
        units = Bt * H
    elif op == "rnea_fd":
        tin = torch.empty_like(q); mb.fill(tin, 2, 3, -lim["effort"], lim["effort"])
        out2 = torch.empty((2 * n, B), dtype=torch.float64, device=dev)
        fn = lambda: mb.rnea_fd(q, dq, x3, tin, out=out2)
        units = 2 * B
    elif op == "crba":
        Bc = min(B, 1 << 22)
        qc = q[:, :Bc].contiguous(); Hout = torch.empty((n * n, Bc), dtype=torch.float64, device=dev)
        fn = lambda: mb.crba(qc, out=Hout)
        units = Bc
    elif op in ("fk", "jac"):
        Bc = min(B, 1 << 23)
        qc = q[:, :Bc].contiguous()
        O = torch.empty((3 if op == "fk" else 6 * n, Bc), dtype=torch.float64, device=dev)
        fn = (lambda: mb.fwd_kin(qc, out=O)) if op == "fk" else (lambda: mb.jac(qc, out=O))
        units = Bc
    elif op in ("rnea_deriv", "fd_deriv"):
        Bc = min(B, 1 << 21)
        blocks = 2 if op == "rnea_deriv" else 3
        qc, dqc, xc = (t[:, :Bc].contiguous() for t in (q, dq, x3))
        Dout = torch.empty((blocks * n * n, Bc), dtype=torch.float64, device=dev)
        fn = (lambda: mb.rnea_derivatives(qc, dqc, xc, out=Dout)) if op == "rnea_deriv" else (lambda: mb.fd_derivatives(qc, dqc, xc, out=Dout))
        units = Bc
    else:
        if a.dtype == "f32":
            q32, dq32, x32, o32 = (t.float() for t in (q, dq, x3, out))
            fn = {"rnea": lambda: mb.rnea(q32, dq32, x32, out=o32), "fd": lambda: mb.forward_dynamics(q32, dq32, x32, out=o32)}[op]
        elif a.layout == "aos":
            qa, dqa, x3a, outa = (t.t().contiguous() for t in (q, dq, x3, out))
            fn = {"rnea": lambda: mb.rnea(qa, dqa, x3a, layout="aos", out=outa),
                  "fd": lambda: mb.forward_dynamics(qa, dqa, x3a, layout="aos", out=outa)}[op]
        else:
            fn = {"rnea": lambda: mb.rnea(q, dq, x3, out=out), "fd": lambda: mb.forward_dynamics(q, dq, x3, out=out)}[op]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    res[op] = {"ms_med": round(med, 4), "ms_min": round(ts[0], 4), "Geval_s": round(units / med / 1e6, 3)}
print(json.dumps(res), flush=True)
