#!/usr/bin/env python
"""End-to-end probe of the single-process multi-device engine: one RB_MEM_HOST multibody_rnea_fd_batch call over all
visible GPUs, pinned host buffers, against the bare pinned-copy ceiling (multibody_gpu_measure_copy_peak).
usage: python tools/e2e_probe.py [--states N] [--devices 0,1,...] [--reps K]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rigidbody_rs_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--states", type=int, default=1 << 24)
ap.add_argument("--devices", default="")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--pageable", action="store_true")
a = ap.parse_args()
devs = [int(x) for x in a.devices.split(",")] if a.devices else list(range(torch.cuda.device_count()))
mb = rb.Multibody.from_urdf("assets/fr3.urdf", devices=devs)
n, B = mb.n, a.states
alloc = (lambda shape: np.empty(shape)) if a.pageable else rb.host_empty
hq, hdq, hddq, htau = (alloc((n, B)) for _ in range(4))
hout = alloc((2 * n, B))
rng = np.random.default_rng(0)
for h in (hq, hdq, hddq, htau):
    h[...] = rng.uniform(-1, 1, (n, 1)) + np.linspace(0, 1, B)[None, :]
mb.rnea_fd(hq, hdq, hddq, htau, out=hout)
ts = []
for _ in range(a.reps):
    t0 = time.perf_counter(); mb.rnea_fd(hq, hdq, hddq, htau, out=hout); ts.append(time.perf_counter() - t0)
sec = min(ts)
up_b, down_b = 4 * n * B * 8, 2 * n * B * 8
cu, cd = mb.copy_peak(up_b, down_b, 2)
print(json.dumps({"devices": devs, "states": B, "pageable": a.pageable, "ms": sec * 1e3, "evals_per_s": 2 * B / sec,
                  "h2d_GBs": up_b / sec / 1e9, "d2h_GBs": down_b / sec / 1e9,
                  "copy_ceiling": {"h2d_GBs": cu, "d2h_GBs": cd, "ms": up_b / cu / 1e6}, "frac_of_ceiling": (up_b / cu / 1e9) / sec}))
