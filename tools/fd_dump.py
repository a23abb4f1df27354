#!/usr/bin/env python
"""Writes forward dynamics of N seeded states of a model to a .npy (variant-vs-variant comparisons of experiment builds).
usage: [RIGIDBODY_B200_LIB=...] python tools/fd_dump.py out.npy [urdf] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rigidbody_rs_b200 as rb
urdf = sys.argv[2] if len(sys.argv) > 2 else "assets/chain32.urdf"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4099
mb = rb.Multibody.from_urdf(urdf)
rng = np.random.default_rng(7)
q, dq, tau = rng.uniform(-3, 3, (mb.n, N)), rng.uniform(-2, 2, (mb.n, N)), rng.uniform(-50, 50, (mb.n, N))
np.save(sys.argv[1], mb.forward_dynamics(q, dq, tau))
