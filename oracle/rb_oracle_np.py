"""Second, independent CPU restatement (numpy, matrix form, vectorised over the batch).

ORACLE = test infrastructure, NOT product code.  PARITY UNPINNED (see oracle/rb_oracle.h).

Where oracle/rb_oracle.c follows the reference's *shape* (quaternion isometries, (m, c, I_c)
inertias, one state at a time), this file restates the same algorithm in the generalised form
SURVEY.md section 7 step 1 asks for: 3x3 rotation matrices, 10-parameter spatial inertias
(m, h = m c, I_o), any serial chain length.  tests/test_oracle.py proves the two equal to <= 1e-12,
which is evidence (1) of SURVEY.md section 8c.  An mpmath evaluation of the same formulas
(`rnea_mp`) bounds the absolute rounding error of both.

Conventions (all citations relative to the reference checkout):
  X_i(q) = parent_i o Rot(z, q_i) is the pose of frame i in frame i-1 (joint.rs:36-38), R_i = R_p Rz(q).
  motion  parent->child: rot' = R^T rot, lin' = R^T (lin - t x rot)          (spatial.rs:110-116)
  force   child->parent: lin' = R lin,  rot' = R rot + t x (R lin)           (spatial.rs:242-248 on X^-1)
  I a: lin = m a.lin - h x a.rot ; rot = I_o a.rot + h x a.lin               (inertia.rs:107-117)
  v x* f: lin = w x f.lin ; rot = w x f.rot + v.lin x f.lin                  (spatial.rs:129-134)
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.81          # multibody.rs:118, base acceleration +z


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


class ChainNP:
    """Flattened serial chain: R_p [n,3,3], t_p [n,3], m [n], h [n,3], I_o [n,3,3]."""

    def __init__(self, model):
        n = model.n
        self.n = n
        self.Rp = np.stack([_rz(y) @ _ry(p) @ _rx(r) for r, p, y in model.rpy])   # joint.rs:59-63
        self.tp = np.array(model.xyz, dtype=np.float64)
        self.m = np.array(model.mass, dtype=np.float64)
        c = np.array(model.com, dtype=np.float64)
        self.h = self.m[:, None] * c
        Io = []
        for i in range(n):
            s = model.inertia6[i]
            Ic = np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])
            C = _skew(c[i])
            Io.append(Ic + self.m[i] * C @ C.T)                                   # inertia.rs:31-32
        self.Io = np.stack(Io)
        self.axis = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))      # the dynamics' axis (multibody.rs:29,130)
        self.parent = np.arange(n) - 1                               # serial chain (multibody.rs:148,165)

    @classmethod
    def from_arrays(cls, Rp, tp, mass, com, inertia_com, axis=None, parent=None):
        """Chain given directly by its flattened descriptor (what RbChainDesc carries).  `axis` [n,3]: joint axes in
        the joint frames, used CONSISTENTLY by kinematics and dynamics (the reference uses them in the kinematics
        only, joint.rs:48-50 vs multibody.rs:130; for +z the two agree).  `parent` [n]: parent link of each joint,
        -1 = base, parent[i] < i (kinematic tree; the reference is the serial case parent[i] = i-1).  The tip of
        fwd_kin / jac is the last link."""
        self = cls.__new__(cls)
        self.n = len(mass)
        self.parent = np.arange(self.n) - 1 if parent is None else np.array(parent, dtype=int)
        ax = np.tile(np.array([0.0, 0.0, 1.0]), (self.n, 1)) if axis is None else np.array(axis, dtype=np.float64)
        self.axis = ax / np.linalg.norm(ax, axis=1, keepdims=True)
        self.Rp = np.array(Rp, dtype=np.float64); self.tp = np.array(tp, dtype=np.float64)
        self.m = np.array(mass, dtype=np.float64)
        c = np.array(com, dtype=np.float64)
        self.h = self.m[:, None] * c
        self.Io = np.stack([np.array(inertia_com[i]) + self.m[i] * _skew(c[i]) @ _skew(c[i]).T for i in range(self.n)])
        return self

    # -- helpers, all batched over leading axis B
    def _R(self, i, q):
        """R_i(q) = R_p Rot(axis_i, q): [B,3,3]  (Rodrigues; Rz(q) for the +z axis)"""
        c, s = np.cos(q)[..., None, None], np.sin(q)[..., None, None]
        a = self.axis[i]
        K = _skew(a)
        Rq = np.eye(3) + s * K + (1.0 - c) * (K @ K)
        return self.Rp[i] @ Rq

    def _Imul(self, i, lin, rot):
        f_lin = self.m[i] * lin - np.cross(self.h[i], rot)
        f_rot = rot @ self.Io[i].T + np.cross(self.h[i], lin)
        return f_lin, f_rot

    def rnea(self, q, dq, ddq):
        """q, dq, ddq: [B, n] -> tau [B, n]  (multibody.rs:111-153)"""
        q, dq, ddq = (np.atleast_2d(np.asarray(x)) for x in (q, dq, ddq))
        dt = np.result_type(q.dtype, dq.dtype, ddq.dtype, np.float64)       # complex inputs: complex-step derivatives
        q, dq, ddq = q.astype(dt), dq.astype(dt), ddq.astype(dt)
        B, n = q.shape
        R = [self._R(i, q[:, i]) for i in range(n)]
        base = (np.zeros((B, 3)), np.zeros((B, 3)), np.tile(np.array([0.0, 0.0, GRAVITY]), (B, 1)), np.zeros((B, 3)))
        state = []
        f_lin, f_rot = [], []
        for i in range(n):
            v_lin, v_rot, a_lin, a_rot = state[self.parent[i]] if self.parent[i] >= 0 else base
            Rt = np.swapaxes(R[i], 1, 2)
            t = self.tp[i]
            v_lin = np.einsum("bij,bj->bi", Rt, v_lin - np.cross(t, v_rot))
            v_rot = np.einsum("bij,bj->bi", Rt, v_rot)
            ax = self.axis[i]
            v_rot = v_rot + ax * dq[:, i:i + 1]
            a_lin = np.einsum("bij,bj->bi", Rt, a_lin - np.cross(t, a_rot))
            a_rot = np.einsum("bij,bj->bi", Rt, a_rot)
            a_rot = a_rot + ax * ddq[:, i:i + 1]
            # v x (S dq), S = (axis; 0):  lin += v_lin x axis dq,  rot += v_rot x axis dq
            a_lin = a_lin + np.cross(v_lin, ax) * dq[:, i:i + 1]
            a_rot = a_rot + np.cross(v_rot, ax) * dq[:, i:i + 1]
            state.append((v_lin, v_rot, a_lin, a_rot))
            Ia_l, Ia_r = self._Imul(i, a_lin, a_rot)
            Iv_l, Iv_r = self._Imul(i, v_lin, v_rot)
            f_lin.append(Ia_l + np.cross(v_rot, Iv_l))
            f_rot.append(Ia_r + np.cross(v_rot, Iv_r) + np.cross(v_lin, Iv_l))
        tau = np.zeros((B, n), dtype=dt)
        for i in range(n - 1, -1, -1):
            tau[:, i] = f_rot[i] @ self.axis[i]
            p = self.parent[i]
            if p >= 0:
                Rl = np.einsum("bij,bj->bi", R[i], f_lin[i])
                f_lin[p] = f_lin[p] + Rl
                f_rot[p] = f_rot[p] + np.einsum("bij,bj->bi", R[i], f_rot[i]) + np.cross(self.tp[i], Rl)
        return tau

    def crba(self, q, symmetric=False):
        """q [B,n] -> H [B,n,n]; reference convention: diagonal + strict upper, lower = 0 (multibody.rs:155-174)."""
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        B, n = q.shape
        R = [self._R(i, q[:, i]) for i in range(n)]
        H = np.zeros((B, n, n))
        m = [np.full(B, self.m[i]) for i in range(n)]
        h = [np.tile(self.h[i], (B, 1)) for i in range(n)]
        Io = [np.tile(self.Io[i], (B, 1, 1)) for i in range(n)]
        for i in range(n - 1, -1, -1):
            ax = self.axis[i]
            F_lin = -np.cross(h[i], ax)                         # I * S,  S = (axis; 0)
            F_rot = Io[i] @ ax
            H[:, i, i] = F_rot @ ax
            j = i
            while self.parent[j] >= 0:                          # up the supporting branch only (zeros elsewhere)
                Rl = np.einsum("bij,bj->bi", R[j], F_lin)
                F_rot = np.einsum("bij,bj->bi", R[j], F_rot) + np.cross(self.tp[j], Rl)
                F_lin = Rl
                j = self.parent[j]
                H[:, j, i] = F_rot @ self.axis[j]
            p = self.parent[i]
            if p >= 0:
                # composite inertia into the parent frame, 10-parameter form:
                # h' = R h + m t ;  I_o' = R I_o R^T - [t]x[Rh]x - [Rh]x[t]x - m [t]x[t]x
                t = self.tp[i]
                Rh = np.einsum("bij,bj->bi", R[i], h[i])
                RIR = R[i] @ Io[i] @ np.swapaxes(R[i], 1, 2)
                T = _skew(t)
                S = np.zeros((B, 3, 3))
                S[:, 0, 1] = -Rh[:, 2]; S[:, 0, 2] = Rh[:, 1]; S[:, 1, 0] = Rh[:, 2]
                S[:, 1, 2] = -Rh[:, 0]; S[:, 2, 0] = -Rh[:, 1]; S[:, 2, 1] = Rh[:, 0]
                Io[p] = Io[p] + RIR - T @ S - S @ T - m[i][:, None, None] * (T @ T)
                h[p] = h[p] + Rh + m[i][:, None] * t
                m[p] = m[p] + m[i]
        if symmetric:
            iu = np.triu_indices(n, 1)
            H[:, iu[1], iu[0]] = H[:, iu[0], iu[1]]
        else:
            il = np.tril_indices(n, -1)
            H[:, il[0], il[1]] = 0.0
        return H

    def forward_dynamics(self, q, dq, tau):
        """qdd = chol_solve(sym(crba(q)), tau - rnea(q,dq,0))."""
        q, dq, tau = (np.atleast_2d(np.asarray(x, dtype=np.float64)) for x in (q, dq, tau))
        c = self.rnea(q, dq, np.zeros_like(q))
        H = self.crba(q, symmetric=True)
        L = np.linalg.cholesky(H)
        y = np.linalg.solve(L, (tau - c)[..., None])
        return np.linalg.solve(np.swapaxes(L, 1, 2), y)[..., 0]

    def rnea_derivatives(self, q, dq, ddq):
        """(d tau / d q, d tau / d dq), each [B, n, n] with [b, r, c] = d tau_r / d x_c, by complex-step
        differentiation of `rnea` itself (h = 1e-30: exact to rounding, no subtraction) -- independent of the
        world-frame closed form the CUDA kernels use."""
        q, dq, ddq = (np.atleast_2d(np.asarray(x, dtype=np.float64)) for x in (q, dq, ddq))
        B, n = q.shape
        h = 1e-30
        Dq, Dv = np.zeros((B, n, n)), np.zeros((B, n, n))
        for c in range(n):
            e = np.zeros(n, dtype=complex); e[c] = 1j * h
            Dq[:, :, c] = self.rnea(q + e, dq, ddq).imag / h
            Dv[:, :, c] = self.rnea(q, dq + e, ddq).imag / h
        return Dq, Dv

    def fd_derivatives(self, q, dq, tau):
        """(d qdd / d q, d qdd / d dq, H^-1 = d qdd / d tau) of qdd = forward_dynamics(q, dq, tau):
        d qdd / d x = -H^-1 (d rnea / d x at ddq = qdd)."""
        qdd = self.forward_dynamics(q, dq, tau)
        Dq, Dv = self.rnea_derivatives(q, dq, qdd)
        Minv = np.linalg.inv(self.crba(q, symmetric=True))
        return -Minv @ Dq, -Minv @ Dv, Minv

    def fwd_kin(self, q):
        """q [B,n] -> (R [B,3,3], p [B,3]) of the tip in the base frame (multibody.rs:87-93)."""
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        B, n = q.shape
        Racc = np.tile(np.eye(3), (B, 1, 1)); p = np.zeros((B, 3))
        i = n - 1
        while i >= 0:                                           # tip = last link; walk its supporting branch
            Ri = self._R(i, q[:, i])
            p = np.einsum("bij,bj->bi", Ri, p) + self.tp[i]
            Racc = Ri @ Racc
            i = self.parent[i]
        return Racc, p

    def jac(self, q):
        """q [B,n] -> J [B,6,n], rows lin then rot, expressed in the tip frame (multibody.rs:95-108)."""
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        B, n = q.shape
        J = np.zeros((B, 6, n))
        Racc = np.tile(np.eye(3), (B, 1, 1)); p = np.zeros((B, 3))   # pose of tip in frame i
        i = n - 1
        while i >= 0:                                           # joints off the tip's branch keep zero columns
            Rt = np.swapaxes(Racc, 1, 2)
            z = self.axis[i]
            J[:, 0:3, i] = np.einsum("bij,bj->bi", Rt, -np.cross(p, z))
            J[:, 3:6, i] = Rt @ z
            Ri = self._R(i, q[:, i])
            p = np.einsum("bij,bj->bi", Ri, p) + self.tp[i]
            Racc = Ri @ Racc
            i = self.parent[i]
        return J

    def potential_energy(self, q):
        """U(q) = sum_i m_i g z_i(com) with the base accelerating +z (gravity points -z)."""
        q = np.atleast_2d(np.asarray(q, dtype=np.float64))
        B, n = q.shape
        U = np.zeros(B)
        pose = []
        for i in range(n):
            Racc, p = pose[self.parent[i]] if self.parent[i] >= 0 else (np.tile(np.eye(3), (B, 1, 1)), np.zeros((B, 3)))
            Ri = self._R(i, q[:, i])
            p = p + np.einsum("bij,j->bi", Racc, self.tp[i])
            Racc = Racc @ Ri
            pose.append((Racc, p))
            com_w = p + np.einsum("bij,j->bi", Racc, self.h[i] / self.m[i])
            U += self.m[i] * GRAVITY * com_w[:, 2]
        return U


def rnea_mp(model, q, dq, ddq, dps=40):
    """mpmath evaluation of the matrix-form RNEA for ONE state at `dps` digits (error bound for the fp64 paths).
    URDF decimal strings are taken at their fp64 values, as the reference does."""
    return np.array([float(x) for x in _rnea_mp(model, q, dq, ddq, dps)])


def fd_mp(model, q, dq, tau, dps=40, return_cond=False):
    """mpmath evaluation of forward dynamics for ONE state: the composition SURVEY.md 3.3 defines,
    qdd = H(q)^-1 (tau - rnea(q, dq, 0)), with H built column by column from the same `dps`-digit recursion
    (H e_j = rnea(q, 0, e_j) - rnea(q, 0, 0), multibody.rs:111-153) and the system solved at `dps` digits: the value every
    fp64 path (oracle LL^T, CUDA LDL^T, quarter-warp elimination) is an approximation OF.  With return_cond also the
    2-norm condition number of H (fp64 is plenty for that)."""
    import mpmath as mp
    mp.mp.dps = dps
    n = model.n
    zero = np.zeros(n)
    g = _rnea_mp(model, q, zero, zero, dps)
    H = mp.matrix(n, n)
    for j in range(n):
        e = np.zeros(n); e[j] = 1.0
        col = _rnea_mp(model, q, zero, e, dps)
        for i in range(n):
            H[i, j] = col[i] - g[i]
    bias = _rnea_mp(model, q, dq, zero, dps)
    b = mp.matrix([mp.mpf(float(tau[i])) - bias[i] for i in range(n)])
    x = mp.lu_solve(H, b)
    out = np.array([float(v) for v in x])
    if return_cond:
        Hf = np.array([[float(H[i, j]) for j in range(n)] for i in range(n)])
        return out, float(np.linalg.cond(0.5 * (Hf + Hf.T)))
    return out


def _rnea_mp(model, q, dq, ddq, dps=40):
    import mpmath as mp
    mp.mp.dps = dps
    n = model.n
    M = mp.matrix

    def rot(axis, a):
        c, s = mp.cos(a), mp.sin(a)
        if axis == "x":
            return M([[1, 0, 0], [0, c, -s], [0, s, c]])
        if axis == "y":
            return M([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return M([[c, -s, 0], [s, c, 0], [0, 0, 1]])

    def cross(a, b):
        return M([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def skew(v):
        return M([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])

    R, t, m, h, Io = [], [], [], [], []
    for i in range(n):
        r, p, y = (mp.mpf(float(x)) for x in model.rpy[i])
        R.append(rot("z", y) * rot("y", p) * rot("x", r) * rot("z", mp.mpf(float(q[i]))))
        t.append(M([mp.mpf(float(x)) for x in model.xyz[i]]))
        mi = mp.mpf(float(model.mass[i])); m.append(mi)
        c = M([mp.mpf(float(x)) for x in model.com[i]])
        h.append(mi * c)
        s = [mp.mpf(float(x)) for x in model.inertia6[i]]
        Ic = M([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])
        Cx = skew(c)
        Io.append(Ic + mi * Cx * Cx.T)
    zero = M([0, 0, 0])
    vl, vr, al, ar = zero.copy(), zero.copy(), M([0, 0, mp.mpf(GRAVITY)]), zero.copy()
    fl, fr = [], []
    for i in range(n):
        Rt = R[i].T
        dqi, ddqi = mp.mpf(float(dq[i])), mp.mpf(float(ddq[i]))
        vl, vr = Rt * (vl - cross(t[i], vr)), Rt * vr
        vr[2] += dqi
        al, ar = Rt * (al - cross(t[i], ar)), Rt * ar
        ar[2] += ddqi
        al[0] += vl[1] * dqi; al[1] += -vl[0] * dqi
        ar[0] += vr[1] * dqi; ar[1] += -vr[0] * dqi
        Ial = m[i] * al - cross(h[i], ar); Iar = Io[i] * ar + cross(h[i], al)
        Ivl = m[i] * vl - cross(h[i], vr); Ivr = Io[i] * vr + cross(h[i], vl)
        fl.append(Ial + cross(vr, Ivl))
        fr.append(Iar + cross(vr, Ivr) + cross(vl, Ivl))
    tau = [None] * n
    for i in range(n - 1, -1, -1):
        tau[i] = fr[i][2]
        if i > 0:
            Rl = R[i] * fl[i]
            fl[i - 1] = fl[i - 1] + Rl
            fr[i - 1] = fr[i - 1] + R[i] * fr[i] + cross(t[i], Rl)
    return tau
