"""Independent inverse dynamics (classical Newton-Euler, world frame, numpy float64) used to pin
the ABA oracle: given a state and the accelerations the oracle produced, the joint forces that
Newton-Euler says are REQUIRED must equal the joint forces that were APPLIED.
Written from the textbook recursion; shares no code or formulation with oracle/aba_oracle.c."""
import numpy as np
from isaacgymenvs_b200.importer import rot


def kinematics(m, root, q, qd, qdd, root_acc):
    """root: (13,) pos quat linvel angvel; root_acc: (dv(3), dw(3)) classical, world.
    Returns per-link dicts of world R, x, w, alpha, a (origin accel), axis_w."""
    nl = m.nl
    R = [None] * nl; x = [None] * nl; w = [None] * nl; al = [None] * nl; a = [None] * nl; axw = [None] * nl
    v = [None] * nl
    R[0] = rot.quat_to_mat(root[3:7]); x[0] = root[0:3].copy()
    if m.root_fixed:
        v[0] = np.zeros(3); w[0] = np.zeros(3); a[0] = np.zeros(3); al[0] = np.zeros(3)
    else:
        v[0] = root[7:10].copy(); w[0] = root[10:13].copy(); a[0] = root_acc[0].copy(); al[0] = root_acc[1].copy()
    for i in range(1, nl):
        p = m.parent[i]
        Rl = rot.quat_to_mat(m.lquat[i])
        ax_p = Rl @ m.axis[i]                     # axis in parent link coords (fixed there)
        wa = R[p] @ ax_p
        axw[i] = wa
        qi, qdi, qddi = q[i - 1], qd[i - 1], qdd[i - 1]
        if m.jtype[i] == 0:
            R[i] = R[p] @ Rl @ rot.axis_angle_to_mat(m.axis[i], qi)
            d = R[p] @ m.lpos[i]
            x[i] = x[p] + d
            w[i] = w[p] + wa * qdi
            al[i] = al[p] + wa * qddi + np.cross(w[p], wa * qdi)
            v[i] = v[p] + np.cross(w[p], d)
            a[i] = a[p] + np.cross(al[p], d) + np.cross(w[p], np.cross(w[p], d))
        else:
            R[i] = R[p] @ Rl
            d = R[p] @ (m.lpos[i] + ax_p * qi)
            x[i] = x[p] + d
            w[i] = w[p].copy(); al[i] = al[p].copy()
            v[i] = v[p] + np.cross(w[p], d) + wa * qdi
            a[i] = a[p] + np.cross(al[p], d) + np.cross(w[p], np.cross(w[p], d)) + 2 * np.cross(w[p], wa * qdi) + wa * qddi
    return R, x, v, w, al, a, axw


def inverse_dynamics(m, root, q, qd, qdd, root_acc, gravity, ext=None):
    """Returns (tau (nd,), root_wrench (force3, torque3 about root origin)) REQUIRED to produce the
    given accelerations.  ext: optional {link: (force_world, point_world)} list of external forces."""
    nl = m.nl
    R, x, v, w, al, a, axw = kinematics(m, root, q, qd, qdd, root_acc)
    f = [np.zeros(3) for _ in range(nl)]     # force parent exerts on link i (at link origin)
    n = [np.zeros(3) for _ in range(nl)]     # torque about link origin
    g = np.asarray(gravity, float)
    for i in range(nl):
        c = R[i] @ m.com[i]
        Iw = R[i] @ rot.sym6_to_mat(m.inertia[i]) @ R[i].T
        ac = a[i] + np.cross(al[i], c) + np.cross(w[i], np.cross(w[i], c))
        F = m.mass[i] * ac - m.mass[i] * g
        N = Iw @ al[i] + np.cross(w[i], Iw @ w[i])
        # AssetOptions.angular_damping / linear_damping: an external wrench -d_a Iw w, -d_l m v_c at the COM (classical form)
        da, dl = float(getattr(m, "angular_damping", 0.0) or 0.0), float(getattr(m, "linear_damping", 0.0) or 0.0)
        F = F + dl * m.mass[i] * (v[i] + np.cross(w[i], c))
        N = N + da * (Iw @ w[i])
        f[i] = f[i] + F
        n[i] = n[i] + N + np.cross(c, F)
        if ext:
            for (li, Fe, pe) in ext:
                if li == i:
                    f[i] = f[i] - Fe
                    n[i] = n[i] - np.cross(pe - x[i], Fe)
    for i in range(nl - 1, 0, -1):
        p = m.parent[i]
        f[p] = f[p] + f[i]
        n[p] = n[p] + n[i] + np.cross(x[i] - x[p], f[i])
    tau = np.zeros(nl - 1)
    for i in range(1, nl):
        tau[i - 1] = axw[i] @ (n[i] if m.jtype[i] == 0 else f[i])
    return tau, (f[0], n[0])


def momentum(m, root, q, qd):
    """Total linear momentum and angular momentum about the world origin."""
    z = np.zeros(m.nl - 1)
    R, x, v, w, al, a, axw = kinematics(m, root, q, qd, z, (np.zeros(3), np.zeros(3)))
    P = np.zeros(3); L = np.zeros(3)
    for i in range(m.nl):
        c = R[i] @ m.com[i]
        vc = v[i] + np.cross(w[i], c)
        Iw = R[i] @ rot.sym6_to_mat(m.inertia[i]) @ R[i].T
        P += m.mass[i] * vc
        L += Iw @ w[i] + np.cross(x[i] + c, m.mass[i] * vc)
    return P, L
