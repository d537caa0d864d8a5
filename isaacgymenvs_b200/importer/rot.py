"""Small float64 rotation helpers for the asset importer (host side, one-off).

Quaternions are xyzw everywhere in this repo, the convention the reference's state tensors use
(`isaacgymenvs/utils/torch_jit_utils.py:41-62` quat_mul reads w at index 3).
"""
import numpy as np


def quat_wxyz_to_xyzw(q):
    q = np.asarray(q, dtype=np.float64)
    return np.array([q[1], q[2], q[3], q[0]])


def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_to_mat(q):
    x, y, z, w = quat_normalize(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def mat_to_quat(R):
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    q = quat_normalize(np.array(q))
    return q if q[3] >= 0 else -q


def axis_angle_to_mat(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = skew(a)
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def rpy_to_mat(rpy):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(y) Ry(p) Rx(r)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def euler_xyz_intrinsic_to_mat(e):
    """MuJoCo default eulerseq 'xyz' (intrinsic): R = Rx(a) Ry(b) Rz(c)."""
    a, b, c = e
    return axis_angle_to_mat([1, 0, 0], a) @ axis_angle_to_mat([0, 1, 0], b) @ axis_angle_to_mat([0, 0, 1], c)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def zaxis_to_mat(z):
    """Minimal rotation taking (0,0,1) to the direction z (MuJoCo `zaxis` / `fromto` convention)."""
    z = np.asarray(z, dtype=np.float64)
    z = z / np.linalg.norm(z)
    e = np.array([0.0, 0.0, 1.0])
    c = float(e @ z)
    if c > 1 - 1e-12:
        return np.eye(3)
    if c < -1 + 1e-12:
        return np.diag([1.0, -1.0, -1.0])
    ax = np.cross(e, z)
    return axis_angle_to_mat(ax, np.arccos(np.clip(c, -1, 1)))


def sym6_to_mat(s):
    """(xx, yy, zz, xy, xz, yz) -> 3x3."""
    return np.array([[s[0], s[3], s[4]], [s[3], s[1], s[5]], [s[4], s[5], s[2]]], dtype=np.float64)


def mat_to_sym6(M):
    return np.array([M[0, 0], M[1, 1], M[2, 2], M[0, 1], M[0, 2], M[1, 2]], dtype=np.float64)
