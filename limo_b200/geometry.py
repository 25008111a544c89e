"""Minimal rigid-body helpers following the Eigen conventions the reference relies on.

Poses are 7-vectors [qw, qx, qy, qz, tx, ty, tz] (internal/definitions.hpp:23) meaning p_dst = R(q) p_src + t
(definitions.hpp:75-83).  4x4 matrices are used where the reference uses Eigen::Isometry3d.
"""
import numpy as np


def quat_to_rot(q):
    """Eigen::Quaternion::toRotationMatrix (no normalisation), q = (w, x, y, z)."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def rot_to_quat(m):
    """Eigen::Quaternion(Matrix3) (definitions.cpp:14-28 uses it through Quaterniond(p.rotation()))."""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)  # w, x, y, z
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def angle_axis(angle, axis):
    """Eigen::AngleAxisd(angle, axis).toRotationMatrix() (axis assumed normalised, as Eigen does)."""
    ax = np.asarray(axis, dtype=float)
    s, c = np.sin(angle), np.cos(angle)
    cc = (1 - c) * ax
    R = np.empty((3, 3))
    tmp = cc[0] * ax[1]; R[0, 1] = tmp - s * ax[2]; R[1, 0] = tmp + s * ax[2]
    tmp = cc[0] * ax[2]; R[0, 2] = tmp + s * ax[1]; R[2, 0] = tmp - s * ax[1]
    tmp = cc[1] * ax[2]; R[1, 2] = tmp - s * ax[0]; R[2, 1] = tmp + s * ax[0]
    R[0, 0] = cc[0] * ax[0] + c; R[1, 1] = cc[1] * ax[1] + c; R[2, 2] = cc[2] * ax[2] + c
    return R


def iso(R=None, t=None):
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if t is not None:
        T[:3, 3] = t
    return T


def translate(T, v):
    """Eigen Transform::translate (post-multiplication)."""
    return T @ iso(t=np.asarray(v, dtype=float))


def rotate(T, R):
    """Eigen Transform::rotate (post-multiplication)."""
    return T @ iso(R=R)


def iso_inv(T):
    """Isometry inverse: R^T, -R^T t."""
    R, t = T[:3, :3], T[:3, 3]
    return iso(R.T, -R.T @ t)


def pose_to_iso(p):
    """convert(const T* pose) (definitions.hpp:75-83)."""
    return iso(quat_to_rot(p[:4]), np.asarray(p[4:7], dtype=float))


def iso_to_pose(T):
    """convert(EigenPose) (definitions.cpp:14-28)."""
    return np.concatenate([rot_to_quat(T[:3, :3]), T[:3, 3]])


def apply(T, p):
    return T[:3, :3] @ np.asarray(p, dtype=float) + T[:3, 3]


def is_approx(A, B, prec):
    """Eigen DenseBase::isApprox on the 4x4 matrices: |A-B|_F^2 <= prec^2 min(|A|_F^2, |B|_F^2)."""
    return np.sum((A - B) ** 2) <= prec * prec * min(np.sum(A * A), np.sum(B * B))


def quaternion_angle(p0, p1):
    """calcQuaternionDiff (definitions.cpp:104-111): angle of AngleAxis(q1^-1 * q0)."""
    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                         a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                         a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    q0, q1 = np.asarray(p0[:4], dtype=float), np.asarray(p1[:4], dtype=float)
    q1i = np.array([q1[0], -q1[1], -q1[2], -q1[3]]) / np.dot(q1, q1)
    q = qmul(q1i, q0)
    n = np.linalg.norm(q[1:])
    if n == 0.0:
        return 0.0
    return 2.0 * np.arctan2(n, abs(q[0]))  # Eigen::AngleAxis(Quaternion): angle = 2 atan2(|vec|, |w|)
