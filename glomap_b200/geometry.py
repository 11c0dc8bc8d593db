"""Small numpy SO(3)/Sim(3) helpers used by the synthetic generators, the
host-side wrappers and the tests (NOT by the solver, which is CUDA).

Conventions follow the reference: ``cam_from_world`` is ``X_c = R X_w + t``;
quaternions cross the C ABI in Eigen ``coeffs()`` order (x, y, z, w)
(reference: glomap/estimators/bundle_adjustment.cc:143); text files use
Hamilton w-first (glomap/io/pose_io.cc:64-68).
"""
from __future__ import annotations

import numpy as np


def quat_xyzw_to_rotmat(q: np.ndarray) -> np.ndarray:
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotmat_to_quat_xyzw(R: np.ndarray) -> np.ndarray:
    """Batched rotation matrix -> unit quaternion (x,y,z,w), w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    batch = R.shape[:-2]
    Rf = R.reshape(-1, 3, 3)
    n = Rf.shape[0]
    q = np.empty((n, 4))
    tr = Rf[:, 0, 0] + Rf[:, 1, 1] + Rf[:, 2, 2]
    for i in range(n):
        m = Rf[i]
        t = tr[i]
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q[i] = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q[i] = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q[i] = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q[i] = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1
    return q.reshape(batch + (4,))


def rotmat_to_quat_xyzw_fast(R: np.ndarray) -> np.ndarray:
    """Vectorised variant (no python loop) for large batches."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    qw = np.sqrt(np.maximum(0, 1 + m00 + m11 + m22)) / 2
    qx = np.sqrt(np.maximum(0, 1 + m00 - m11 - m22)) / 2
    qy = np.sqrt(np.maximum(0, 1 - m00 + m11 - m22)) / 2
    qz = np.sqrt(np.maximum(0, 1 - m00 - m11 + m22)) / 2
    # Pick the largest component as pivot for sign recovery.
    comps = np.stack([qx, qy, qz, qw], axis=1)
    piv = np.argmax(comps, axis=1)
    q = np.empty_like(comps)
    for p in range(4):
        sel = piv == p
        if not sel.any():
            continue
        Rs = R[sel]
        c = comps[sel, p]
        if p == 3:
            q[sel] = np.stack([(Rs[:, 2, 1] - Rs[:, 1, 2]) / (4 * c), (Rs[:, 0, 2] - Rs[:, 2, 0]) / (4 * c),
                               (Rs[:, 1, 0] - Rs[:, 0, 1]) / (4 * c), c], axis=1)
        elif p == 0:
            q[sel] = np.stack([c, (Rs[:, 0, 1] + Rs[:, 1, 0]) / (4 * c), (Rs[:, 0, 2] + Rs[:, 2, 0]) / (4 * c),
                               (Rs[:, 2, 1] - Rs[:, 1, 2]) / (4 * c)], axis=1)
        elif p == 1:
            q[sel] = np.stack([(Rs[:, 0, 1] + Rs[:, 1, 0]) / (4 * c), c, (Rs[:, 1, 2] + Rs[:, 2, 1]) / (4 * c),
                               (Rs[:, 0, 2] - Rs[:, 2, 0]) / (4 * c)], axis=1)
        else:
            q[sel] = np.stack([(Rs[:, 0, 2] + Rs[:, 2, 0]) / (4 * c), (Rs[:, 1, 2] + Rs[:, 2, 1]) / (4 * c), c,
                               (Rs[:, 1, 0] - Rs[:, 0, 1]) / (4 * c)], axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1
    return q


def skew(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]
    S[..., 2, 1] = v[..., 0]
    return S


def so3_exp(w: np.ndarray) -> np.ndarray:
    """Rodrigues, batched, exact series below 1e-8."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th * th / 6.0, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th * th / 24.0, (1 - np.cos(ths)) / (ths * ths))
    K = skew(w)
    I = np.broadcast_to(np.eye(3), K.shape)
    return I + a[..., None, None] * K + b[..., None, None] * (K @ K)


def so3_log(R: np.ndarray) -> np.ndarray:
    """Batched log map via quaternion (robust near pi), angle in [0, pi]."""
    q = rotmat_to_quat_xyzw_fast(np.asarray(R).reshape(-1, 3, 3))
    n = np.linalg.norm(q[:, :3], axis=1)
    ang = 2 * np.arctan2(n, q[:, 3])
    scale = np.where(n < 1e-15, 2.0, ang / np.where(n < 1e-15, 1.0, n))
    out = q[:, :3] * scale[:, None]
    return out.reshape(np.asarray(R).shape[:-2] + (3,))


def rotation_angle_deg(Ra: np.ndarray, Rb: np.ndarray) -> np.ndarray:
    """Angle of Ra^T Rb in degrees (reference: glomap/math/rigid3d.cc:22-27)."""
    M = np.swapaxes(Ra, -1, -2) @ Rb
    c = (np.trace(M, axis1=-2, axis2=-1) - 1) / 2
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


def centers_from_pose(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """c = -R^T t  (reference: glomap/math/rigid3d.cc:65-67)."""
    return -np.einsum("...ji,...j->...i", R, t)


def umeyama_sim3(src: np.ndarray, dst: np.ndarray):
    """Least-squares Sim3 (s, R, t) with dst ~= s R src + t.  This mirrors the
    compare-after-alignment methodology of the reference's tests
    (global_mapper_test.cc:27-33, AlignReconstructionsViaProjCenters)."""
    mu_s, mu_d = src.mean(0), dst.mean(0)
    xs, xd = src - mu_s, dst - mu_d
    cov = xd.T @ xs / len(src)
    U, D, Vt = np.linalg.svd(cov)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    var_s = (xs ** 2).sum() / len(src)
    s = np.trace(np.diag(D) @ S) / var_s
    t = mu_d - s * R @ mu_s
    return s, R, t


def compare_reconstructions(Ra, ta, Rb, tb):
    """Align (a) onto (b) through projection centres and return
    (max rotation error [deg], max centre error).  Rotations are compared as
    cam_from_world after rotating world (a) into world (b)."""
    ca, cb = centers_from_pose(Ra, ta), centers_from_pose(Rb, tb)
    s, R, t = umeyama_sim3(ca, cb)
    ca_al = (s * (R @ ca.T)).T + t
    Ra_al = Ra @ R.T
    rot_err = rotation_angle_deg(Ra_al, Rb)
    cen_err = np.linalg.norm(ca_al - cb, axis=1)
    return float(rot_err.max()), float(cen_err.max()), (s, R, t)
