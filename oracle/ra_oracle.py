"""ORACLE (test infrastructure, NOT product code) -- CPU/numpy restatement of
the reference's robust rotation averaging,
glomap/estimators/global_rotation_averaging.cc (3-DoF path, trivial rigs).

PARITY UNPINNED: the reference holds no golden vectors for RotationEstimator
and needs Eigen + SuiteSparse + COLMAP (`colmap::LeastAbsoluteDeviationSolver`,
pinned b6b7b54, un-vendored) which cannot be built here.  The ADMM L1 solver is
restated from the public COLMAP source (colmap/optim/least_absolute_deviations.cc,
UPSTREAM-UNVERIFIED): rho = 1, alpha = 1, absolute_tolerance 1e-4,
relative_tolerance 1e-2, the inner iteration cap of 10 that the reference sets
(global_rotation_averaging.cc:484; the doubling at :536-537 mutates a local
options copy after the solver was constructed and does not reach it).

What is restated, with the reference lines it follows:
  * unknowns: angle-axis per frame (.cc:223-225); the first frame is the gauge
    reference (.cc:248-256) with 3 extra rows pinning it (.cc:455-460).
  * A: rows of -I at image 1 and +I at image 2 per edge (.cc:396-415) -- the
    first-order model dR_ij = dR_j - dR_i (header .h:98-100).
  * ComputeResiduals (.cc:696-756): r_e = -log(R_j^T R_rel R_i); gauge rows
    log(R_fixed0^T R_fixed).
  * SolveL1Regression (.cc:479-541): <= max_num_l1_iterations times
    { ADMM on min |W A x - W r|_1 ; UpdateGlobalRotations ; ComputeResiduals },
    stop when the average step < l1_step_convergence_threshold or the step norm
    stalls (|last - cur| < EPS = 1e-12).
  * SolveIRLS (.cc:543-625): weights sigma^2/(e^2+sigma^2)^2 (GEMAN_MCCLURE) or
    (e^2)^(-0.75) (HALF_NORM, .cc:587), gauge rows weight 1 (.cc:557-560);
    step = (A^T W A)^-1 A^T W r by sparse Cholesky every iteration (.cc:603-611).
  * UpdateGlobalRotations (.cc:627-644): theta <- log(exp(theta) exp(-step)).
  * AngleAxisToRotation first-order fallback below 1e-12 (math/rigid3d.cc:45-63);
    RotationToAngleAxis through Eigen's quaternion conversion (math/rigid3d.cc:39-43).
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = 1e-12   # glomap/types.h


@dataclasses.dataclass
class RAOptions:
    """Mirror of RotationEstimatorOptions (global_rotation_averaging.h:39-75)."""
    max_num_l1_iterations: int = 5
    l1_step_convergence_threshold: float = 0.001
    max_num_irls_iterations: int = 100
    irls_step_convergence_threshold: float = 0.001
    irls_loss_parameter_sigma: float = 5.0     # degrees
    weight_type: str = "GEMAN_MCCLURE"         # or "HALF_NORM"
    use_weight: bool = False


def aa_to_R(v):
    """AngleAxisToRotation (math/rigid3d.cc:45-63), batched."""
    v = np.asarray(v, dtype=np.float64)
    n = np.linalg.norm(v, axis=-1)
    small = n <= EPS
    ns = np.where(small, 1.0, n)
    k = v / ns[..., None]
    K = np.zeros(v.shape[:-1] + (3, 3))
    K[..., 0, 1] = -k[..., 2]; K[..., 0, 2] = k[..., 1]
    K[..., 1, 0] = k[..., 2]; K[..., 1, 2] = -k[..., 0]
    K[..., 2, 0] = -k[..., 1]; K[..., 2, 1] = k[..., 0]
    s, c = np.sin(ns)[..., None, None], np.cos(ns)[..., None, None]
    R = np.eye(3) + s * K + (1 - c) * (K @ K)
    if small.any():
        Ks = np.zeros(v.shape[:-1] + (3, 3))
        Ks[..., 0, 1] = -v[..., 2]; Ks[..., 0, 2] = v[..., 1]
        Ks[..., 1, 0] = v[..., 2]; Ks[..., 1, 2] = -v[..., 0]
        Ks[..., 2, 0] = -v[..., 1]; Ks[..., 2, 1] = v[..., 0]
        R = np.where(small[..., None, None], np.eye(3) + Ks, R)
    return R


def R_to_aa(R):
    """RotationToAngleAxis: Eigen AngleAxis(Matrix3) = quaternion conversion,
    angle = 2 atan2(|vec|, |w|) with the axis sign following w."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    n = len(R)
    q = np.empty((n, 4))   # x y z w
    t = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    pos = t > 0
    # Eigen: if (t > 0) {...} else pick i = argmax diag
    tt = np.sqrt(np.where(pos, t, 0) + 1.0)
    q[pos, 3] = 0.5 * tt[pos]
    f = 0.5 / tt[pos]
    q[pos, 0] = (R[pos, 2, 1] - R[pos, 1, 2]) * f
    q[pos, 1] = (R[pos, 0, 2] - R[pos, 2, 0]) * f
    q[pos, 2] = (R[pos, 1, 0] - R[pos, 0, 1]) * f
    idx = np.nonzero(~pos)[0]
    for m in idx:
        M = R[m]
        i = 0
        if M[1, 1] > M[0, 0]:
            i = 1
        if M[2, 2] > M[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        tv = np.sqrt(M[i, i] - M[j, j] - M[k, k] + 1.0)
        qq = np.zeros(4)
        qq[i] = 0.5 * tv
        tv = 0.5 / tv
        qq[3] = (M[k, j] - M[j, k]) * tv
        qq[j] = (M[j, i] + M[i, j]) * tv
        qq[k] = (M[k, i] + M[i, k]) * tv
        q[m] = qq
    nv = np.linalg.norm(q[:, :3], axis=1)
    ang = 2 * np.arctan2(nv, np.abs(q[:, 3]))
    sgn = np.where(q[:, 3] < 0, -1.0, 1.0)
    safe = np.where(nv > 0, nv, 1.0)
    out = q[:, :3] * (sgn * ang / safe)[:, None]
    out[nv == 0] = 0.0
    return out


def compute_residuals(theta, ei, ej, R_rel, fixed, theta_fixed0):
    Ri, Rj = aa_to_R(theta[ei]), aa_to_R(theta[ej])
    r_e = -R_to_aa(np.swapaxes(Rj, -1, -2) @ R_rel @ Ri)
    r_g = R_to_aa(aa_to_R(theta_fixed0[None]).transpose(0, 2, 1) @ aa_to_R(theta[fixed][None]))[0]
    return np.concatenate([r_e.ravel(), r_g])


def build_A(n, ei, ej, fixed):
    E = len(ei)
    rows = np.concatenate([3 * np.arange(E)[:, None] + np.arange(3), 3 * np.arange(E)[:, None] + np.arange(3)]).ravel()
    cols = np.concatenate([3 * ei[:, None] + np.arange(3), 3 * ej[:, None] + np.arange(3)]).ravel()
    vals = np.concatenate([-np.ones(3 * E), np.ones(3 * E)])
    rows = np.concatenate([rows, 3 * E + np.arange(3)])
    cols = np.concatenate([cols, 3 * fixed + np.arange(3)])
    vals = np.concatenate([vals, np.ones(3)])
    return sp.csc_matrix((vals, (rows, cols)), shape=(3 * E + 3, 3 * n))


def l1_admm(A, b, max_iter=10, rho=1.0, alpha=1.0, abs_tol=1e-4, rel_tol=1e-2):
    """colmap::LeastAbsoluteDeviationSolver::Solve (UPSTREAM-UNVERIFIED restatement)."""
    m, n = A.shape
    lu = spla.splu((A.T @ A).tocsc())
    z = np.zeros(m); u = np.zeros(m); x = np.zeros(n)
    b_norm = np.linalg.norm(b)
    eps_pri_thr, eps_dual_thr = np.sqrt(m) * abs_tol, np.sqrt(n) * abs_tol
    its = 0
    for _ in range(max_iter):
        its += 1
        x = lu.solve(A.T @ (b + z - u))
        Ax = A @ x
        Ax_hat = alpha * Ax + (1 - alpha) * (z + b)
        z_old = z
        v = Ax_hat - b + u
        z = np.maximum(0, v - 1 / rho) - np.maximum(0, -v - 1 / rho)
        u = u + Ax_hat - z - b
        r_norm = np.linalg.norm(Ax - z - b)
        s_norm = np.linalg.norm(-rho * (A.T @ (z - z_old)))
        eps_pri = eps_pri_thr + rel_tol * max(b_norm, np.linalg.norm(Ax), np.linalg.norm(z))
        eps_dual = eps_dual_thr + rel_tol * np.linalg.norm(rho * (A.T @ u))
        if r_norm < eps_pri and s_norm < eps_dual:
            break
    return x, its


def update_rotations(theta, step):
    return R_to_aa(aa_to_R(theta) @ aa_to_R(-step)).reshape(theta.shape)


def estimate_rotations(n, ei, ej, R_rel, theta0, edge_weight=None, opts: RAOptions | None = None, fixed=0, verbose=False):
    """Oracle counterpart of RotationEstimator::EstimateRotations after the
    (host-side) initialisation: SetupLinearSystem + SolveL1Regression + SolveIRLS.
    theta0 [n,3] initial angle-axis; returns (theta, info dict)."""
    o = opts or RAOptions()
    ei = np.asarray(ei, dtype=np.int64); ej = np.asarray(ej, dtype=np.int64)
    E = len(ei)
    theta = np.array(theta0, dtype=np.float64)
    theta_fixed0 = theta[fixed].copy()
    A = build_A(n, ei, ej, fixed)
    w_rows = np.ones(3 * E + 3)
    if o.use_weight and edge_weight is not None:
        ew = np.where(np.asarray(edge_weight) >= 0, edge_weight, 1.0)
        w_rows[:3 * E] = np.repeat(ew, 3)
    info = dict(l1_iterations=0, irls_iterations=0, admm_iterations=0)
    res = compute_residuals(theta, ei, ej, R_rel, fixed, theta_fixed0)
    # ---- L1 -------------------------------------------------------------------
    if o.max_num_l1_iterations > 0:
        Aw = sp.diags(w_rows) @ A
        last_norm = curr_norm = 0.0
        for it in range(o.max_num_l1_iterations):
            last_norm = curr_norm
            step, n_admm = l1_admm(Aw.tocsc(), w_rows * res, max_iter=10)
            info["admm_iterations"] += n_admm
            if np.isnan(step).any():
                info["failed"] = True
                return theta, info
            curr_norm = np.linalg.norm(step)
            theta = update_rotations(theta, step.reshape(n, 3))
            res = compute_residuals(theta, ei, ej, R_rel, fixed, theta_fixed0)
            info["l1_iterations"] += 1
            avg = np.linalg.norm(step.reshape(n, 3), axis=1).sum() / n
            if verbose:
                print(f"  L1 {it}: avg step {avg:.3e} |res|_1 {np.abs(res).sum():.4e}")
            if avg < o.l1_step_convergence_threshold or abs(last_norm - curr_norm) < EPS:
                break
    # ---- IRLS -----------------------------------------------------------------
    if o.max_num_irls_iterations > 0:
        sigma = np.radians(o.irls_loss_parameter_sigma)
        for it in range(o.max_num_irls_iterations):
            err2 = (res[:3 * E].reshape(E, 3) ** 2).sum(1)
            if o.weight_type == "GEMAN_MCCLURE":
                tmp = err2 + sigma * sigma
                w = sigma * sigma / (tmp * tmp)
            else:
                with np.errstate(divide="ignore"):
                    w = err2 ** ((0.5 - 2) / 2)
            if np.isnan(w).any():
                info["failed"] = True
                return theta, info
            w_irls = np.concatenate([np.repeat(w, 3), np.ones(3)])
            W = sp.diags(w_irls * w_rows)
            AtW = A.T @ W
            step = spla.splu((AtW @ A).tocsc()).solve(AtW @ res)
            theta = update_rotations(theta, step.reshape(n, 3))
            res = compute_residuals(theta, ei, ej, R_rel, fixed, theta_fixed0)
            info["irls_iterations"] += 1
            avg = np.linalg.norm(step.reshape(n, 3), axis=1).sum() / n
            if verbose:
                print(f"  IRLS {it}: avg step {avg:.3e}")
            if avg < o.irls_step_convergence_threshold:
                break
    return theta, info


def max_pairwise_rotation_error_deg(theta, R_gt):
    """All-pairs relative rotation error (rotation_averager_test.cc:85-106)."""
    R = aa_to_R(theta)
    n = len(R)
    worst = 0.0
    for i in range(n):
        Rel = R @ R[i].T
        Rel_gt = R_gt @ R_gt[i].T
        M = np.swapaxes(Rel, -1, -2) @ Rel_gt
        c = np.clip((np.trace(M, axis1=-2, axis2=-1) - 1) / 2, -1, 1)
        worst = max(worst, float(np.degrees(np.arccos(c)).max()))
    return worst


# ---------------------------------------------------------------------------
# Gravity-aligned (1-DoF) frames -- options_.use_gravity
# ---------------------------------------------------------------------------
def get_align_rot(gravity):
    """GetAlignRot (math/gravity.cc:11-24): a rotation whose second column is
    the (normalised) gravity direction.  The other two columns are any
    orthonormal completion (the reference takes them from a Householder QR;
    a different completion only shifts the 1-DoF angle by a constant)."""
    v = np.asarray(gravity, dtype=np.float64)
    v = v / np.linalg.norm(v)
    a = np.array([1.0, 0, 0]) if abs(v[0]) < 0.9 else np.array([0, 0, 1.0])
    x = np.cross(v, a); x /= np.linalg.norm(x)
    z = np.cross(x, v)
    R = np.stack([x, v, z], axis=1)
    if np.linalg.det(R) < 0:
        R[:, 2] = -R[:, 2]
    return R


def rel_angle_error(angle_12, angle_1, angle_2):
    """RelAngleError (global_rotation_averaging.cc:19-36) without the rand()
    jitter near +-pi (which is not reproducible in the reference either)."""
    est = (angle_2 - angle_1) - angle_12
    return (est + np.pi) % (2 * np.pi) - np.pi


def estimate_rotations_gravity(n, ei, ej, R_rel, R0, has_gravity, R_align, opts: RAOptions | None = None, verbose=False):
    """use_gravity path (trivial rigs): frames with gravity carry ONE unknown
    (the angle about the gravity axis, .cc:207-217), the others three.
    R0 [n,3,3] initial rotations, R_align [n,3,3] (identity where no gravity).
    Rows follow SetupLinearSystem (.cc:386-421), residuals ComputeResiduals
    (.cc:709-743), weights SolveIRLS (.cc:574-580), update (.cc:634-643).
    Returns (R [n,3,3], info)."""
    o = opts or RAOptions()
    ei = np.asarray(ei, dtype=np.int64); ej = np.asarray(ej, dtype=np.int64)
    hg = np.asarray(has_gravity, dtype=bool)
    E = len(ei)
    # unknown layout
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.where(hg, 1, 3))
    ndof = int(off[-1])
    est = np.zeros(ndof)
    for i in range(n):
        if hg[i]:
            est[off[i]] = R_to_aa((R_align[i].T @ R0[i])[None])[0, 1]        # RotUpToAngle (.cc:208-210)
        else:
            est[off[i]:off[i] + 3] = R_to_aa(R0[i][None])[0]
    # gravity-aligned relative rotations (.cc:311-340)
    Rr = np.array(R_rel, dtype=np.float64, copy=True)
    for e in range(E):
        if hg[ei[e]]:
            Rr[e] = Rr[e] @ R_align[ei[e]]
        if hg[ej[e]]:
            Rr[e] = R_align[ej[e]].T @ Rr[e]
    both = hg[ei] & hg[ej]
    aa = R_to_aa(Rr)
    xz_error = np.where(both, aa[:, 0] ** 2 + aa[:, 2] ** 2, 0.0)
    angle_rel = aa[:, 1]
    # fixed camera: first frame with gravity, else first frame (.cc:213-217,248-256)
    fixed = int(np.nonzero(hg)[0][0]) if hg.any() else 0
    fixed0 = est[off[fixed]:off[fixed + 1]].copy()
    rows, cols, vals = [], [], []
    row_of = np.zeros(E, dtype=np.int64)
    pos = 0
    for e in range(E):
        i, j = ei[e], ej[e]
        row_of[e] = pos
        if both[e]:
            rows += [pos, pos]; cols += [off[i], off[j]]; vals += [-1, 1]; pos += 1
        else:
            if not hg[i]:
                rows += [pos, pos + 1, pos + 2]; cols += [off[i], off[i] + 1, off[i] + 2]; vals += [-1, -1, -1]
            else:
                rows.append(pos + 1); cols.append(off[i]); vals.append(-1)
            if not hg[j]:
                rows += [pos, pos + 1, pos + 2]; cols += [off[j], off[j] + 1, off[j] + 2]; vals += [1, 1, 1]
            else:
                rows.append(pos + 1); cols.append(off[j]); vals.append(1)
            pos += 3
    gauge_row = pos
    if hg[fixed]:
        rows.append(pos); cols.append(off[fixed]); vals.append(1); pos += 1
    else:
        rows += [pos, pos + 1, pos + 2]; cols += [off[fixed], off[fixed] + 1, off[fixed] + 2]; vals += [1, 1, 1]; pos += 3
    m = pos
    A = sp.csc_matrix((vals, (rows, cols)), shape=(m, ndof))

    def node_R(i):
        return aa_to_R(np.array([[0, est[off[i]], 0]]))[0] if hg[i] else aa_to_R(est[off[i]:off[i] + 3][None])[0]

    def residuals():
        r = np.zeros(m)
        for e in range(E):
            i, j = ei[e], ej[e]
            if both[e]:
                r[row_of[e]] = rel_angle_error(angle_rel[e], est[off[i]], est[off[j]])
            else:
                r[row_of[e]:row_of[e] + 3] = -R_to_aa((node_R(j).T @ Rr[e] @ node_R(i))[None])[0]
        if hg[fixed]:
            r[gauge_row] = est[off[fixed]] - fixed0[0]
        else:
            r[gauge_row:gauge_row + 3] = R_to_aa((aa_to_R(fixed0[None])[0].T @ node_R(fixed))[None])[0]
        return r

    def update(step):
        for i in range(n):
            if hg[i]:
                est[off[i]] -= step[off[i]]
            else:
                Rn = aa_to_R(est[off[i]:off[i] + 3][None])[0] @ aa_to_R(-step[off[i]:off[i] + 3][None])[0]
                est[off[i]:off[i] + 3] = R_to_aa(Rn[None])[0]

    def avg_step(step):
        return sum(abs(step[off[i]]) if hg[i] else np.linalg.norm(step[off[i]:off[i] + 3]) for i in range(n)) / n

    info = dict(l1_iterations=0, irls_iterations=0, admm_iterations=0)
    res = residuals()
    if o.max_num_l1_iterations > 0:
        last_norm = curr_norm = 0.0
        for it in range(o.max_num_l1_iterations):
            last_norm = curr_norm
            step, n_admm = l1_admm(A, res, max_iter=10)
            info["admm_iterations"] += n_admm
            curr_norm = np.linalg.norm(step)
            update(step)
            res = residuals()
            info["l1_iterations"] += 1
            if avg_step(step) < o.l1_step_convergence_threshold or abs(last_norm - curr_norm) < EPS:
                break
    if o.max_num_irls_iterations > 0:
        sigma = np.radians(o.irls_loss_parameter_sigma)
        for it in range(o.max_num_irls_iterations):
            w_rows = np.ones(m)
            for e in range(E):
                if both[e]:
                    err2 = res[row_of[e]] ** 2 + xz_error[e]
                    nrow = 1
                else:
                    err2 = (res[row_of[e]:row_of[e] + 3] ** 2).sum()
                    nrow = 3
                if o.weight_type == "GEMAN_MCCLURE":
                    tmp = err2 + sigma * sigma
                    w = sigma * sigma / (tmp * tmp)
                else:
                    w = err2 ** ((0.5 - 2) / 2)
                w_rows[row_of[e]:row_of[e] + nrow] = w
            AtW = A.T @ sp.diags(w_rows)
            step = spla.splu((AtW @ A).tocsc()).solve(AtW @ res)
            update(step)
            res = residuals()
            info["irls_iterations"] += 1
            if avg_step(step) < o.irls_step_convergence_threshold:
                break
    R = np.stack([R_align[i] @ node_R(i) if hg[i] else node_R(i) for i in range(n)])   # ConvertResults (.cc:787-799)
    info["fixed"] = fixed
    return R, info


# ---------------------------------------------------------------------------
# Unknown cam_from_rig rotations (non-trivial rigs whose sensors are not calibrated yet)
# ---------------------------------------------------------------------------
def average_quaternions(R_list):
    """colmap::AverageQuaternions with unit weights (UPSTREAM-UNVERIFIED restatement): the eigenvector of
    sum q q^T with the largest eigenvalue, returned as a rotation matrix."""
    qs = []
    for R in R_list:
        # Eigen::Quaterniond(Matrix3d): the same branch structure as R_to_aa's first half
        t = np.trace(R)
        q = np.zeros(4)
        if t > 0:
            tt = np.sqrt(t + 1.0); q[3] = 0.5 * tt; tt = 0.5 / tt
            q[0] = (R[2, 1] - R[1, 2]) * tt; q[1] = (R[0, 2] - R[2, 0]) * tt; q[2] = (R[1, 0] - R[0, 1]) * tt
        else:
            i = 0
            if R[1, 1] > R[0, 0]:
                i = 1
            if R[2, 2] > R[i, i]:
                i = 2
            j, k = (i + 1) % 3, (i + 2) % 3
            tt = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
            q[i] = 0.5 * tt; tt = 0.5 / tt
            q[3] = (R[k, j] - R[j, k]) * tt; q[j] = (R[j, i] + R[i, j]) * tt; q[k] = (R[k, i] + R[i, k]) * tt
        qs.append(q)
    qs = np.array(qs)
    w, v = np.linalg.eigh(qs.T @ qs)
    x, y, z, ww = v[:, -1] / np.linalg.norm(v[:, -1])
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                     [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                     [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])


def estimate_rotations_rig_unknown(n_frames, n_cams, ei, ej, eci, ecj, R_rel, theta0, cam_frames, opts: RAOptions | None = None,
                                   fixed=0, edge_weight=None):
    """RotationEstimator with unknown cam_from_rig rotations (global_rotation_averaging.cc:173-245 unknown layout,
    :425-440 rows, :646-693 update with quaternion averaging, :726-736 residuals).
      unknowns: theta [n_frames + n_cams, 3] = frame rotations followed by the cam_from_rig rotations of the sensors that
      are not calibrated; edge e: frames (ei, ej), unknown-camera nodes (eci, ecj) as indices into theta or -1;
      R_rel already carries the KNOWN cam_from_rig factors (:305-309); cam_frames[c] = frames that hold an image of
      unknown camera c (:660-671).  Row e of A: -I at ei, +I at ej, -I at eci, +I at ecj (same-frame pairs cancel)."""
    o = opts or RAOptions()
    ei = np.asarray(ei, np.int64); ej = np.asarray(ej, np.int64); eci = np.asarray(eci, np.int64); ecj = np.asarray(ecj, np.int64)
    E, n = len(ei), n_frames + n_cams
    theta = np.array(theta0, dtype=np.float64)
    theta_fixed0 = theta[fixed].copy()
    rows, cols, vals = [], [], []
    for nodes, sgn in ((ei, -1.0), (ej, 1.0), (eci, -1.0), (ecj, 1.0)):
        m = nodes >= 0
        idx = np.nonzero(m)[0]
        for k in range(3):
            rows.append(3 * idx + k); cols.append(3 * nodes[m] + k); vals.append(np.full(len(idx), sgn))
    rows.append(3 * E + np.arange(3)); cols.append(3 * fixed + np.arange(3)); vals.append(np.ones(3))
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * E + 3, 3 * n))
    A.sum_duplicates()
    w_rows = np.ones(3 * E + 3)
    if o.use_weight and edge_weight is not None:
        w_rows[:3 * E] = np.repeat(np.where(np.asarray(edge_weight) >= 0, edge_weight, 1.0), 3)

    def residuals(th):
        R1, R2 = aa_to_R(th[ei]), aa_to_R(th[ej])
        m = eci >= 0
        if m.any():
            R1[m] = aa_to_R(th[eci[m]]) @ R1[m]          # R_1 = R_cam1 R_frame1 (:726-730)
        m = ecj >= 0
        if m.any():
            R2[m] = aa_to_R(th[ecj[m]]) @ R2[m]
        r_e = -R_to_aa(np.swapaxes(R2, -1, -2) @ R_rel @ R1)
        r_g = R_to_aa(aa_to_R(theta_fixed0[None]).transpose(0, 2, 1) @ aa_to_R(th[fixed][None]))[0]
        return np.concatenate([r_e.ravel(), r_g])

    def update(th, step):
        out = th.copy()
        out[:n_frames] = R_to_aa(aa_to_R(th[:n_frames]) @ aa_to_R(-step[:n_frames]))      # frames first (:631-644)
        Rf = aa_to_R(out[:n_frames])
        for c in range(n_cams):                                                              # :675-693
            R_ori = aa_to_R(th[n_frames + c][None])[0]
            R_upd = aa_to_R(-step[n_frames + c][None])[0]
            prods = [R_ori @ Rf[f] @ R_upd @ Rf[f].T for f in cam_frames[c]]
            out[n_frames + c] = R_to_aa(average_quaternions(prods)[None])[0]
        return out

    info = dict(l1_iterations=0, irls_iterations=0, admm_iterations=0)
    res = residuals(theta)
    if o.max_num_l1_iterations > 0:
        Aw = (sp.diags(w_rows) @ A).tocsc()
        last_norm = curr_norm = 0.0
        for it in range(o.max_num_l1_iterations):
            last_norm = curr_norm
            step, n_admm = l1_admm(Aw, w_rows * res, max_iter=10)
            info["admm_iterations"] += n_admm
            curr_norm = np.linalg.norm(step)
            st = step.reshape(n, 3)
            theta = update(theta, st)
            res = residuals(theta)
            info["l1_iterations"] += 1
            avg = np.linalg.norm(st[:n_frames], axis=1).sum() / n_frames                  # frames only (:758-772)
            if avg < o.l1_step_convergence_threshold or abs(last_norm - curr_norm) < EPS:
                break
    if o.max_num_irls_iterations > 0:
        sigma = np.radians(o.irls_loss_parameter_sigma)
        for it in range(o.max_num_irls_iterations):
            err2 = (res[:3 * E].reshape(E, 3) ** 2).sum(1)
            if o.weight_type == "GEMAN_MCCLURE":
                tmp = err2 + sigma * sigma
                w = sigma * sigma / (tmp * tmp)
            else:
                with np.errstate(divide="ignore"):
                    w = err2 ** ((0.5 - 2) / 2)
            W = sp.diags(np.concatenate([np.repeat(w, 3), np.ones(3)]) * w_rows)
            AtW = A.T @ W
            step = spla.splu((AtW @ A).tocsc()).solve(AtW @ res)
            st = step.reshape(n, 3)
            theta = update(theta, st)
            res = residuals(theta)
            info["irls_iterations"] += 1
            if np.linalg.norm(st[:n_frames], axis=1).sum() / n_frames < o.irls_step_convergence_threshold:
                break
    return theta, info
