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
