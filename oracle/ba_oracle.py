"""ORACLE (test infrastructure, NOT product code) -- CPU/numpy restatement of
the reference's global bundle adjustment, glomap/estimators/bundle_adjustment.cc.

PARITY UNPINNED (see oracle/ceres_lm.py): neither Ceres nor COLMAP
(`colmap::ReprojErrorCostFunctor`, pinned b6b7b54, un-vendored) can be built
here and the reference holds no golden vectors for BundleAdjuster.

What is restated, with the reference lines it follows:
  * residual per observation r = ImgFromCam(params, q (x) X + t) - xy
    (bundle_adjustment.cc:135-146; colmap ReprojErrorCostFunctor,
    UPSTREAM-UNVERIFIED).  Observations with z <= eps contribute zero.
  * tracks with fewer than min_num_view_per_track observations are skipped
    (bundle_adjustment.cc:122).
  * Huber(thres_loss_function = 1.0) on every residual block
    (bundle_adjustment.h:29-35, .cc:112,142).
  * EigenQuaternionManifold on every frame rotation (.cc:258): Plus(q, d) =
    [sin|d| d/|d|, cos|d|] (x) q, i.e. a LEFT perturbation by angle 2|d|.
  * the first frame is held constant, rotations and/or translations constant
    when optimize_rotations / optimize_translation are off (.cc:261-266);
    points constant when optimize_points is off (.cc:310-316).
  * intrinsics: constant (optimize_intrinsics = false) or optimised with the
    principal point held fixed (SubsetManifold, .cc:273-286).
  * Ceres LM + exact solve (SPARSE_SCHUR, .cc:95) -> oracle/ceres_lm.py.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp

from .ceres_lm import LMOptions, LMSummary, huber_rho, solve_lm

SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL = 0, 1, 2, 3
# (focal idxs, principal point idxs, extra idxs) per COLMAP model
MODEL_LAYOUT = {
    SIMPLE_PINHOLE: ((0,), (1, 2), ()),
    PINHOLE: ((0, 1), (2, 3), ()),
    SIMPLE_RADIAL: ((0,), (1, 2), (3,)),
    RADIAL: ((0,), (1, 2), (3, 4)),
}
Z_EPS = 1e-12


@dataclasses.dataclass
class BAOptions:
    """Mirror of BundleAdjusterOptions (bundle_adjustment.h:12-37)."""
    optimize_rotations: bool = True
    optimize_translation: bool = True
    optimize_intrinsics: bool = False
    optimize_principal_point: bool = False
    optimize_points: bool = True
    optimize_rig_poses: bool = False
    thres_loss_function: float = 1.0
    min_num_view_per_track: int = 3
    max_num_iterations: int = 200
    function_tolerance: float = 1e-5


def quat_rotmat(q):
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def quat_mul(a, b):
    """Hamilton product of xyzw quaternions."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def quat_plus(q, d):
    """EigenQuaternionManifold::Plus: q_new = exp(d) (x) q with
    exp(d) = [sin|d| d/|d|, cos|d|] (Ceres manifold.h, UPSTREAM-UNVERIFIED)."""
    n = np.linalg.norm(d, axis=-1, keepdims=True)
    small = n < 1e-300
    sn = np.where(small, 1.0, np.sin(n) / np.where(small, 1.0, n))
    dq = np.concatenate([sn * d, np.cos(n)], -1)
    out = quat_mul(dq, q)
    return out / np.linalg.norm(out, axis=-1, keepdims=True)


def project_with_jac(model, params, Xc, want_jac=True):
    """Pixel projection and its derivatives wrt the camera-frame point (2x3)
    and wrt all params (2 x nparams), per observation.  ``params`` is [n, k]."""
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    iz = 1.0 / z
    u, v = x * iz, y * iz
    n = len(x)
    npar = params.shape[1]
    Jp = np.zeros((n, 2, 3)) if want_jac else None
    Jk = np.zeros((n, 2, npar)) if want_jac else None
    if model in (SIMPLE_PINHOLE, PINHOLE):
        if model == SIMPLE_PINHOLE:
            fx = fy = params[:, 0]; cx, cy = params[:, 1], params[:, 2]
        else:
            fx, fy, cx, cy = params[:, 0], params[:, 1], params[:, 2], params[:, 3]
        px = np.stack([fx * u + cx, fy * v + cy], 1)
        if want_jac:
            Jp[:, 0, 0] = fx * iz; Jp[:, 0, 2] = -fx * u * iz
            Jp[:, 1, 1] = fy * iz; Jp[:, 1, 2] = -fy * v * iz
            if model == SIMPLE_PINHOLE:
                Jk[:, 0, 0] = u; Jk[:, 1, 0] = v; Jk[:, 0, 1] = 1; Jk[:, 1, 2] = 1
            else:
                Jk[:, 0, 0] = u; Jk[:, 1, 1] = v; Jk[:, 0, 2] = 1; Jk[:, 1, 3] = 1
        return px, Jp, Jk
    f, cx, cy = params[:, 0], params[:, 1], params[:, 2]
    r2 = u * u + v * v
    if model == SIMPLE_RADIAL:
        k1 = params[:, 3]; k2 = np.zeros(n)
    elif model == RADIAL:
        k1, k2 = params[:, 3], params[:, 4]
    else:
        raise ValueError(model)
    d = 1 + k1 * r2 + k2 * r2 * r2
    dd = k1 + 2 * k2 * r2            # d(d)/d(r2)
    px = np.stack([f * u * d + cx, f * v * d + cy], 1)
    if want_jac:
        # d(ud, vd)/d(u, v)
        a00 = d + 2 * u * u * dd; a01 = 2 * u * v * dd; a11 = d + 2 * v * v * dd
        # d(u,v)/dXc = [[iz, 0, -u iz], [0, iz, -v iz]]
        Jp[:, 0, 0] = f * a00 * iz; Jp[:, 0, 1] = f * a01 * iz; Jp[:, 0, 2] = -f * iz * (a00 * u + a01 * v)
        Jp[:, 1, 0] = f * a01 * iz; Jp[:, 1, 1] = f * a11 * iz; Jp[:, 1, 2] = -f * iz * (a01 * u + a11 * v)
        Jk[:, 0, 0] = u * d; Jk[:, 1, 0] = v * d; Jk[:, 0, 1] = 1; Jk[:, 1, 2] = 1
        Jk[:, 0, 3] = f * u * r2; Jk[:, 1, 3] = f * v * r2
        if model == RADIAL:
            Jk[:, 0, 4] = f * u * r2 * r2; Jk[:, 1, 4] = f * v * r2 * r2
    return px, Jp, Jk


def skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]; S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]; S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]; S[..., 2, 1] = v[..., 0]
    return S


class BAProblem:
    """Flat BA problem (the arrays of include/b200sfm.h: b200sfm_ba_solve)."""

    def __init__(self, quat, trans, points, pt_obs_begin, obs_cam, obs_xy, cam_intr, intr_model, intr_params,
                 opts: BAOptions, cam_const_mask=None, rig=None):
        """``rig`` (known, constant camera rigs; bundle_adjustment.cc:147-161,
        RigReprojErrorConstantRigCostFunctor): dict(obs_img [N], img_q [I,4], img_t [I,3], img_intr [I]) --
        then quat/trans/obs_cam refer to FRAMES (rig_from_world) and every observation's image carries a
        constant cam_from_rig transform and its own intrinsics block.
        With ``opts.optimize_rig_poses`` (bundle_adjustment.cc:162-180,297-308, RigReprojErrorCostFunctor) the rig dict also
        carries img_sensor [I] (sensor of each image, -1 = reference sensor / constant) and sensor_q [S,4], sensor_t [S,3]:
        the cam_from_rig of every non-reference sensor is then an unknown (quaternion manifold + translation) shared by
        all images of that sensor."""
        self.opts = opts
        self.C, self.P = len(quat), len(points)
        lens = np.diff(pt_obs_begin)
        keep_pt = lens >= opts.min_num_view_per_track
        pt_of_obs = np.repeat(np.arange(self.P), lens)
        keep = keep_pt[pt_of_obs]
        self.obs_pt = pt_of_obs[keep]
        self.obs_cam = np.asarray(obs_cam)[keep].astype(np.int64)
        self.obs_xy = np.asarray(obs_xy, dtype=np.float64)[keep]
        self.N = len(self.obs_pt)
        self.cam_intr = np.asarray(cam_intr).astype(np.int64)
        self.rig = None
        if rig is not None:
            oi = np.asarray(rig["obs_img"])[keep].astype(np.int64)
            self.rig = dict(obs_img=oi, R_cr=quat_rotmat(np.asarray(rig["img_q"], dtype=np.float64))[oi],
                            t_cr=np.asarray(rig["img_t"], dtype=np.float64)[oi],
                            obs_intr=np.asarray(rig["img_intr"]).astype(np.int64)[oi], obs_sensor=None)
            if opts.optimize_rig_poses and "img_sensor" in rig:
                self.rig["obs_sensor"] = np.asarray(rig["img_sensor"]).astype(np.int64)[oi]
        self.intr_model = np.asarray(intr_model).astype(np.int64)
        self.K = len(self.intr_model)
        self.x0 = dict(quat=np.array(quat, dtype=np.float64), trans=np.array(trans, dtype=np.float64),
                       points=np.array(points, dtype=np.float64), intr=np.array(intr_params, dtype=np.float64))
        self.S = 0
        if self.rig is not None and self.rig["obs_sensor"] is not None:
            self.x0["sq"] = np.array(rig["sensor_q"], dtype=np.float64)
            self.x0["st"] = np.array(rig["sensor_t"], dtype=np.float64)
            self.S = len(self.x0["sq"])
        # ---- variable layout (tangent space) --------------------------------
        # cam_const_mask bit0: rotation constant, bit1: translation constant.
        mask = np.zeros(self.C, dtype=np.int64) if cam_const_mask is None else np.asarray(cam_const_mask).astype(np.int64)
        cam_used = np.zeros(self.C, dtype=bool)
        cam_used[self.obs_cam] = True
        rot_var = cam_used & ((mask & 1) == 0) & opts.optimize_rotations
        trn_var = cam_used & ((mask & 2) == 0) & opts.optimize_translation
        col = 0
        self.rot_col = np.full(self.C, -1)
        self.trn_col = np.full(self.C, -1)
        for c in range(self.C):
            if rot_var[c]:
                self.rot_col[c] = col; col += 3
            if trn_var[c]:
                self.trn_col[c] = col; col += 3
        # unknown cam_from_rig blocks (group 1 as well, bundle_adjustment.cc:222-236)
        self.sq_col = np.full(self.S, -1)
        self.st_col = np.full(self.S, -1)
        if self.S:
            s_used = np.zeros(self.S, dtype=bool)
            s_used[self.rig["obs_sensor"][self.rig["obs_sensor"] >= 0]] = True
            for si in range(self.S):
                if s_used[si]:
                    self.sq_col[si] = col; col += 3
                    self.st_col[si] = col; col += 3
        self.intr_cols = []  # per intrinsics block: list of (param idx, col)
        used_intr = np.zeros(self.K, dtype=bool)
        used_intr[self.rig["obs_intr"] if self.rig is not None else self.cam_intr[self.obs_cam]] = True
        for k in range(self.K):
            ent = []
            # bundle_adjustment.cc:273-293: optimize_principal_point -> no manifold and no constant block is set, EVERY
            # parameter is variable (whatever optimize_intrinsics says); else optimize_intrinsics -> SubsetManifold that
            # holds the principal point; else the block is constant
            if (opts.optimize_intrinsics or opts.optimize_principal_point) and used_intr[k]:
                foc, pp, extra = MODEL_LAYOUT[int(self.intr_model[k])]
                idxs = list(foc) + list(extra) + (list(pp) if opts.optimize_principal_point else [])
                for i in sorted(idxs):
                    ent.append((i, col)); col += 1
            self.intr_cols.append(ent)
        pt_used = np.zeros(self.P, dtype=bool)
        pt_used[self.obs_pt] = True
        self.pt_col = np.full(self.P, -1)
        if opts.optimize_points:
            idx = np.nonzero(pt_used)[0]
            self.pt_col[idx] = col + 3 * np.arange(len(idx))
            col += 3 * len(idx)
        self.ncols = col

    # -- state helpers --------------------------------------------------------
    def plus(self, x, delta):
        out = {k: v.copy() for k, v in x.items()}
        rv = self.rot_col >= 0
        if rv.any():
            d = delta[self.rot_col[rv][:, None] + np.arange(3)]
            out["quat"][rv] = quat_plus(x["quat"][rv], d)
        tv = self.trn_col >= 0
        if tv.any():
            out["trans"][tv] += delta[self.trn_col[tv][:, None] + np.arange(3)]
        pv = self.pt_col >= 0
        if pv.any():
            out["points"][pv] += delta[self.pt_col[pv][:, None] + np.arange(3)]
        for k, ent in enumerate(self.intr_cols):
            for i, c in ent:
                out["intr"][k, i] += delta[c]
        sv = self.sq_col >= 0
        if sv.any():
            out["sq"][sv] = quat_plus(x["sq"][sv], delta[self.sq_col[sv][:, None] + np.arange(3)])
            out["st"][sv] += delta[self.st_col[sv][:, None] + np.arange(3)]
        return out

    def x_norm(self, x, y=None):
        """Ambient norm over the variable parameter blocks (or of x - y)."""
        tot = 0.0
        rv, tv, pv = self.rot_col >= 0, self.trn_col >= 0, self.pt_col >= 0
        blocks = [("quat", rv), ("trans", tv), ("points", pv)]
        if self.S:
            blocks += [("sq", self.sq_col >= 0), ("st", self.st_col >= 0)]
        for key, m in blocks:
            a = x[key][m] if y is None else x[key][m] - y[key][m]
            tot += float((a * a).sum())
        for k, ent in enumerate(self.intr_cols):
            if ent:
                npar = len(sum(MODEL_LAYOUT[int(self.intr_model[k])], ()))
                a = x["intr"][k, :npar] if y is None else x["intr"][k, :npar] - y["intr"][k, :npar]
                tot += float((a * a).sum())
        return np.sqrt(tot)

    # -- residuals / Jacobian -------------------------------------------------
    def residuals(self, x, want_jac):
        R = quat_rotmat(x["quat"])[self.obs_cam]
        X = x["points"][self.obs_pt]
        RX = np.einsum("nij,nj->ni", R, X)
        Xc = RX + x["trans"][self.obs_cam]
        R_cr = RY = None
        if self.rig is not None:      # X_c = R_cr (R_f X + t_f) + t_cr
            R_cr, t_cr = self.rig["R_cr"], self.rig["t_cr"]
            os_ = self.rig["obs_sensor"]
            if os_ is not None:       # unknown cam_from_rig of the non-reference sensors comes from the state
                m = os_ >= 0
                R_cr, t_cr = R_cr.copy(), t_cr.copy()
                R_cr[m] = quat_rotmat(x["sq"])[os_[m]]
                t_cr[m] = x["st"][os_[m]]
            RY = np.einsum("nij,nj->ni", R_cr, Xc)
            Xc = RY + t_cr
        valid = Xc[:, 2] > Z_EPS
        Xs = np.where(valid[:, None], Xc, np.array([0.0, 0.0, 1.0]))
        res = np.zeros((self.N, 2))
        Jp = np.zeros((self.N, 2, 3)) if want_jac else None
        Jk_all = {}
        ci = self.rig["obs_intr"] if self.rig is not None else self.cam_intr[self.obs_cam]
        for k in range(self.K):
            mk = ci == k
            if not mk.any():
                continue
            model = int(self.intr_model[k])
            npar = len(sum(MODEL_LAYOUT[model], ()))
            par = np.broadcast_to(x["intr"][k, :npar], (int(mk.sum()), npar))
            px, jp, jk = project_with_jac(model, par, Xs[mk], want_jac)
            res[mk] = px - self.obs_xy[mk]
            if want_jac:
                Jp[mk] = jp
                Jk_all[k] = (mk, jk)
        res[~valid] = 0.0
        if not want_jac:
            return res, None
        Jp[~valid] = 0.0
        Jsq = Jst = None
        if self.rig is not None:
            if self.S:                # d X_c / d(sensor rotation) = -2 [R_cr Y]x (left perturbation), d X_c / d t_cr = I
                Jsq = np.einsum("nij,njk->nik", Jp, -2.0 * skew(RY))
                Jst = Jp
            Jp = np.einsum("nij,njk->nik", Jp, R_cr)      # chain the frame / point blocks through R_cr
        # d Xc / d delta (quaternion manifold, left perturbation angle 2|d|): -2 [R X]x
        Jrot = np.einsum("nij,njk->nik", Jp, -2.0 * skew(RX))
        Jtrn = Jp
        Jpt = np.einsum("nij,njk->nik", Jp, R)
        return res, (Jrot, Jtrn, Jpt, Jk_all, valid, Jsq, Jst)

    def evaluate(self, x, want_jac):
        res, jac = self.residuals(x, want_jac)
        s = (res * res).sum(1)
        rho0, rho1 = huber_rho(s, self.opts.thres_loss_function)
        cost = 0.5 * float(rho0.sum())
        w = np.sqrt(rho1)
        r = (res * w[:, None]).ravel()
        if not want_jac:
            return cost, r, None
        Jrot, Jtrn, Jpt, Jk_all, valid, Jsq, Jst = jac
        rows_l, cols_l, vals_l = [], [], []
        row0 = 2 * np.arange(self.N)

        def add_block(Jb, col_start, sel):
            """Jb [n,2,3] blocks for observations `sel` at columns col_start+0..2."""
            idx = np.nonzero(sel)[0]
            if len(idx) == 0:
                return
            rr = (row0[idx][:, None, None] + np.arange(2)[None, :, None] + np.zeros((1, 1, 3), dtype=np.int64))
            cc = (col_start[idx][:, None, None] + np.zeros((1, 2, 1), dtype=np.int64) + np.arange(3)[None, None, :])
            rows_l.append(rr.ravel()); cols_l.append(cc.ravel())
            vals_l.append((Jb[idx] * w[idx][:, None, None]).ravel())

        rc = self.rot_col[self.obs_cam]
        add_block(Jrot, rc, rc >= 0)
        tc = self.trn_col[self.obs_cam]
        add_block(Jtrn, tc, tc >= 0)
        pc = self.pt_col[self.obs_pt]
        add_block(Jpt, pc, pc >= 0)
        if Jsq is not None:
            os_ = self.rig["obs_sensor"]
            qc = np.where(os_ >= 0, self.sq_col[np.maximum(os_, 0)], -1)
            tc2 = np.where(os_ >= 0, self.st_col[np.maximum(os_, 0)], -1)
            add_block(Jsq, qc, qc >= 0)
            add_block(Jst, tc2, tc2 >= 0)
        for k, (mk, jk) in Jk_all.items():
            idx = np.nonzero(mk)[0]
            vmask = valid[idx]
            for i, c in self.intr_cols[k]:
                for a in range(2):
                    rows_l.append(row0[idx] + a)
                    cols_l.append(np.full(len(idx), c))
                    vals_l.append(jk[:, a, i] * w[idx] * vmask)
        if rows_l:
            J = sp.csr_matrix((np.concatenate(vals_l), (np.concatenate(rows_l), np.concatenate(cols_l))),
                              shape=(2 * self.N, self.ncols))
        else:
            J = sp.csr_matrix((2 * self.N, self.ncols))
        return cost, r, J


def solve_ba(quat, trans, points, pt_obs_begin, obs_cam, obs_xy, cam_intr, intr_model, intr_params,
             opts: BAOptions | None = None, cam_const_mask=None, verbose=False, rig=None):
    """Oracle counterpart of BundleAdjuster::Solve (bundle_adjustment.cc:11-106).
    ``cam_const_mask`` [C] uint8: bit0 rotation constant, bit1 translation
    constant (the caller marks the first frame with 3, .cc:261-266).
    Returns (state dict, LMSummary)."""
    opts = opts or BAOptions()
    prob = BAProblem(quat, trans, points, pt_obs_begin, obs_cam, obs_xy, cam_intr, intr_model, intr_params, opts,
                     cam_const_mask, rig)
    lm = LMOptions(max_num_iterations=opts.max_num_iterations, function_tolerance=opts.function_tolerance,
                   verbose=verbose)
    if prob.N == 0 or prob.ncols == 0:
        s = LMSummary(termination="empty problem")
        return prob.x0, s
    x, summ = solve_lm(prob.x0, prob.evaluate, prob.plus, lm, x_norm_fn=prob.x_norm)
    return x, summ
