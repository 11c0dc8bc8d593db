"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of
glomap/processors/track_filter.cc (pixel-space reprojection filter :7-52,
angle filter :54-90, triangulation-angle filter :92-127) on the flat arrays.
PARITY UNPINNED (no golden vectors in the reference; COLMAP's ImgFromCam is
restated for the four supported models in glomap_b200.synthetic.project)."""
import numpy as np

from .ba_oracle import quat_rotmat

EPS = 1e-12


def _cam_points(quat, trans, points, pt_obs_begin, obs_cam):
    pt = np.repeat(np.arange(len(points)), np.diff(pt_obs_begin))
    R = quat_rotmat(np.asarray(quat, dtype=np.float64))[obs_cam]
    return np.einsum("nij,nj->ni", R, points[pt]) + trans[obs_cam], pt


def filter_reprojection(scene, max_err, project):
    Xc, pt = _cam_points(scene.quat, scene.trans, scene.points, scene.pt_obs_begin, scene.obs_cam)
    keep = np.zeros(scene.N, bool)
    ok = ~(Xc[:, 2] < EPS)
    ci = scene.cam_intr[scene.obs_cam]
    for k in range(len(scene.intr_model)):
        m = ok & (ci == k)
        if m.any():
            px = project(int(scene.intr_model[k]), scene.intr_params[k], Xc[m])
            keep[m] = np.linalg.norm(px - scene.obs_xy[m], axis=1) < max_err
    changed = np.zeros(scene.P, bool)
    np.logical_or.at(changed, pt, ~keep)
    return keep, int(changed.sum())


def filter_reprojection_normalized(scene, bearings, max_err):
    """in_normalized_image = true (track_filter.cc:24-31): |X_c.xy / X_c.z - b.xy / (b.z + EPS)| < max_err."""
    Xc, pt = _cam_points(scene.quat, scene.trans, scene.points, scene.pt_obs_begin, scene.obs_cam)
    ok = ~(Xc[:, 2] < EPS)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = Xc[:, :2] / Xc[:, 2:3] - bearings[:, :2] / (bearings[:, 2:3] + EPS)
    keep = ok & (np.linalg.norm(d, axis=1) < max_err)
    changed = np.zeros(scene.P, bool)
    np.logical_or.at(changed, pt, ~keep)
    return keep, int(changed.sum())


def filter_angle(scene, bearings, max_angle_deg, calibrated=None):
    Xc, pt = _cam_points(scene.quat, scene.trans, scene.points, scene.pt_obs_begin, scene.obs_cam)
    thres, thres_u = np.cos(np.radians(max_angle_deg)), np.cos(np.radians(2 * max_angle_deg))
    ok = ~(Xc[:, 2] < EPS)
    d = (Xc / np.linalg.norm(Xc, axis=1, keepdims=True) * bearings).sum(1)
    th = thres if calibrated is None else np.where(np.asarray(calibrated).astype(bool)[scene.obs_cam], thres, thres_u)
    keep = ok & (d > th)
    changed = np.zeros(scene.P, bool)
    np.logical_or.at(changed, pt, ~keep)
    return keep, int(changed.sum())


def filter_triangulation_angle(scene, min_angle_deg):
    R = quat_rotmat(np.asarray(scene.quat, dtype=np.float64))
    c = -np.einsum("nji,nj->ni", R, scene.trans)
    thres = np.cos(np.radians(min_angle_deg))
    keep = np.zeros(scene.P, bool)
    for p in range(scene.P):
        cams = scene.obs_cam[scene.pt_obs_begin[p]:scene.pt_obs_begin[p + 1]]
        rays = scene.points[p] - c[cams]
        rays /= np.linalg.norm(rays, axis=1, keepdims=True)
        G = rays @ rays.T
        iu = np.triu_indices(len(cams), 1)
        keep[p] = bool((G[iu] < thres).any())
    return keep, int((~keep).sum())
