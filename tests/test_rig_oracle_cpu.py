"""CPU tests of the known-rig oracle extensions and host-side rig preparation (no GPU):
 * the rig BA oracle (oracle/ba_oracle.py, rig=) is the same function of the state as the trivial-frame oracle
   over the F*S images with composed poses, and its analytic Jacobian matches finite differences
   (bundle_adjustment.cc:147-161, RigReprojErrorConstantRigCostFunctor);
 * rig GP terms (global_positioning.cc:325-346) and the frame view graph for rotation averaging
   (global_rotation_averaging.cc:274-309) agree with their per-image definitions."""
import numpy as np

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ba_oracle as B, gp_oracle as GPO, ra_oracle as RO


def _rig_problem(rs, st, opts):
    return B.BAProblem(st.quat, st.trans, st.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy, np.zeros(rs.F, np.int32),
                       rs.intr_model, st.intr_params, opts, rig=rs.rig_dict())


def test_rig_scene_is_consistent_and_zero_cost_at_ground_truth():
    rs = S.make_rig_scene(8, 3, 200, seed=4, model=S.RADIAL)
    assert rs.N == rs.pt_obs_begin[-1] and (np.diff(rs.pt_obs_begin) >= 3).all()
    assert rs.obs_sensor.max() < rs.S and rs.obs_frame.max() < rs.F
    prob = _rig_problem(rs, rs, B.BAOptions())
    assert prob.evaluate(prob.x0, False)[0] < 1e-18


def test_rig_oracle_equals_image_oracle_and_fd_jacobian():
    rs = S.make_rig_scene(10, 3, 300, seed=4, model=S.SIMPLE_RADIAL)
    st = S.perturb_rig_scene(rs)
    prob = _rig_problem(rs, st, B.BAOptions())
    im = st.images_scene()
    pim = B.BAProblem(im.quat, im.trans, im.points, im.pt_obs_begin, im.obs_cam, im.obs_xy, im.cam_intr, im.intr_model,
                      im.intr_params, B.BAOptions())
    c_r, c_i = prob.evaluate(prob.x0, False)[0], pim.evaluate(pim.x0, False)[0]
    assert abs(c_r - c_i) <= 1e-12 * c_i
    for oi in (False, True):   # the loss is switched off so that the residual is differentiable everywhere
        p = _rig_problem(rs, st, B.BAOptions(thres_loss_function=1e9, optimize_intrinsics=oi))
        _, _, J = p.evaluate(p.x0, True)
        d = np.random.default_rng(0).normal(size=J.shape[1]) * 1e-6
        r1 = p.evaluate(p.plus(p.x0, d), False)[1]
        r0 = p.evaluate(p.plus(p.x0, -d), False)[1]
        assert np.abs((r1 - r0) / 2 - J @ d).max() < 1e-9 * max(1.0, np.abs(J @ d).max() / 1e-3)
    assert J.shape[1] == 6 * (rs.F - 0) + 3 * rs.P + 2 * rs.S   # SIMPLE_RADIAL: f and k per sensor block


def test_rig_ba_oracle_recovers_ground_truth():
    rs = S.make_rig_scene(8, 3, 250, seed=6)
    st = S.perturb_rig_scene(rs)
    mask = E.first_frame_mask(rs.F)
    x, summ = B.solve_ba(st.quat, st.trans, st.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy, np.zeros(rs.F, np.int32),
                         rs.intr_model, st.intr_params, B.BAOptions(), mask, rig=rs.rig_dict())
    assert summ.final_cost < 1e-10 * summ.initial_cost
    rot, cen = G.compare_reconstructions(G.quat_xyzw_to_rotmat(x["quat"]), x["trans"], G.quat_xyzw_to_rotmat(rs.quat),
                                         rs.trans)[:2]
    assert rot < 1e-2 and cen < 1e-4


def test_rig_world_terms_match_the_image_definitions():
    rs = S.make_rig_scene(6, 4, 120, seed=7)
    im = rs.images_scene()
    bear = S.bearings_from_scene(im)
    t_obs, t_rig = E.rig_world_terms(rs.quat, rs.sensor_quat, rs.sensor_trans, bear, rs.obs_frame, rs.obs_sensor)
    assert np.abs(t_obs - E.world_bearings(im.quat, bear, im.obs_cam)).max() < 1e-13
    # with exact bearings the residual t_obs - s (X - c_frame + t_rig) vanishes for s = 1 / |X - c_image|
    c_f = G.centers_from_pose(G.quat_xyzw_to_rotmat(rs.quat), rs.trans)
    pt = np.repeat(np.arange(rs.P), np.diff(rs.pt_obs_begin))
    d = rs.points[pt] - c_f[rs.obs_frame] + t_rig
    c_i = G.centers_from_pose(G.quat_xyzw_to_rotmat(im.quat), im.trans)
    assert np.abs(d - (rs.points[pt] - c_i[im.obs_cam])).max() < 1e-12
    s = 1.0 / np.linalg.norm(d, axis=1)
    assert np.abs(t_obs - s[:, None] * d).max() < 1e-9
    prob = GPO.GPProblem(c_f, rs.points, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(), s, obs_offset=t_rig)
    assert prob.evaluate(prob.x0, False)[0] < 1e-16


def test_rig_view_graph_folds_image_pairs_onto_frames():
    rs = S.make_rig_scene(12, 3, 10, seed=11)
    Ri, _ = rs.image_poses()
    rng = np.random.default_rng(5)
    n_img = rs.F * rs.S
    ei, ej = np.triu_indices(n_img, 1)
    sel = rng.uniform(size=len(ei)) < 0.3
    ei, ej = ei[sel].astype(np.int32), ej[sel].astype(np.int32)
    vg_img = S.ViewGraph(n_img, ei, ej, Ri[ej] @ np.swapaxes(Ri[ei], -1, -2), np.ones(len(ei)), Ri)
    img_frame, img_sensor = np.repeat(np.arange(rs.F), rs.S), np.tile(np.arange(rs.S), rs.F)
    Rf = G.quat_xyzw_to_rotmat(rs.quat)
    vg = E.rig_view_graph(vg_img, img_frame, img_sensor, rs.sensor_quat, Rf)
    same = img_frame[ei] == img_frame[ej]
    assert vg.E == int((~same).sum()) and (vg.ei != vg.ej).all()
    # exact relative rotations of the frames
    assert np.abs(vg.R_rel - Rf[vg.ej] @ np.swapaxes(Rf[vg.ei], -1, -2)).max() < 1e-12
    R0 = E.initialize_from_maximum_spanning_tree(vg, None)
    th, info = RO.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
    R = G.so3_exp(th)
    # (matrix comparison: the arccos-based angle helper has a ~1e-6 deg floor)
    assert np.abs(R @ np.swapaxes(R[:1], -1, -2) - Rf @ np.swapaxes(Rf[:1], -1, -2)).max() < 1e-9


def test_optimised_rig_poses_oracle():
    """optimize_rig_poses (bundle_adjustment.cc:162-180,297-308, RigReprojErrorCostFunctor): the cam_from_rig of the
    non-reference sensors are unknown blocks shared by all frames.  Oracle only so far (no device path yet): analytic
    Jacobian against finite differences and recovery of perturbed rig extrinsics on noise-free data."""
    rs = S.make_rig_scene(10, 3, 300, seed=4, model=S.SIMPLE_RADIAL)
    st = S.perturb_rig_scene(rs)
    rig = st.rig_dict()
    rng = np.random.default_rng(1)
    rig["sensor_q"][1:] = G.rotmat_to_quat_xyzw_fast(G.so3_exp(rng.normal(size=(2, 3)) * 0.01) @ G.quat_xyzw_to_rotmat(rig["sensor_q"][1:]))
    rig["sensor_t"][1:] += rng.normal(size=(2, 3)) * 0.02
    args = (st.quat, st.trans, st.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy, np.zeros(rs.F, np.int32), rs.intr_model,
            st.intr_params)
    p = B.BAProblem(*args, B.BAOptions(thres_loss_function=1e9, optimize_rig_poses=True), rig=rig)
    _, _, J = p.evaluate(p.x0, True)
    assert J.shape[1] == 6 * rs.F + 3 * rs.P + 6 * (rs.S - 1)          # the reference sensor stays the identity
    d = rng.normal(size=J.shape[1]) * 1e-6
    r1 = p.evaluate(p.plus(p.x0, d), False)[1]
    r0 = p.evaluate(p.plus(p.x0, -d), False)[1]
    assert np.abs((r1 - r0) / 2 - J @ d).max() < 1e-10
    x, summ = B.solve_ba(*args, B.BAOptions(optimize_rig_poses=True), E.first_frame_mask(rs.F), rig=rig)
    assert summ.final_cost < 1e-12 * summ.initial_cost
    assert np.abs(G.quat_xyzw_to_rotmat(x["sq"]) - G.quat_xyzw_to_rotmat(rs.sensor_quat)).max() < 1e-6
    # nothing metric is held fixed any more: the rig baselines are recovered up to the global scale
    s = np.linalg.norm(x["st"][1]) / np.linalg.norm(rs.sensor_trans[1])
    assert np.abs(x["st"][1:] - s * rs.sensor_trans[1:]).max() < 1e-6 and abs(s - 1) < 0.05
    # without the flag the same dictionary is the constant-rig problem
    q = B.BAProblem(*args, B.BAOptions(), rig=rig)
    assert q.S == 0 and "sq" not in q.x0


def test_rig_unknown_bata_oracle():
    """RigUnknownBATAPairwiseDirectionError (cost_function.h:90-134, global_positioning.cc:347-364): the camera centre in
    the rig frame of a sensor without a known cam_from_rig is an unknown shared by its images.  Oracle only so far.
    With u_s = -R_cr^T t_cr the term -R_rw^T u_s equals the known-rig offset R_cw^T t_cr, so the residual vanishes at
    the ground truth; Jacobian against finite differences; the solve drives the cost to zero."""
    rs = S.make_rig_scene(10, 3, 250, seed=12)
    bear = S.bearings_from_scene(rs.images_scene())
    t_obs, t_rig = E.rig_world_terms(rs.quat, rs.sensor_quat, rs.sensor_trans, bear, rs.obs_frame, rs.obs_sensor)
    Rf = G.quat_xyzw_to_rotmat(rs.quat)
    c_f = G.centers_from_pose(Rf, rs.trans)
    u = -np.einsum("sji,sj->si", G.quat_xyzw_to_rotmat(rs.sensor_quat), rs.sensor_trans)
    ru = dict(obs_sensor=np.where(rs.obs_sensor > 0, rs.obs_sensor.astype(np.int64), -1), R_rw=Rf[rs.obs_frame], centers=u)
    pt = np.repeat(np.arange(rs.P), np.diff(rs.pt_obs_begin))
    s = 1.0 / np.linalg.norm(rs.points[pt] - c_f[rs.obs_frame] + t_rig, axis=1)
    p = GPO.GPProblem(c_f, rs.points, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(), s, rig_unknown=ru)
    assert p.evaluate(p.x0, False)[0] < 1e-20
    rng = np.random.default_rng(0)
    ru2 = dict(ru, centers=u + rng.normal(size=u.shape) * 0.1)
    c0, p0 = c_f + rng.normal(size=c_f.shape) * 0.3, rs.points + rng.normal(size=rs.points.shape) * 0.3
    q = GPO.GPProblem(c0, p0, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(thres_loss_function=1e9), s, rig_unknown=ru2)
    _, _, J = q.evaluate(q.x0, True)
    assert J.shape[1] == 3 * rs.F + 3 * rs.P + (q.N - 1) + 3 * (rs.S - 1)
    d = rng.normal(size=J.shape[1]) * 1e-6
    r1, r0 = q.evaluate(q.plus(q.x0, d), False)[1], q.evaluate(q.plus(q.x0, -d), False)[1]
    assert np.abs((r1 - r0) / 2 - J @ d).max() < 1e-12
    x, summ = GPO.solve_gp(c0, p0, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(), None, rig_unknown=ru2)
    assert summ.final_cost < 1e-12 * summ.initial_cost
