"""GPU parity tests for KNOWN camera rigs (non-trivial frames with constant cam_from_rig):
BundleAdjuster (bundle_adjustment.cc:147-161, RigReprojErrorConstantRigCostFunctor),
GlobalPositioner (RigBATAPairwiseDirectionError with constant rig scale,
global_positioning.cc:325-346,493-497), RotationEstimator over frames
(global_rotation_averaging.cc:274-309) and the track filters -- through the C ABI against the
CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ba_oracle as B, filter_oracle as FO, gp_oracle as GPO, ra_oracle as RO

pytestmark = pytest.mark.gpu


def _rig_oracle(rs, init, opts, mask):
    return B.solve_ba(init.quat, init.trans, init.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy,
                      np.zeros(rs.F, np.int32), rs.intr_model, init.intr_params, opts, mask, rig=rs.rig_dict())


def _device(init, mask, design=0, tol=1e-12, **kw):
    opts = E.BundleAdjusterOptions(**kw)
    opts.design = design
    opts.solver_options.pcg_rel_tolerance = tol
    opts.solver_options.pcg_max_iterations = 3000
    ba = E.BundleAdjuster(opts)
    dev = init.copy()
    ok = ba.Solve(dev, mask)
    return ok, dev, ba.summary


@pytest.mark.parametrize("design", [1, 2])
def test_rig_ba_tracks_oracle(design):
    rs = S.make_rig_scene(12, 3, 500, seed=5, pixel_sigma=0.5)
    init = S.perturb_rig_scene(rs)
    mask = E.first_frame_mask(rs.F)
    ok, dev, st = _device(init, mask, design=design, optimize_intrinsics=False)
    x, summ = _rig_oracle(rs, init, B.BAOptions(), mask)
    assert ok and st.usable
    assert st.iterations == summ.iterations
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert abs(st.final_cost - summ.final_cost) <= 1e-8 * summ.final_cost
    assert np.abs(dev.quat - x["quat"]).max() < 1e-6
    assert np.abs(dev.trans - x["trans"]).max() < 1e-6
    assert np.abs(dev.points - x["points"]).max() < 1e-6


def test_rig_ba_noise_free_recovers_ground_truth():
    rs = S.make_rig_scene(10, 4, 500, seed=6, model=S.SIMPLE_RADIAL)
    init = S.perturb_rig_scene(rs)
    ok, dev, st = _device(init, None, tol=1e-10, optimize_intrinsics=False)
    assert ok and st.final_cost < 1e-8 * st.initial_cost
    rot, cen = G.compare_reconstructions(G.quat_xyzw_to_rotmat(dev.quat), dev.trans, G.quat_xyzw_to_rotmat(rs.quat),
                                         rs.trans)[:2]
    assert rot < 1e-2 and cen < 1e-4, (rot, cen)


@pytest.mark.parametrize("model", [S.SIMPLE_PINHOLE, S.SIMPLE_RADIAL])
def test_rig_ba_with_intrinsics_matches_oracle(model):
    """optimize_intrinsics (the reference default): one intrinsics block per SENSOR, shared by the
    images of that sensor in all frames -- the dense border of the reduced system."""
    rs = S.make_rig_scene(12, 3, 600, seed=8, pixel_sigma=0.3, model=model)
    init = S.perturb_rig_scene(rs)
    init.intr_params = init.intr_params.copy()
    init.intr_params[:, 0] *= 1.01
    mask = E.first_frame_mask(rs.F)
    ok, dev, st = _device(init, mask, optimize_intrinsics=True)
    x, summ = _rig_oracle(rs, init, B.BAOptions(optimize_intrinsics=True), mask)
    assert ok
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert abs(st.final_cost - summ.final_cost) <= 1e-6 * summ.final_cost
    npar = S.MODEL_NUM_PARAMS[model]
    assert np.abs(dev.intr_params[:, :npar] - x["intr"][:, :npar]).max() < 1e-4 * 1000
    assert np.abs(dev.quat - x["quat"]).max() < 1e-5


def test_rig_cost_and_filters_equal_the_equivalent_image_problem():
    """A rig problem and the same observations posed over the F*S images with composed poses are the
    same functions of the state: cost and all three track filters must agree."""
    rs = S.make_rig_scene(10, 3, 400, seed=9, pixel_sigma=1.0)
    st_ = S.perturb_rig_scene(rs, rot_deg=0.3)
    im = st_.images_scene()
    ctx = E.default_context()
    opts = E.BundleAdjusterOptions(optimize_intrinsics=False)
    pr = E.BAProblem(ctx, st_, 3)
    pi = E.BAProblem(ctx, im, 3)
    pr.set_state(st_.intr_params, st_.quat, st_.trans, st_.points)
    pi.set_state(im.intr_params, im.quat, im.trans, im.points)
    c_r, c_i = pr.cost(opts), pi.cost(opts)
    assert abs(c_r - c_i) <= 1e-11 * c_i
    k_r, n_r = pr.filter_reprojection(20.0)
    k_i, n_i = pi.filter_reprojection(20.0)
    k_o, n_o = FO.filter_reprojection(im, 20.0, S.project)
    assert np.array_equal(k_r, k_i) and n_r == n_i and np.array_equal(k_r, k_o) and n_r == n_o
    assert 0 < n_r < rs.P
    bear = S.bearings_from_scene(im)
    k_r, n_r = pr.filter_angle(bear, 1.0)
    k_i, n_i = pi.filter_angle(bear, 1.0)
    assert np.array_equal(k_r, k_i) and n_r == n_i and 0 < n_r < rs.P
    t_r, m_r = pr.filter_triangulation_angle(100.0)
    t_i, m_i = pi.filter_triangulation_angle(100.0)
    t_o, m_o = FO.filter_triangulation_angle(im, 100.0)
    assert np.array_equal(t_r, t_i) and m_r == m_i and np.array_equal(t_r, t_o) and m_r > 0
    pr.free(); pi.free()


def test_rig_global_positioning_matches_oracle():
    rs = S.make_rig_scene(14, 3, 500, seed=10, pixel_sigma=0.3)
    im = rs.images_scene()
    bear = S.bearings_from_scene(im)
    rng = np.random.default_rng(3)
    cen0 = G.centers_from_pose(G.quat_xyzw_to_rotmat(rs.quat), rs.trans) + rng.normal(size=(rs.F, 3)) * 0.5
    pts0 = rs.points + rng.normal(size=rs.points.shape) * 0.5
    opts = E.GlobalPositionerOptions(generate_random_positions=False, generate_random_points=False, generate_scales=True)
    opts.solver_options.pcg_rel_tolerance = 1e-12
    opts.solver_options.pcg_max_iterations = 3000
    prob = E.PositioningProblem(rs.quat, rs.pt_obs_begin, rs.obs_frame, bear, centers=cen0.copy(), points=pts0.copy(),
                                obs_sensor=rs.obs_sensor, sensor_quat=rs.sensor_quat, sensor_trans=rs.sensor_trans,
                                sensor_calibrated=np.array([1, 0, 1], np.uint8))
    gp = E.GlobalPositioner(opts)
    assert gp.Solve(prob)
    t_obs, t_rig = E.rig_world_terms(rs.quat, rs.sensor_quat, rs.sensor_trans, bear, rs.obs_frame, rs.obs_sensor)
    # the oracle takes the loss scale per "camera": pose the calibrated flag per image (F*S pseudo cameras)
    ocal = np.array([1, 0, 1], bool)[rs.obs_sensor]
    po = GPO.GPProblem(cen0, pts0, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(), None, obs_offset=t_rig)
    po.loss_scale = np.where(ocal[po.keep], 1.0, 0.5)
    from oracle.ceres_lm import LMOptions, solve_lm
    x, summ = solve_lm(po.x0, po.evaluate, po.plus, LMOptions(max_num_iterations=100, function_tolerance=1e-5),
                       project=po.project, x_norm_fn=po.x_norm)
    st = gp.summary
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert abs(st.final_cost - summ.final_cost) <= 1e-6 * max(summ.final_cost, 1e-12)
    assert np.abs(prob.centers - x["centers"]).max() < 1e-5
    assert np.abs(prob.points - x["points"]).max() < 1e-5


def test_rig_rotation_averaging_over_frames():
    """Image-pair rotations folded onto frames (R_c2r2^T R_21 R_c1r1), same-frame pairs skipped."""
    rs = S.make_rig_scene(16, 3, 10, seed=11)
    Ri, _ = rs.image_poses()
    rng = np.random.default_rng(5)
    n_img = rs.F * rs.S
    ei, ej = np.triu_indices(n_img, 1)
    sel = rng.uniform(size=len(ei)) < 0.25
    ei, ej = ei[sel].astype(np.int32), ej[sel].astype(np.int32)
    noise = G.so3_exp(rng.normal(size=(len(ei), 3)) * np.radians(0.5))
    R_rel = noise @ Ri[ej] @ np.swapaxes(Ri[ei], -1, -2)
    vg_img = S.ViewGraph(n_img, ei, ej, R_rel, np.ones(len(ei)), Ri)
    img_frame = np.repeat(np.arange(rs.F), rs.S)
    img_sensor = np.tile(np.arange(rs.S), rs.F)
    Rf = G.quat_xyzw_to_rotmat(rs.quat)
    vg = E.rig_view_graph(vg_img, img_frame, img_sensor, rs.sensor_quat, Rf)
    assert vg.E < vg_img.E and (vg.ei != vg.ej).all()
    est = E.RotationEstimator(E.RotationEstimatorOptions())
    ok, R = est.EstimateRotations(vg)
    assert ok
    R0 = E.initialize_from_maximum_spanning_tree(vg, None)
    th, info = RO.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
    assert np.abs(R - G.so3_exp(th)).max() < 1e-7
    err = G.rotation_angle_deg(R @ np.swapaxes(R[:1], -1, -2), Rf @ np.swapaxes(Rf[:1], -1, -2))
    assert err.max() < 1.0


# ---- unknown cam_from_rig: optimize_rig_poses (bundle_adjustment.cc:162-180,296-308, RigReprojErrorCostFunctor) ----------
def _perturbed_rig(rs, seed=1):
    init = S.perturb_rig_scene(rs)
    rng = np.random.default_rng(seed)
    init.sensor_quat = init.sensor_quat.copy(); init.sensor_trans = init.sensor_trans.copy()
    init.sensor_quat[1:] = G.rotmat_to_quat_xyzw_fast(G.so3_exp(rng.normal(size=(rs.S - 1, 3)) * 0.01) @
                                                      G.quat_xyzw_to_rotmat(init.sensor_quat[1:]))
    init.sensor_trans[1:] += rng.normal(size=(rs.S - 1, 3)) * 0.02
    return init


@pytest.mark.parametrize("with_intr", [False, True])
def test_optimised_rig_poses_track_oracle(with_intr):
    """The shape of the reference's own NonTrivialUnknownRig test (global_mapper_test.cc:128): 3 cameras per rig, the
    cam_from_rig of the two non-reference sensors unknown and shared by all frames; device vs oracle iteration by
    iteration, with and without the (shared, per-sensor) intrinsics refined in the same solve."""
    rs = S.make_rig_scene(12, 3, 600, seed=14, pixel_sigma=0.3, model=S.SIMPLE_RADIAL)
    init = _perturbed_rig(rs)
    mask = E.first_frame_mask(rs.F)
    ok, dev, st = _device(init, mask, optimize_intrinsics=with_intr, optimize_rig_poses=True)
    rig = init.rig_dict()
    x, summ = B.solve_ba(init.quat, init.trans, init.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy, np.zeros(rs.F, np.int32),
                         rs.intr_model, init.intr_params, B.BAOptions(optimize_intrinsics=with_intr, optimize_rig_poses=True), mask,
                         rig=rig)
    assert ok and st.usable
    assert st.iterations == summ.iterations, (st.iterations, summ.iterations)
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert abs(st.final_cost - summ.final_cost) <= 1e-7 * summ.final_cost
    assert np.abs(G.quat_xyzw_to_rotmat(dev.sensor_quat) - G.quat_xyzw_to_rotmat(x["sq"])).max() < 1e-6
    assert np.abs(dev.sensor_trans - x["st"]).max() < 1e-6
    assert np.array_equal(dev.sensor_quat[0], init.sensor_quat[0]) and np.array_equal(dev.sensor_trans[0], init.sensor_trans[0])
    assert np.abs(dev.quat - x["quat"]).max() < 1e-6 and np.abs(dev.points - x["points"]).max() < 1e-5


def test_optimised_rig_poses_recover_the_extrinsics_noise_free():
    rs = S.make_rig_scene(10, 3, 500, seed=4, model=S.SIMPLE_RADIAL)
    init = _perturbed_rig(rs)
    ok, dev, st = _device(init, E.first_frame_mask(rs.F), tol=1e-10, optimize_intrinsics=False, optimize_rig_poses=True)
    assert ok and st.final_cost < 1e-10 * st.initial_cost
    assert np.abs(G.quat_xyzw_to_rotmat(dev.sensor_quat) - G.quat_xyzw_to_rotmat(rs.sensor_quat)).max() < 1e-5
    s = np.linalg.norm(dev.sensor_trans[1]) / np.linalg.norm(rs.sensor_trans[1])      # nothing metric is fixed: up to scale
    assert np.abs(dev.sensor_trans[1:] - s * rs.sensor_trans[1:]).max() < 1e-5 and abs(s - 1) < 0.05


def test_rig_unknown_bata_matches_oracle():
    """RigUnknownBATAPairwiseDirectionError (cost_function.h:90-136, global_positioning.cc:347-364): sensor 0 is the
    reference sensor, sensor 1 is calibrated (RigBATA offset), the camera centre of sensor 2 in the rig frame is an
    unknown shared by all its images.  Same start as the oracle, tight PCG: same initial / final cost, same centres."""
    rs = S.make_rig_scene(14, 3, 500, seed=11, pixel_sigma=0.3)
    bear = S.bearings_from_scene(rs.images_scene())
    rng = np.random.default_rng(4)
    Rf = G.quat_xyzw_to_rotmat(rs.quat)
    cen0 = G.centers_from_pose(Rf, rs.trans) + rng.normal(size=(rs.F, 3)) * 0.3
    pts0 = rs.points + rng.normal(size=rs.points.shape) * 0.3
    u_gt = -np.einsum("sji,sj->si", G.quat_xyzw_to_rotmat(rs.sensor_quat), rs.sensor_trans)     # centre in the rig frame
    unk = np.array([False, False, True])
    u0 = np.zeros((3, 3)); u0[2] = u_gt[2] + rng.normal(size=3) * 0.1
    opts = E.GlobalPositionerOptions(generate_random_positions=False, generate_random_points=False, generate_scales=True)
    opts.solver_options.pcg_rel_tolerance = 1e-12
    opts.solver_options.pcg_max_iterations = 3000
    prob = E.PositioningProblem(rs.quat, rs.pt_obs_begin, rs.obs_frame, bear, centers=cen0.copy(), points=pts0.copy(),
                                obs_sensor=rs.obs_sensor, sensor_quat=rs.sensor_quat, sensor_trans=rs.sensor_trans.copy(),
                                sensor_unknown=unk, rig_centers=u0.copy())
    gp = E.GlobalPositioner(opts)
    assert gp.Solve(prob)
    st_known = rs.sensor_trans.copy(); st_known[2] = 0.0
    t_obs, t_rig = E.rig_world_terms(rs.quat, rs.sensor_quat, st_known, bear, rs.obs_frame, rs.obs_sensor)
    ru = dict(obs_sensor=np.where(rs.obs_sensor == 2, 0, -1).astype(np.int64), R_rw=Rf[rs.obs_frame], centers=u0[2:3].copy())
    x, summ = GPO.solve_gp(cen0, pts0, rs.pt_obs_begin, rs.obs_frame, t_obs, None, GPO.GPOptions(), None, obs_offset=t_rig,
                           rig_unknown=ru)
    st = gp.summary
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert abs(st.final_cost - summ.final_cost) <= 1e-6 * max(summ.final_cost, 1e-12), (st.final_cost, summ.final_cost)
    assert np.abs(prob.centers - x["centers"]).max() < 1e-5
    assert np.abs(prob.rig_centers[2] - x["rig_centers"][0]).max() < 1e-5
    # ConvertResults (.cc:578-582): the estimated centre becomes the cam_from_rig translation -R_cr u
    want_t = -G.quat_xyzw_to_rotmat(rs.sensor_quat[2:3])[0] @ prob.rig_centers[2]
    assert np.abs(prob.sensor_trans[2] - want_t).max() < 1e-12
    assert np.array_equal(prob.sensor_trans[:2], rs.sensor_trans[:2])


def test_rotation_averaging_with_unknown_cam_from_rig():
    """global_rotation_averaging.cc:173-245,425-440,646-693: sensor 1 of a 3-camera rig is calibrated, sensor 2 is not --
    its cam_from_rig rotation is an extra 3-dof node shared by all frames, updated by quaternion averaging over the
    frames.  Device vs oracle (same L1 / IRLS iteration counts, same rotations), and recovery of the extrinsic rotation."""
    rs = S.make_rig_scene(16, 3, 10, seed=13)
    Ri, _ = rs.image_poses()
    rng = np.random.default_rng(6)
    n_img = rs.F * rs.S
    ei, ej = np.triu_indices(n_img, 1)
    sel = rng.uniform(size=len(ei)) < 0.3
    ei, ej = ei[sel].astype(np.int32), ej[sel].astype(np.int32)
    noise = G.so3_exp(rng.normal(size=(len(ei), 3)) * np.radians(0.3))
    R_rel = noise @ Ri[ej] @ np.swapaxes(Ri[ei], -1, -2)
    vg_img = S.ViewGraph(n_img, ei, ej, R_rel, np.ones(len(ei)), Ri)
    img_frame = np.repeat(np.arange(rs.F), rs.S)
    img_sensor = np.tile(np.arange(rs.S), rs.F)
    known = np.array([True, True, False])
    g = E.rig_view_graph_unknown(vg_img, img_frame, img_sensor, rs.sensor_quat, known)
    assert g["n_cams"] == 1 and (g["eci"] >= rs.F).any() and ((g["ei"] == g["ej"]) & ((g["eci"] >= 0) | (g["ecj"] >= 0))).any()
    Rf = G.quat_xyzw_to_rotmat(rs.quat)
    Rs = G.quat_xyzw_to_rotmat(rs.sensor_quat)
    R_f0 = G.so3_exp(rng.normal(size=(rs.F, 3)) * 0.03) @ Rf
    R_f0[0] = Rf[0]
    # the reference starts the camera from the spanning-tree estimate (rotation_initializer.cc:45-89 -> .cc:186-190), i.e.
    # near the truth; the frames-only convergence test (.cc:758-772) does not wait for a camera that starts far away
    R_c0 = G.so3_exp(rng.normal(size=(1, 3)) * 0.03) @ Rs[2:3]
    est = E.RotationEstimator(E.RotationEstimatorOptions(pcg_rel_tolerance=1e-12))
    ok, R_frames, R_cams = E.estimate_rotations_rig_unknown(est, g, R_f0, R_c0)
    assert ok
    cam_frames = [g["cam_frames"][g["cam_frames_begin"][c]:g["cam_frames_begin"][c + 1]] for c in range(g["n_cams"])]
    theta0 = np.concatenate([G.so3_log(R_f0), G.so3_log(R_c0)])
    th, info = RO.estimate_rotations_rig_unknown(g["n_frames"], g["n_cams"], g["ei"], g["ej"], g["eci"], g["ecj"], g["R_rel"],
                                                 theta0, cam_frames)
    assert (est.summary.l1_iterations, est.summary.irls_iterations) == (info["l1_iterations"], info["irls_iterations"])
    assert np.abs(R_frames - G.so3_exp(th[:rs.F])).max() < 1e-7
    assert np.abs(R_cams - G.so3_exp(th[rs.F:])).max() < 1e-7
    assert G.rotation_angle_deg(R_cams, Rs[2:3]).max() < 1.0          # the extrinsic rotation is recovered (0.3 deg noise)
