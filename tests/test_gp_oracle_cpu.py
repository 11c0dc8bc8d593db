"""CPU tests pinning the BATA global-positioning oracle to synthetic ground
truth (the reference holds no golden vectors for GlobalPositioner): from random
positions it must recover the camera centres up to a similarity
(global_mapper_test.cc:84-86: centre error < 1e-4 after alignment)."""
import numpy as np

from glomap_b200 import geometry as G, synthetic as S
from oracle import gp_oracle as GP


def _setup(seed, C=12, P=150, sigma=0.0):
    sc = S.make_scene(C, P, mean_track_len=5, seed=seed, pixel_sigma=sigma)
    t_obs = GP.world_bearings(sc.quat, S.bearings_from_scene(sc), sc.obs_cam)
    rng = np.random.default_rng(seed)
    c0 = 100 * rng.uniform(-1, 1, size=(sc.C, 3))
    X0 = 100 * rng.uniform(-1, 1, size=(sc.P, 3))                      # global_positioning.cc:158-159,262
    return sc, t_obs, c0, X0


def test_random_start_recovers_centres():
    sc, t_obs, c0, X0 = _setup(3)
    x, summ = GP.solve_gp(c0, X0, sc.pt_obs_begin, sc.obs_cam, t_obs)
    cg = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    s, R, t = G.umeyama_sim3(x["centers"], cg)
    assert np.linalg.norm((s * (R @ x["centers"].T)).T + t - cg, axis=1).max() < 1e-4
    assert summ.final_cost < 1e-10 * summ.initial_cost
    assert x["scales"].min() >= GP.SCALE_LOWER_BOUND and x["scales"][0] == 1.0     # .cc:373, :484-489


def test_analytic_jacobian_matches_finite_differences():
    sc, t_obs, c0, X0 = _setup(4)
    prob = GP.GPProblem(c0, X0, sc.pt_obs_begin, sc.obs_cam, t_obs, None, GP.GPOptions(thres_loss_function=1e9))
    _, r, J = prob.evaluate(prob.x0, True)
    rng = np.random.default_rng(0)
    d = rng.normal(size=prob.ncols) * 1e-6
    x1 = prob.plus(prob.x0, d); x2 = prob.plus(prob.x0, -d)
    _, r1, _ = prob.evaluate(x1, False); _, r2, _ = prob.evaluate(x2, False)
    assert np.abs((r1 - r2) / 2 - J @ d).max() < 1e-9


def test_uncalibrated_cameras_halve_the_loss():
    sc, t_obs, c0, X0 = _setup(5)
    cal = np.zeros(sc.C, np.uint8)
    p_all = GP.GPProblem(c0, X0, sc.pt_obs_begin, sc.obs_cam, t_obs, None, GP.GPOptions())
    p_unc = GP.GPProblem(c0, X0, sc.pt_obs_begin, sc.obs_cam, t_obs, cal, GP.GPOptions())
    assert abs(p_unc.evaluate(p_unc.x0, False)[0] - 0.5 * p_all.evaluate(p_all.x0, False)[0]) < 1e-9   # ScaledLoss 0.5 (.cc:242-247)
