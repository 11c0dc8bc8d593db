"""CPU tests of the BA oracle (oracle/): the reference holds no golden vectors
for BundleAdjuster, so the oracle is pinned to synthetic
ground truth with the reference's own end-to-end thresholds
(global_mapper_test.cc:84-86: rotation < 1e-2 deg, centre < 1e-4) and its two
implementations (numpy + splu, C/OpenMP + dense Schur) are pinned to each other."""
import numpy as np
import pytest

from glomap_b200 import geometry as G, synthetic as S
from oracle import ba_oracle as B, ba_oracle_fast as F


def _args(sc, init):
    return (init.quat, init.trans, init.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr, sc.intr_model,
            sc.intr_params)


def _mask(C):
    m = np.zeros(C, np.uint8)
    m[0] = 3
    return m


@pytest.mark.parametrize("model", [S.SIMPLE_PINHOLE, S.PINHOLE, S.SIMPLE_RADIAL, S.RADIAL])
def test_jacobian_matches_finite_differences(model):
    sc = S.make_scene(8, 60, mean_track_len=5, seed=3, model=model, num_intrinsics=2)
    init = S.perturb_scene(sc)
    prob = B.BAProblem(*_args(sc, init), B.BAOptions(thres_loss_function=1e9, optimize_intrinsics=True,
                                                     optimize_principal_point=True), _mask(sc.C))
    _, _, J = prob.evaluate(prob.x0, True)
    rng = np.random.default_rng(0)
    d = rng.normal(size=prob.ncols) * 1e-6
    _, r2, _ = prob.evaluate(prob.plus(prob.x0, d), False)
    _, r3, _ = prob.evaluate(prob.plus(prob.x0, -d), False)
    assert np.abs((r2 - r3) / 2 - J @ d).max() < 1e-9 * max(1.0, np.abs(J @ d).max() / 1e-6 * 1e-3)


def test_oracle_recovers_ground_truth_noise_free():
    sc = S.make_scene(30, 800, mean_track_len=6, seed=1)
    init = S.perturb_scene(sc)
    x, summ = B.solve_ba(*_args(sc, init), cam_const_mask=_mask(sc.C))
    rot, cen, _ = G.compare_reconstructions(G.quat_xyzw_to_rotmat(x["quat"]), x["trans"],
                                            G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    assert summ.usable and rot < 1e-2 and cen < 1e-4          # global_mapper_test.cc:84-86
    assert summ.final_cost < 1e-8 * summ.initial_cost


def test_oracle_noisy_within_reference_noisy_thresholds():
    sc = S.make_scene(30, 800, mean_track_len=6, seed=1, pixel_sigma=0.5)
    init = S.perturb_scene(sc)
    x, summ = B.solve_ba(*_args(sc, init), cam_const_mask=_mask(sc.C))
    rot, cen, _ = G.compare_reconstructions(G.quat_xyzw_to_rotmat(x["quat"]), x["trans"],
                                            G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    assert rot < 1e-1 and cen < 1e-1                           # global_mapper_test.cc:213-215
    assert summ.termination == "function tolerance"


def test_oracle_intrinsics_refinement():
    sc = S.make_scene(20, 600, mean_track_len=6, seed=4, model=S.SIMPLE_RADIAL)
    init = S.perturb_scene(sc, rot_deg=0.2)
    init.intr_params[:, 0] *= 1.01
    init.intr_params[:, 3] = 0.0
    x, summ = B.solve_ba(*_args(sc, init), B.BAOptions(optimize_intrinsics=True), _mask(sc.C))
    assert abs(x["intr"][0, 0] / sc.intr_params[0, 0] - 1) < 1e-6
    assert abs(x["intr"][0, 3] - sc.intr_params[0, 3]) < 1e-6
    assert np.allclose(x["intr"][0, 1:3], sc.intr_params[0, 1:3])   # principal point held (bundle_adjustment.cc:273-286)


@pytest.mark.parametrize("sigma", [0.0, 0.5])
def test_c_port_matches_numpy_oracle(sigma):
    sc = S.make_scene(25, 700, mean_track_len=6, seed=2, pixel_sigma=sigma, model=S.RADIAL, num_intrinsics=3)
    init = S.perturb_scene(sc)
    x, s1 = B.solve_ba(*_args(sc, init), cam_const_mask=_mask(sc.C))
    y, s2 = F.solve_ba_fast(*_args(sc, init), cam_const_mask=_mask(sc.C))
    assert s1.iterations == s2.iterations and s1.termination == s2.termination
    assert np.allclose(s1.costs, s2.costs, rtol=1e-9, atol=1e-18)
    for k in ("quat", "trans", "points"):
        assert np.abs(x[k] - y[k]).max() < 1e-9


def test_short_tracks_are_skipped():
    """bundle_adjustment.cc:122: tracks with < min_num_view_per_track observations do not enter the problem."""
    sc = S.make_scene(12, 200, mean_track_len=4, seed=5)
    init = S.perturb_scene(sc)
    lens = np.diff(sc.pt_obs_begin)
    x, _ = B.solve_ba(*_args(sc, init), B.BAOptions(min_num_view_per_track=4), _mask(sc.C))
    short = lens < 4
    assert short.any() and np.array_equal(x["points"][short], init.points[short])
