"""Device-resident NormalizeReconstruction / UndistortImages against the host restatements of
glomap/processors/reconstruction_normalizer.cc:5-104 and image_undistorter.cc:7-53 (glomap_b200/processors.py, which
tests/test_processors_cpu.py pins to closed-form expectations)."""
import numpy as np
import pytest

from glomap_b200 import estimators as E, geometry as G, processors as PR, synthetic as S

pytestmark = pytest.mark.gpu


def _problem(sc):
    prob = E.BAProblem(E.default_context(), sc, 3, E.first_frame_mask(sc.C))
    prob.set_state(sc.intr_params, sc.quat, sc.trans, sc.points)
    return prob


@pytest.mark.parametrize("model", [S.SIMPLE_PINHOLE, S.PINHOLE, S.SIMPLE_RADIAL, S.RADIAL])
def test_undistort_matches_host_and_feeds_the_filters(model):
    sc = S.make_scene(25, 900, mean_track_len=6, seed=61, pixel_sigma=0.5, model=model, num_intrinsics=3)
    prob = _problem(sc)
    want = PR.undistort_images(sc)
    got = prob.undistort()
    assert np.abs(got - want).max() < 1e-14
    assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 1e-15
    # the bearing-based filters give the same masks from the resident bearings as from uploaded ones
    k1, n1 = prob.filter_angle(want, 1.0)
    k2, n2 = prob.filter_angle("resident", 1.0)
    assert np.array_equal(k1, k2) and n1 == n2
    k3, n3 = prob.filter_reprojection(1e-3, bearings=want)
    k4, n4 = prob.filter_reprojection(1e-3, bearings="resident")
    assert np.array_equal(k3, k4) and n3 == n4
    prob.free()


@pytest.mark.parametrize("fixed_scale", [False, True])
def test_normalize_matches_host(fixed_scale):
    sc = S.make_scene(200, 3000, mean_track_len=6, seed=62, pixel_sigma=0.5)
    prob = _problem(sc)
    ref = sc.copy()
    s_ref, t_ref = PR.normalize_reconstruction(ref, fixed_scale=fixed_scale)
    s_dev, t_dev = prob.normalize(fixed_scale=fixed_scale)
    assert abs(s_dev - s_ref) <= 1e-12 * s_ref and np.abs(t_dev - t_ref).max() <= 1e-10 * max(1.0, np.abs(t_ref).max())
    intr, q, t, X = prob.get_state()
    assert np.abs(G.quat_xyzw_to_rotmat(q) - G.quat_xyzw_to_rotmat(sc.quat)).max() < 1e-15    # rotations untouched (identity-rotation similarity)
    assert np.abs(t - ref.trans).max() < 1e-9 and np.abs(X - ref.points).max() < 1e-9
    # the robust box of the normalised scene has the requested extent and is centred
    c = G.centers_from_pose(G.quat_xyzw_to_rotmat(q), t)
    srt = np.sort(c.astype(np.float32), axis=0)
    i0, i1 = int(0.1 * (len(c) - 1)), int(0.9 * (len(c) - 1))
    if not fixed_scale:
        assert abs(np.linalg.norm(srt[i1].astype(float) - srt[i0].astype(float)) - 10.0) < 1e-4
    assert np.abs(srt[i0:i1 + 1].astype(float).mean(0)).max() < 1e-4
    # a second call is (numerically) the identity
    s2, t2 = prob.normalize(fixed_scale=fixed_scale)
    assert abs(s2 - 1) < 1e-5 and np.abs(t2).max() < 1e-4
    prob.free()


def test_normalize_rig_problem_scales_cam_from_rig():
    rsc = S.make_rig_scene(12, 3, 600, mean_track_len=6, seed=63, pixel_sigma=0.3)
    prob = E.BAProblem(E.default_context(), rsc, 3, E.first_frame_mask(rsc.F))
    prob.set_state(rsc.intr_params, rsc.quat, rsc.trans, rsc.points)
    ref = rsc.copy()
    s_ref, t_ref = PR.normalize_reconstruction(ref)
    s_dev, t_dev = prob.normalize()
    assert abs(s_dev - s_ref) <= 1e-12 * s_ref and np.abs(t_dev - t_ref).max() <= 1e-10 * max(1.0, np.abs(t_ref).max())
    intr, q, t, X = prob.get_state()
    assert np.abs(t - ref.trans).max() < 1e-9 and np.abs(X - ref.points).max() < 1e-9
    sq, st = prob.get_sensor_poses()
    assert np.abs(st - ref.sensor_trans).max() < 1e-12
    prob.free()
