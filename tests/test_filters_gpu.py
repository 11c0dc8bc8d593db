"""Track filters on the resident BA problem vs the numpy restatement of
glomap/processors/track_filter.cc (bit-exact masks: pure threshold tests)."""
import numpy as np
import pytest

from glomap_b200 import estimators as E, synthetic as S
from oracle import filter_oracle as FO

pytestmark = pytest.mark.gpu


def _problem(sc):
    prob = E.BAProblem(E.default_context(), sc, 3, E.first_frame_mask(sc.C))
    prob.set_state(sc.intr_params, sc.quat, sc.trans, sc.points)
    return prob


def _scene():
    sc = S.make_scene(30, 1500, mean_track_len=6, seed=51, pixel_sigma=1.0, model=S.SIMPLE_RADIAL, num_intrinsics=2)
    rng = np.random.default_rng(0)
    bad = rng.uniform(size=sc.N) < 0.05
    sc.obs_xy[bad] += rng.normal(scale=30.0, size=(int(bad.sum()), 2))      # outlier observations
    sc.quat = sc.quat / np.linalg.norm(sc.quat, axis=1, keepdims=True)
    return sc


def test_reprojection_filter_matches_reference_restatement():
    sc = _scene()
    prob = _problem(sc)
    for thr in (1.0, 3.0, 12.0):                                            # mapper uses max(3 - ite, 1) * thr
        keep, cnt = prob.filter_reprojection(thr)
        k0, c0 = FO.filter_reprojection(sc, thr, S.project)
        assert np.array_equal(keep, k0) and cnt == c0
    assert 0 < cnt < sc.P
    prob.free()


def test_angle_filter_matches_reference_restatement():
    sc = _scene()
    b = S.bearings_from_scene(sc)
    cal = (np.arange(sc.C) % 3 != 0).astype(np.uint8)
    prob = _problem(sc)
    for thr in (0.05, 1.0):
        keep, cnt = prob.filter_angle(b, thr, cal)
        k0, c0 = FO.filter_angle(sc, b, thr, cal)
        assert np.array_equal(keep, k0) and cnt == c0
    prob.free()


def test_triangulation_angle_filter_matches_reference_restatement():
    sc = _scene()
    prob = _problem(sc)
    for thr in (1.0, 8.0, 25.0):
        keep, cnt = prob.filter_triangulation_angle(thr)
        k0, c0 = FO.filter_triangulation_angle(sc, thr)
        assert np.array_equal(keep, k0) and cnt == c0
    prob.free()


def test_points_behind_cameras_are_dropped():
    sc = _scene()
    sc.points[:50] *= -40.0        # far behind most cameras -> z < EPS (track_filter.cc:20,71)
    prob = _problem(sc)
    keep, cnt = prob.filter_reprojection(1e9)
    k0, c0 = FO.filter_reprojection(sc, 1e9, S.project)
    assert np.array_equal(keep, k0) and cnt == c0 and not keep.all()
    prob.free()


def test_normalized_reprojection_filter_matches_oracle():
    sc = S.make_scene(12, 400, mean_track_len=5, seed=3, pixel_sigma=2.0)
    st = S.perturb_scene(sc, rot_deg=0.2)
    bear = S.bearings_from_scene(st)
    prob = E.BAProblem(E.default_context(), st, 3)
    prob.set_state(st.intr_params, st.quat, st.trans, st.points)
    for thr in (1e-3, 1e-2, 1e-1):
        keep, n = prob.filter_reprojection(thr, bear)
        k0, n0 = FO.filter_reprojection_normalized(st, bear, thr)
        assert np.array_equal(keep, k0) and n == n0
    prob.free()
