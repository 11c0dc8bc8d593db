"""Host-side processors either side of the solvers: NormalizeReconstruction
(glomap/processors/reconstruction_normalizer.cc:5-104) and UndistortImages (image_undistorter.cc:7-53)."""
import numpy as np

from glomap_b200 import geometry as G, processors as PR, synthetic as S
from oracle import ba_oracle as B


def _cost(sc):
    p = B.BAProblem(sc.quat, sc.trans, sc.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr, sc.intr_model,
                    sc.intr_params, B.BAOptions())
    return p.evaluate(p.x0, False)[0]


def test_normalize_reconstruction_is_a_similarity_and_fixes_extent():
    sc = S.make_scene(40, 300, mean_track_len=5, seed=2, pixel_sigma=0.5)
    c0 = _cost(sc)
    before = sc.copy()
    scale, t = PR.normalize_reconstruction(sc)
    # reprojection residuals are invariant under the similarity
    assert abs(_cost(sc) - c0) <= 1e-9 * c0
    assert np.abs(sc.points - (scale * before.points + t)).max() < 1e-12
    c = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    assert np.abs(c - (scale * G.centers_from_pose(G.quat_xyzw_to_rotmat(before.quat), before.trans) + t)).max() < 1e-9
    # robust box (10..90 %) of the centres now has diagonal 10 and its trimmed mean sits at the origin (float32 sorting)
    cs = np.sort(c, axis=0)
    n = len(cs); i0, i1 = int(0.1 * (n - 1)), int(0.9 * (n - 1))
    assert abs(np.linalg.norm(cs[i1] - cs[i0]) - 10.0) < 1e-4
    assert np.abs(cs[i0:i1 + 1].mean(0)).max() < 1e-4
    # idempotent
    s2, t2 = PR.normalize_reconstruction(sc)
    assert abs(s2 - 1.0) < 1e-5 and np.abs(t2).max() < 1e-4


def test_normalize_fixed_scale_and_tiny_inputs():
    sc = S.make_scene(3, 30, mean_track_len=3, seed=3)     # <= 3 images: the whole range is used (.cc:36-39)
    before = sc.copy()
    scale, t = PR.normalize_reconstruction(sc, fixed_scale=True)
    assert scale == 1.0
    c = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    assert np.abs(c.mean(0)).max() < 1e-5 and np.abs(sc.points - (before.points + t)).max() < 1e-12


def test_normalize_rig_scene_scales_cam_from_rig():
    rs = S.make_rig_scene(10, 3, 100, seed=4, pixel_sigma=0.3)
    before = rs.copy()
    scale, t = PR.normalize_reconstruction(rs)
    assert np.abs(rs.sensor_trans - scale * before.sensor_trans).max() < 1e-15
    # the image poses moved by the same similarity: residuals unchanged
    a, b = before.images_scene(), rs.images_scene()
    assert abs(_cost(a) - _cost(b)) <= 1e-9 * _cost(a)


def test_undistort_images_inverts_the_projection():
    for model in (S.SIMPLE_PINHOLE, S.PINHOLE, S.SIMPLE_RADIAL, S.RADIAL):
        sc = S.make_scene(6, 80, mean_track_len=4, seed=5, model=model)
        bear = PR.undistort_images(sc)
        assert np.abs(np.linalg.norm(bear, axis=1) - 1).max() < 1e-14
        pt = np.repeat(np.arange(sc.P), np.diff(sc.pt_obs_begin))
        Xc = np.einsum("nij,nj->ni", G.quat_xyzw_to_rotmat(sc.quat)[sc.obs_cam], sc.points[pt]) + sc.trans[sc.obs_cam]
        assert np.abs(bear - Xc / np.linalg.norm(Xc, axis=1, keepdims=True)).max() < 1e-9
