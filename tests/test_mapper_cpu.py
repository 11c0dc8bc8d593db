"""Host helpers of the mapper driver (glomap_b200/mapper.py): observation compaction (what the track filters do to
Track::observations), RelPoseFilter::FilterRotations (processors/relpose_filter.cc:7-33) and the largest connected
component (scene/view_graph.cc:56)."""
import numpy as np

from glomap_b200 import geometry as G, mapper as M, synthetic as S


def test_compact_observations_and_drop_tracks():
    sc = S.make_scene(8, 100, mean_track_len=4, seed=1)
    keep = np.ones(sc.N, bool)
    keep[::7] = False
    c = M.compact_observations(sc, keep)
    assert c.N == int(keep.sum()) == c.pt_obs_begin[-1] and c.P == sc.P
    pt = np.repeat(np.arange(sc.P), np.diff(sc.pt_obs_begin))
    assert np.array_equal(np.diff(c.pt_obs_begin), np.bincount(pt[keep], minlength=sc.P))
    assert np.array_equal(c.obs_xy, sc.obs_xy[keep]) and np.array_equal(c.points, sc.points)
    kt = np.ones(sc.P, bool); kt[[0, 5, sc.P - 1]] = False
    d = M.drop_tracks(sc, kt)
    lens = np.diff(d.pt_obs_begin)
    assert (lens[[0, 5, sc.P - 1]] == 0).all() and np.array_equal(lens[kt], np.diff(sc.pt_obs_begin)[kt])


def test_filter_rotations_and_connected_component():
    sc = S.make_scene(10, 200, mean_track_len=5, seed=2)
    vg = S.view_graph_from_scene(sc, min_shared=5, noise_deg=0.0)
    R = G.quat_xyzw_to_rotmat(sc.quat)
    assert M.filter_rotations(vg, R, 1.0).all()
    bad = vg.R_rel.copy()
    bad[3] = G.so3_exp(np.array([[0.0, np.radians(30.0), 0.0]]))[0] @ bad[3]
    vg2 = S.ViewGraph(vg.n_images, vg.ei, vg.ej, bad, vg.weight, vg.R_gt)
    valid = M.filter_rotations(vg2, R, 10.0)
    assert not valid[3] and valid.sum() == vg.E - 1
    # two components: {0,1,2} and {3,4}
    m = M.largest_connected_component(5, np.array([0, 1, 3]), np.array([1, 2, 4]))
    assert m.tolist() == [True, True, True, False, False]
