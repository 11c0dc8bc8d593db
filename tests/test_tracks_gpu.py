"""Device track establishment (b200sfm_tracks_establish: union-find over the inlier matches, track collection, inconsistency
rule -- glomap/controllers/track_establishment.cc:5-150) against the host restatement, which tests/test_track_establishment_cpu.py
pins to the reference's rules.  Index work: the comparison is exact."""
import numpy as np
import pytest

import importlib.util
import os

from glomap_b200 import synthetic as S, track_establishment as T

_spec = importlib.util.spec_from_file_location("_te_cpu", os.path.join(os.path.dirname(__file__), "test_track_establishment_cpu.py"))
_te_cpu = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_te_cpu)


def _pairs_from_scene(sc, rng, drop=0.0):   # (features, pairs): the generator of the CPU test (image ids are 1-based)
    return _te_cpu._pairs_from_scene(sc, rng, drop)[:2]

pytestmark = pytest.mark.gpu


def _same(a: T.Tracks, b: T.Tracks):
    assert np.array_equal(a.track_ids, b.track_ids)
    assert np.array_equal(a.begin, b.begin)
    assert np.array_equal(a.obs_image, b.obs_image) and np.array_equal(a.obs_feature, b.obs_feature)


@pytest.mark.parametrize("drop", [0.0, 0.3])
def test_device_tracks_equal_host_tracks(drop):
    sc = S.make_scene(40, 4000, mean_track_len=6, seed=71)
    rng = np.random.default_rng(5)
    features, pairs = _pairs_from_scene(sc, rng, drop=drop)
    want, dis_w = T.establish_full_tracks(pairs, features)
    got, dis_g = T.establish_full_tracks_device(pairs, features)
    _same(got, want)
    assert dis_g == dis_w
    sel_w = T.find_tracks_for_problem(want, range(1, sc.C + 1))
    sel_g = T.find_tracks_for_problem(got, range(1, sc.C + 1))
    _same(sel_g, sel_w)


def test_wrong_matches_merge_and_discard_like_the_host():
    """Random wrong matches glue tracks together; merged tracks that put two distant features into one image are discarded
    (observations cleared, id kept), invalid pairs and pairs without inliers are ignored."""
    sc = S.make_scene(30, 3000, mean_track_len=5, seed=72)
    rng = np.random.default_rng(6)
    features, pairs = _pairs_from_scene(sc, rng, drop=0.1)
    for p in pairs[::7]:                                   # corrupt every 7th pair: shuffle the second column of a few matches
        m = np.array(p.matches)
        k = min(5, len(m))
        m[:k, 1] = rng.permutation(m[:, 1])[:k]
        p.matches = m
    pairs[3].is_valid = False
    pairs[5].inliers = np.zeros(0, np.int64)
    want, dis_w = T.establish_full_tracks(pairs, features)
    got, dis_g = T.establish_full_tracks_device(pairs, features)
    assert dis_w > 0
    _same(got, want)
    assert dis_g == dis_w


def test_no_matches():
    got, dis = T.establish_full_tracks_device([], {})
    assert len(got) == 0 and dis == 0 and got.begin.tolist() == [0]
