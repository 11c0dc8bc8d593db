"""Host-side track establishment against the rules of glomap::TrackEngine
(glomap/controllers/track_establishment.cc:5-234)."""
import numpy as np

from glomap_b200 import synthetic as S, track_establishment as T


def _pairs_from_scene(sc, rng, drop=0.0):
    """Feature tables + pairwise matches from ground-truth tracks: feature f of image i is its f-th observation."""
    order = np.argsort(sc.obs_cam, kind="stable")
    counts = np.bincount(sc.obs_cam, minlength=sc.C)
    starts = np.concatenate([[0], np.cumsum(counts)])
    feat_of_obs = np.empty(sc.N, np.int64)
    feat_of_obs[order] = np.arange(sc.N) - np.repeat(starts[:-1], counts)
    features = {i + 1: sc.obs_xy[order][starts[i]:starts[i + 1]] for i in range(sc.C)}    # image ids are 1-based
    pt = np.repeat(np.arange(sc.P), np.diff(sc.pt_obs_begin))
    by_pair = {}
    for p in range(sc.P):
        a, b = sc.pt_obs_begin[p], sc.pt_obs_begin[p + 1]
        for u in range(a, b):
            for v in range(u + 1, b):
                if rng.uniform() < drop:
                    continue
                i, j = int(sc.obs_cam[u]), int(sc.obs_cam[v])
                fu, fv = int(feat_of_obs[u]), int(feat_of_obs[v])
                if i > j:
                    i, j, fu, fv = j, i, fv, fu
                by_pair.setdefault((i + 1, j + 1), []).append((fu, fv))
    pairs = []
    for (i, j), m in sorted(by_pair.items()):
        m = np.asarray(m, np.int64)
        # two junk rows that are NOT inliers must be ignored
        mm = np.concatenate([m, [[0, 0], [1, 1]]])
        pairs.append(T.ImagePairMatches(i, j, mm, np.arange(len(m))))
    return features, pairs, feat_of_obs, pt


def test_ground_truth_tracks_are_recovered():
    sc = S.make_scene(10, 120, mean_track_len=4, seed=3)
    rng = np.random.default_rng(0)
    features, pairs, feat_of_obs, pt = _pairs_from_scene(sc, rng, drop=0.3)   # a spanning subset of the matches suffices
    tracks, discarded = T.establish_full_tracks(pairs, features)
    assert discarded == 0
    got = set()
    for t in range(len(tracks)):
        im, ft = tracks.observations(t)
        got.add(frozenset(zip(im.tolist(), ft.tolist())))
        assert int(tracks.track_ids[t]) == min((int(i) << 32) | int(f) for i, f in zip(im, ft))
    want = set()
    for p in range(sc.P):
        a, b = sc.pt_obs_begin[p], sc.pt_obs_begin[p + 1]
        want.add(frozenset((int(sc.obs_cam[o]) + 1, int(feat_of_obs[o])) for o in range(a, b)))
    # with 30 % of the pairwise matches dropped a track may split; every recovered track is a subset of a true one
    assert all(any(g <= w for w in want) for g in got)
    full, _ = T.establish_full_tracks(_pairs_from_scene(sc, rng, drop=0.0)[1], features)
    got_full = {frozenset(zip(*[x.tolist() for x in full.observations(t)])) for t in range(len(full))}
    assert got_full == want


def test_invalid_pairs_are_ignored_and_inconsistent_tracks_discarded():
    features = {1: np.array([[0.0, 0.0], [100.0, 0.0], [3.0, 0.0]]), 2: np.array([[5.0, 5.0], [50.0, 5.0]]),
                3: np.array([[9.0, 9.0], [1.0, 1.0]])}
    P = T.ImagePairMatches
    # track A: (1,0)-(2,0)-(3,0).  A wrong match (3,0)-(1,1) pulls (1,1) into it: image 1 then holds features 0 and 1,
    # 100 px apart > thres_inconsistency -> the whole track is discarded (observations cleared, id kept)
    pairs = [P(1, 2, np.array([[0, 0]]), np.array([0])), P(2, 3, np.array([[0, 0]]), np.array([0])),
             P(1, 3, np.array([[1, 0]]), np.array([0])),
             P(1, 2, np.array([[2, 1]]), np.array([0]), is_valid=False)]          # invalid pair: ignored
    tracks, discarded = T.establish_full_tracks(pairs, features)
    assert discarded == 1 and len(tracks) == 1 and tracks.begin[-1] == 0
    assert int(tracks.track_ids[0]) == (1 << 32) | 0
    # two features of ONE image closer than the threshold stay in the track (duplicates are allowed, :126-134)
    pairs2 = [P(1, 2, np.array([[0, 0]]), np.array([0])), P(2, 3, np.array([[0, 0]]), np.array([0])),
              P(1, 3, np.array([[2, 0]]), np.array([0]))]                         # (1,2) is 3 px from (1,0)
    tracks2, discarded2 = T.establish_full_tracks(pairs2, features)
    assert discarded2 == 0 and len(tracks2) == 1 and tracks2.begin[-1] == 4
    im, ft = tracks2.observations(0)
    assert sorted(zip(im.tolist(), ft.tolist())) == [(1, 0), (1, 2), (2, 0), (3, 0)]


def _toy_tracks():
    # 4 tracks over images 1..4: lengths 4, 3, 3, 2
    ids = np.array([10, 20, 30, 40], np.uint64)
    begin = np.array([0, 4, 7, 10, 12])
    img = np.array([1, 2, 3, 4, 1, 2, 3, 2, 3, 4, 1, 2], np.uint32)
    return T.Tracks(ids, begin, img, np.arange(12, dtype=np.uint32))


def test_selection_rules():
    tr = _toy_tracks()
    o = T.TrackEstablishmentOptions()                       # min views 3, quota -1 == no quota (unsigned comparison)
    sel = T.find_tracks_for_problem(tr, [1, 2, 3, 4], o)
    assert sel.track_ids.tolist() == [10, 30, 20]           # (length, id) descending; the 2-view track is skipped
    # only registered images count: without image 4 track 30 keeps 2 views and is dropped, track 10 keeps 3
    sel = T.find_tracks_for_problem(tr, [1, 2, 3], o)
    assert sel.track_ids.tolist() == [10, 20] and np.diff(sel.begin).tolist() == [3, 3]
    # per-camera quota 0: a camera accepts a track while its counter <= 0, i.e. exactly one; all four cameras are
    # saturated by the first track and the loop stops (cameras_left == 0)
    sel = T.find_tracks_for_problem(tr, [1, 2, 3, 4], T.TrackEstablishmentOptions(min_num_tracks_per_view=0))
    assert sel.track_ids.tolist() == [10]
    # max_num_tracks: the loop stops once size() > max (one more than the limit is kept, :226)
    sel = T.find_tracks_for_problem(tr, [1, 2, 3, 4], T.TrackEstablishmentOptions(max_num_tracks=0))
    assert sel.track_ids.tolist() == [10]
    sel = T.find_tracks_for_problem(tr, [1, 2, 3, 4], T.TrackEstablishmentOptions(max_num_view_per_track=3))
    assert sel.track_ids.tolist() == [30, 20]


def test_tracks_to_scene_layout():
    sc = S.make_scene(6, 40, mean_track_len=4, seed=5)
    features, pairs, _, _ = _pairs_from_scene(sc, np.random.default_rng(1))
    tracks, _ = T.establish_full_tracks(pairs, features)
    sel = T.find_tracks_for_problem(tracks, range(1, sc.C + 1))
    flat = T.tracks_to_scene(sel, features, range(1, sc.C + 1), sc.cam_intr, sc.intr_model, sc.intr_params)
    assert flat.P == len(sel) and flat.N == sel.begin[-1] and flat.C == sc.C
    assert (np.diff(flat.pt_obs_begin) >= 3).all()
    # the pixel of every observation is the feature the track refers to
    t0 = int(np.argsort(sel.track_ids)[0])
    im, ft = sel.observations(t0)
    assert np.array_equal(flat.obs_xy[:len(im)], np.stack([features[int(i)][int(f)] for i, f in zip(im, ft)]))
