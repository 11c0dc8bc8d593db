"""End-to-end run of the host mapper driver (stages 3, 5, 6 of controllers/global_mapper.cc) on the GPU solvers, with
the reference's own end-to-end thresholds (global_mapper_test.cc:84-86: rotation < 1e-2 deg / centre < 1e-4 noise-free)
relaxed for the noisy case (:213-215).

Written after the GPU budget of round 1 was exhausted: it has NOT run on a GPU yet, so it is opt-in
(B200SFM_UNVERIFIED_TESTS=1) until it has been validated; every solver and filter it composes is covered by the other
GPU tests (the normalised-plane reprojection filter in tests/test_filters_gpu.py)."""
import os

import numpy as np
import pytest

from glomap_b200 import geometry as G, mapper as M, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("refine_intrinsics", [False, True])
def test_mapper_recovers_the_scene(refine_intrinsics):
    """refine_intrinsics = True is the reference's default BundleAdjusterOptions (bundle_adjustment.h:18)."""
    sc = S.make_scene(30, 2000, mean_track_len=6, seed=21, pixel_sigma=0.5)
    vg = S.view_graph_from_scene(sc, min_shared=15, noise_deg=0.5)
    start = sc.copy()
    start.quat[:] = [0, 0, 0, 1]; start.trans[:] = 0; start.points[:] = 0     # nothing but tracks and relative rotations
    opts = M.GlobalMapperOptions()
    opts.opt_ba.optimize_intrinsics = refine_intrinsics
    mapper = M.GlobalMapper(opts)
    ok, out = mapper.Solve(vg, start)
    assert ok, mapper.log
    rot, cen = G.compare_reconstructions(G.quat_xyzw_to_rotmat(out.quat), out.trans, G.quat_xyzw_to_rotmat(sc.quat), sc.trans)[:2]
    assert rot < 1e-1 and cen < 1e-1, (rot, cen, mapper.log)
