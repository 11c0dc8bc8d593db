"""End-to-end run of the mapper driver (glomap_b200/mapper.py: stages 3, 5, 6 of controllers/global_mapper.cc) on the
CPU with the three estimators and the track filters replaced by the ORACLE (test-only fakes, monkeypatched into the
driver): validates the driver's glue -- undistortion, filter/compaction sequence, normalisation between the solves,
staged BA with the option flip -- with the reference's end-to-end thresholds on a small synthetic scene.  The GPU
counterpart (tests/test_mapper_gpu.py) runs the same driver on the CUDA solvers."""
import dataclasses

import numpy as np

from glomap_b200 import estimators as E, geometry as G, mapper as M, synthetic as S
from oracle import ba_oracle as B, filter_oracle as FO, gp_oracle as GPO, ra_oracle as RO


class FakeRA:
    def __init__(self, options, ctx=None):
        self.o = options

    def EstimateRotations(self, vg, R_init=None, fixed=0, gravity=None):
        R0 = E.initialize_from_maximum_spanning_tree(vg, R_init)
        th, info = RO.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
        return not info.get("failed", False), G.so3_exp(th)


class FakeGP:
    def __init__(self, options, ctx=None):
        self.o, self.rng = options, np.random.default_rng(options.seed)

    def Solve(self, prob):
        cen = 100.0 * self.rng.uniform(-1, 1, size=(prob.C, 3))
        pts = 100.0 * self.rng.uniform(-1, 1, size=(prob.P, 3))
        t_obs = E.world_bearings(prob.quat, prob.bearings, prob.obs_cam)
        x, summ = GPO.solve_gp(cen, pts, prob.pt_obs_begin, prob.obs_cam, t_obs, None, GPO.GPOptions())
        prob.centers, prob.points = x["centers"], x["points"]
        prob.trans = -np.einsum("nij,nj->ni", G.quat_xyzw_to_rotmat(prob.quat), prob.centers)
        return True


@dataclasses.dataclass
class _Summary:
    final_cost: float = 0.0
    usable: int = 1


class FakeBA:
    def __init__(self, options, ctx=None):
        self.options_ = dataclasses.replace(options)
        self.summary = _Summary()

    def GetOptions(self):
        return self.options_

    def Solve(self, sc, cam_const_mask=None):
        o = self.options_
        opts = B.BAOptions(optimize_rotations=o.optimize_rotations, optimize_translation=o.optimize_translation,
                           optimize_intrinsics=o.optimize_intrinsics, optimize_points=o.optimize_points)
        x, summ = B.solve_ba(sc.quat, sc.trans, sc.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr, sc.intr_model,
                             sc.intr_params, opts, E.first_frame_mask(sc.C))
        sc.quat, sc.trans, sc.points, sc.intr_params = x["quat"], x["trans"], x["points"], x["intr"]
        self.summary = _Summary(summ.final_cost)
        return True


class FakeBAProblem:
    def __init__(self, ctx, scene, min_views=3, mask=None):
        self.sc = scene

    def set_state(self, intr, quat, trans, points):
        pass

    @staticmethod
    def _bearings(sc, bearings):   # "resident": the problem's own UndistortImages (device side: b200sfm_ba_problem_undistort)
        from glomap_b200 import processors as PR
        return PR.undistort_images(sc) if isinstance(bearings, str) else bearings

    def filter_angle(self, bearings, thr, cal=None):
        return FO.filter_angle(self.sc, self._bearings(self.sc, bearings), thr)

    def filter_reprojection(self, thr, bearings=None):
        return FO.filter_reprojection_normalized(self.sc, self._bearings(self.sc, bearings), thr)

    def filter_triangulation_angle(self, thr):
        return FO.filter_triangulation_angle(self.sc, thr)

    def free(self):
        pass


def test_mapper_driver_end_to_end_with_oracle_solvers(monkeypatch):
    monkeypatch.setattr(M.E, "RotationEstimator", FakeRA)
    monkeypatch.setattr(M.E, "GlobalPositioner", FakeGP)
    monkeypatch.setattr(M.E, "BundleAdjuster", FakeBA)
    monkeypatch.setattr(M.E, "BAProblem", FakeBAProblem)
    monkeypatch.setattr(M.E, "default_context", lambda: None)
    sc = S.make_scene(12, 300, mean_track_len=5, seed=21, pixel_sigma=0.5)
    vg = S.view_graph_from_scene(sc, min_shared=8, noise_deg=0.5)
    start = sc.copy()
    start.quat[:] = [0, 0, 0, 1]; start.trans[:] = 0; start.points[:] = 0      # nothing but tracks and relative rotations
    opts = M.GlobalMapperOptions()
    opts.opt_ba.optimize_intrinsics = False
    mapper = M.GlobalMapper(opts)
    ok, out = mapper.Solve(vg, start)
    assert ok, mapper.log
    rot, cen = G.compare_reconstructions(G.quat_xyzw_to_rotmat(out.quat), out.trans, G.quat_xyzw_to_rotmat(sc.quat), sc.trans)[:2]
    assert rot < 1e-1 and cen < 1e-1, (rot, cen, mapper.log)            # global_mapper_test.cc:213-215 (noisy case)
    # NormalizeReconstruction ran last on the poses: robust extent of the centres is 10
    c = np.sort(G.centers_from_pose(G.quat_xyzw_to_rotmat(out.quat), out.trans), axis=0)
    n = len(c)
    assert abs(np.linalg.norm(c[int(0.9 * (n - 1))] - c[int(0.1 * (n - 1))]) - 10.0) < 1e-3
    assert out.N <= sc.N and out.N > 0.8 * sc.N
