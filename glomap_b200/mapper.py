"""Host driver over the three estimators: stages 3, 5 and 6 of ``glomap::GlobalMapper::Solve``
(glomap/controllers/global_mapper.cc:79-276) on the flat SoA scene -- rotation averaging (run twice, with
``RelPoseFilter::FilterRotations`` and the largest connected component in between), global positioning followed by the
three track filters and ``NormalizeReconstruction``, then the staged bundle adjustment loop (positions only, then
rotations too; normalise; reprojection filters with the tightening threshold ``max(3 - ite, 1) * thr``).

It mirrors the reference's control flow so that the GPU solvers are exercised in the order, and with the option
mutations, the real mapper uses; it is host glue (as in the reference) and owns no numerics: every solve and every
filter goes through ``libb200sfm.so``.  Trivial frames only.  Not covered (they are COLMAP / PoseLib code in the
reference): view-graph calibration, relative-pose estimation, track establishment, retriangulation, pruning."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import estimators as E, geometry as geo, processors as PR, synthetic as S


@dataclasses.dataclass
class InlierThresholdOptions:
    """glomap/types.h:18-33."""
    max_angle_error: float = 1.0            # degrees, global positioning
    max_reprojection_error: float = 1e-2    # normalised image plane, bundle adjustment
    min_triangulation_angle: float = 1.0    # degrees
    max_rotation_error: float = 10.0        # degrees, rotation averaging


@dataclasses.dataclass
class GlobalMapperOptions:
    """controllers/global_mapper.h:14-44 (the fields of the stages implemented here)."""
    opt_ra: E.RotationEstimatorOptions = dataclasses.field(default_factory=E.RotationEstimatorOptions)
    opt_gp: E.GlobalPositionerOptions = dataclasses.field(default_factory=E.GlobalPositionerOptions)
    opt_ba: E.BundleAdjusterOptions = dataclasses.field(default_factory=E.BundleAdjusterOptions)
    inlier_thresholds: InlierThresholdOptions = dataclasses.field(default_factory=InlierThresholdOptions)
    num_iteration_bundle_adjustment: int = 3
    skip_rotation_averaging: bool = False
    skip_global_positioning: bool = False
    skip_bundle_adjustment: bool = False


def compact_observations(scene: S.Scene, keep: np.ndarray) -> S.Scene:
    """Drop the observations with keep == False (what the filters do to ``Track::observations``)."""
    keep = np.asarray(keep, bool)
    pt = np.repeat(np.arange(scene.P), np.diff(scene.pt_obs_begin))
    lens = np.bincount(pt[keep], minlength=scene.P)
    out = scene.copy()
    out.obs_cam, out.obs_xy = scene.obs_cam[keep], scene.obs_xy[keep]
    out.pt_obs_begin = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return out


def drop_tracks(scene: S.Scene, keep_track: np.ndarray) -> S.Scene:
    """FilterTrackTriangulationAngle clears the observations of the removed tracks (track_filter.cc:118-121)."""
    pt = np.repeat(np.arange(scene.P), np.diff(scene.pt_obs_begin))
    return compact_observations(scene, np.asarray(keep_track, bool)[pt])


def filter_rotations(vg: S.ViewGraph, R: np.ndarray, max_angle_deg: float) -> np.ndarray:
    """RelPoseFilter::FilterRotations (processors/relpose_filter.cc:7-33): valid-edge mask."""
    R_calc = R[vg.ej] @ np.swapaxes(R[vg.ei], -1, -2)
    return geo.rotation_angle_deg(R_calc, vg.R_rel) <= max_angle_deg


def largest_connected_component(n: int, ei, ej) -> np.ndarray:
    """ViewGraph::KeepLargestConnectedComponents (scene/view_graph.cc:56): image mask."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    g = sp.coo_matrix((np.ones(len(ei)), (ei, ej)), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    return lab == np.bincount(lab).argmax()


def _sub_view_graph(vg: S.ViewGraph, edge_mask) -> S.ViewGraph:
    return S.ViewGraph(vg.n_images, vg.ei[edge_mask], vg.ej[edge_mask], vg.R_rel[edge_mask], np.asarray(vg.weight)[edge_mask],
                       vg.R_gt)


class GlobalMapper:
    def __init__(self, options: GlobalMapperOptions | None = None, ctx: E.Context | None = None):
        self.options_ = options or GlobalMapperOptions()
        self.ctx = ctx
        self.log: list[str] = []

    # -- helpers ------------------------------------------------------------------------------------
    def _filters(self, scene: S.Scene, what) -> S.Scene:
        """Run a list of (kind, threshold) filters on ONE resident problem per filter (the observation set changes
        after each, as in the reference where every filter rewrites Track::observations)."""
        ctx = self.ctx or E.default_context()
        for kind, thr in what:
            prob = E.BAProblem(ctx, scene, self.options_.opt_ba.min_num_view_per_track)
            try:
                prob.set_state(scene.intr_params, scene.quat, scene.trans, scene.points)
                if kind == "angle":
                    keep, n = prob.filter_angle("resident", thr)                  # bearings: the device's own UndistortImages
                    scene = compact_observations(scene, keep)
                elif kind == "reprojection":
                    keep, n = prob.filter_reprojection(thr, "resident")
                    scene = compact_observations(scene, keep)
                else:
                    keep_t, n = prob.filter_triangulation_angle(thr)
                    scene = drop_tracks(scene, keep_t)
            finally:
                prob.free()
            self.log.append(f"filter {kind} thr={thr:g}: {n} tracks changed, {scene.N} observations left")
            self.last_filtered = n
        return scene

    # -- controllers/global_mapper.cc:19-355 (stages 3, 5, 6) -----------------------------------------
    def Solve(self, view_graph: S.ViewGraph, scene: S.Scene):
        """Returns (ok, scene): poses / points / intrinsics of ``scene`` estimated from the relative rotations of
        ``view_graph`` and the tracks of ``scene`` (its poses and points are only used when a stage is skipped)."""
        o, thr = self.options_, self.options_.inlier_thresholds
        scene = scene.copy()
        # 3. rotation averaging: first run for filtering, second for the estimate (:84-116)
        if not o.skip_rotation_averaging:
            vg = view_graph
            for run in range(2):
                ra = E.RotationEstimator(o.opt_ra, self.ctx)
                ok, R = ra.EstimateRotations(vg)
                if not ok:
                    if run == 1:
                        return False, scene
                    continue
                valid = filter_rotations(vg, R, thr.max_rotation_error)
                vg = _sub_view_graph(vg, valid)
                if not largest_connected_component(vg.n_images, vg.ei, vg.ej).all():
                    raise NotImplementedError("images outside the largest connected component must be removed by the caller")
                self.log.append(f"rotation averaging run {run + 1}: {int((~valid).sum())} edges filtered")
            scene.quat = geo.rotmat_to_quat_xyzw_fast(R)
        # 5. global positioning (:143-189)
        if not o.skip_global_positioning:
            bear = PR.undistort_images(scene)
            gp = E.GlobalPositioner(o.opt_gp, self.ctx)
            prob = E.PositioningProblem(scene.quat, scene.pt_obs_begin, scene.obs_cam, bear, centers=None, points=None)
            if not gp.Solve(prob):
                return False, scene
            scene.trans, scene.points = prob.trans, prob.points
            scene = self._filters(scene, [("angle", thr.max_angle_error), ("triangulation", thr.min_triangulation_angle),
                                          ("reprojection", 10 * thr.max_reprojection_error)])
            PR.normalize_reconstruction(scene)
        # 6. bundle adjustment (:191-280)
        if not o.skip_bundle_adjustment:
            ite = 0
            while ite < o.num_iteration_bundle_adjustment:
                ba = E.BundleAdjuster(o.opt_ba, self.ctx)
                inner = ba.GetOptions()
                inner.optimize_rotations = False                          # 6.1 positions only (:207-211)
                if not ba.Solve(scene):
                    return False, scene
                inner.optimize_rotations = o.opt_ba.optimize_rotations    # 6.2 (:217-222)
                if inner.optimize_rotations and not ba.Solve(scene):
                    return False, scene
                self.log.append(f"bundle adjustment iteration {ite + 1}: cost {ba.summary.final_cost:.6g}")
                PR.normalize_reconstruction(scene)
                status, filtered = True, 0                                 # 6.3 (:236-262)
                while status and ite < o.num_iteration_bundle_adjustment:
                    scaling = max(3 - ite, 1)
                    scene = self._filters(scene, [("reprojection", scaling * thr.max_reprojection_error)])
                    filtered += self.last_filtered
                    if filtered > 1e-3 * scene.P:
                        status = False
                    else:
                        ite += 1
                if status:
                    self.log.append("fewer than 0.1% tracks are filtered, stop the iteration")
                    break
                ite += 1
            scene = self._filters(scene, [("reprojection", thr.max_reprojection_error),
                                          ("triangulation", thr.min_triangulation_angle)])
        return True, scene
