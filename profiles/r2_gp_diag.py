#!/usr/bin/env python
"""Diagnostic for the config-2 GP solve (tests/test_config2_gpu.py): where does the LM loop stop, and which variables
carry the remaining projected gradient?  Measurement tooling, not part of the product."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import gp_oracle as GP

sc = S.make_scene(1000, 200_000, 10.0, seed=1, pixel_sigma=0.0, chunk=50_000)
for tol, maxit in ((1e-3, 400), (1e-6, 400)):
    prob = E.PositioningProblem(sc.quat, sc.pt_obs_begin, sc.obs_cam, S.bearings_from_scene(sc))
    opts = E.GlobalPositionerOptions()
    opts.solver_options.pcg_rel_tolerance = tol
    opts.solver_options.pcg_max_iterations = 3000
    opts.solver_options.function_tolerance = 1e-12
    opts.solver_options.max_num_iterations = maxit
    gp = E.GlobalPositioner(opts)
    ok = gp.Solve(prob)
    st = gp.summary
    print(f"tol {tol}: ok {ok} its {st.iterations} successful {st.num_successful_steps} term {st.termination} "
          f"cost {st.initial_cost:.6e} -> {st.final_cost:.6e} pcg {st.pcg_iterations} ms {st.ms_total:.1f}", flush=True)
    t_obs = GP.world_bearings(sc.quat, prob.bearings, sc.obs_cam)
    o = GP.GPProblem(prob.centers, prob.points, sc.pt_obs_begin, sc.obs_cam, t_obs, None, GP.GPOptions(), scales=prob.scales)
    cost, r, J = o.evaluate(o.x0, True)
    g = J.T @ r
    step = np.abs(o.project(o.x0, -g))
    nC = 3 * sc.C; nP = 3 * sc.P
    print("  oracle cost", cost, "ncols", o.ncols, "3C", nC, "3P", nP)
    for name, sl in (("centres", slice(0, nC)), ("points", slice(nC, nC + nP)), ("scales", slice(nC + nP, None))):
        s_ = step[sl]
        if len(s_) == 0: continue
        print(f"  {name}: max {s_.max():.3e} q99.9 {np.quantile(s_, 0.999):.3e} q99 {np.quantile(s_, 0.99):.3e} median {np.median(s_):.3e} "
              f"count>2e-2 {(s_ > 2e-2).sum()} of {len(s_)}")
    print("  scales at bound:", int((prob.scales <= 1e-5 * (1 + 1e-9)).sum()), "of", len(prob.scales), "min", prob.scales.min())
    r2 = (r.reshape(-1, 3) ** 2).sum(1)
    print("  residual^2 per obs: max", r2.max(), "q99.99", np.quantile(r2, 0.9999), "sum", r2.sum(), " #>1e-3:", int((r2 > 1e-3).sum()))
