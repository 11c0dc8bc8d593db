#!/usr/bin/env python
"""Kernel-tuning sweep / profiling target (measurement tooling, not part of the product): the config-4 scene is
generated ONCE (cached under /dev/shm), then one worker process per (library variant, carve-out, extra env) reports
CUDA-event times of a warm BundleAdjuster solve.  The worker is also the process profiled by the ncu commands quoted
in profiles/*.md (it starts in seconds, unlike bench.py).
  python profiles/sweep_worker.py                # driver: glomap_b200/libb200sfm.so + build/variants/*.so
  python profiles/sweep_worker.py --worker       # one variant (env B200SFM_LIB / B200SFM_CARVEOUT / SWEEP_REPS)"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/dev/shm/b200sfm_cfg4"
FIELDS = ["quat", "trans", "points", "pt_obs_begin", "obs_cam", "obs_xy", "cam_intr", "intr_model", "intr_params"]


def gen():
    from glomap_b200 import synthetic as S
    t0 = time.time()
    sc = S.make_scene(10_000, 2_000_000, 10.0, seed=1, pixel_sigma=0.5, chunk=50_000)
    init = S.perturb_scene(sc, chunk=50_000)
    os.makedirs(CACHE, exist_ok=True)
    for f in FIELDS:
        np.save(f"{CACHE}/{f}.npy", getattr(sc, f))
    for f in ("quat", "trans", "points"):
        np.save(f"{CACHE}/init_{f}.npy", getattr(init, f))
    print(f"scene generated in {time.time() - t0:.1f} s, N = {sc.N}", flush=True)


def worker(design, reps):
    from glomap_b200 import estimators as E, synthetic as S
    sc = S.Scene(*[np.load(f"{CACHE}/{f}.npy") for f in FIELDS])
    init = {f: np.load(f"{CACHE}/init_{f}.npy") for f in ("quat", "trans", "points")}
    ctx = E.default_context()
    opts = E.BundleAdjusterOptions(optimize_intrinsics=False, profile_kernels=True, design=design)
    opts.solver_options.max_num_iterations = 20
    opts.solver_options.pcg_rel_tolerance = 0.1
    prob = E.BAProblem(ctx, sc, 3, E.first_frame_mask(sc.C))
    prob.set_state(sc.intr_params, init["quat"], init["trans"], init["points"])
    prob.save_state()
    out = []
    for r in range(reps + 2):
        prob.restore_state()
        st = prob.solve(opts)
        if r >= 2:
            out.append(st)
    ms = np.mean([s.ms_total for s in out])
    mv = np.mean([s.ms_matvec / max(s.n_matvec, 1) for s in out])
    li = np.mean([s.ms_linearize / max(s.n_linearize, 1) for s in out])
    s = out[-1]
    n_used = s.num_observations
    print(json.dumps({"lib": os.path.basename(os.environ.get("B200SFM_LIB", "default")),
                      "carve": os.environ.get("B200SFM_CARVEOUT", "max"), "design": design, "ms_solve": round(ms, 2),
                      "ms_matvec": round(mv, 4), "ms_linearize_points": round(li, 4), "lm_its": s.iterations,
                      "pcg_its": s.pcg_iterations, "final_cost": s.final_cost,
                      "Gobs_per_s_per_LMit": round(n_used * s.iterations / ms / 1e6, 4)}), flush=True)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        design = int(os.environ.get("SWEEP_DESIGN", "0"))
        worker(design, int(os.environ.get("SWEEP_REPS", "3")))
        sys.exit(0)
    if not os.path.exists(f"{CACHE}/obs_xy.npy"):
        gen()
    vdir = os.path.join(ROOT, "build", "variants")   # variant libraries built with -DB200_* tunables (scratch, git-ignored)
    libs = [os.path.join(ROOT, "glomap_b200", "libb200sfm.so")] + sorted(
        os.path.join(vdir, f) for f in (os.listdir(vdir) if os.path.isdir(vdir) else []) if f.endswith(".so"))
    carves = os.environ.get("SWEEP_CARVES", "100,75,50,25").split(",")
    extras = os.environ.get("SWEEP_EXTRA", "").split(";")
    for lib in libs:
      for extra in extras:
        for carve in carves:
            env = dict(os.environ, B200SFM_LIB=lib, B200SFM_CARVEOUT=carve)
            if extra:
                k, v = extra.split("=")
                env[k] = v
                print(f"-- {extra}", flush=True)
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, timeout=240, check=False)
            except subprocess.TimeoutExpired:
                print(f"TIMEOUT {lib} {carve}", flush=True)
