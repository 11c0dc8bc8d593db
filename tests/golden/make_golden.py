#!/usr/bin/env python
"""Regenerates the fixtures in this directory:  python tests/golden/make_golden.py

PROVENANCE: the reference (colmap/glomap) cannot be built in this image and ships no golden vectors for its three
estimators, so these are NOT reference outputs.  They are small, fully self-contained problems
(inputs included) with the outputs of the CPU oracle (oracle/*.py) at the time of writing.  They pin the oracle -- and,
through the GPU parity tests, the CUDA path -- against drift between rounds; parity with the reference itself stays
"unpinned" (DESIGN.md section 5)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from glomap_b200 import estimators as E, geometry as G, synthetic as S   # noqa: E402
from oracle import ba_oracle as B, gp_oracle as GPO, ra_oracle as RO      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ba_fixture():
    sc = S.make_scene(12, 150, mean_track_len=5, seed=101, pixel_sigma=0.5, model=S.SIMPLE_RADIAL, num_intrinsics=2)
    init = S.perturb_scene(sc, seed=102)
    mask = E.first_frame_mask(sc.C)
    out = {}
    for name, opts in (("const_intr", B.BAOptions()), ("opt_intr", B.BAOptions(optimize_intrinsics=True))):
        x, summ = B.solve_ba(init.quat, init.trans, init.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr,
                             sc.intr_model, init.intr_params, opts, mask)
        out.update({f"{name}_quat": x["quat"], f"{name}_trans": x["trans"], f"{name}_points": x["points"],
                    f"{name}_intr": x["intr"], f"{name}_cost": np.array([summ.initial_cost, summ.final_cost]),
                    f"{name}_iterations": np.array([summ.iterations])})
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), quat=init.quat, trans=init.trans, points=init.points,
                        pt_obs_begin=sc.pt_obs_begin, obs_cam=sc.obs_cam, obs_xy=sc.obs_xy, cam_intr=sc.cam_intr,
                        intr_model=sc.intr_model, intr_params=init.intr_params, cam_const_mask=mask, **out)


def gp_fixture():
    sc = S.make_scene(14, 200, mean_track_len=5, seed=111, pixel_sigma=0.3)
    bear = S.bearings_from_scene(sc)
    t_obs = E.world_bearings(sc.quat, bear, sc.obs_cam)
    rng = np.random.default_rng(112)
    cen0 = G.centers_from_pose(G.quat_xyzw_to_rotmat(sc.quat), sc.trans) + rng.normal(size=(sc.C, 3))
    pts0 = sc.points + rng.normal(size=sc.points.shape)
    cal = (np.arange(sc.C) % 4 != 0).astype(np.uint8)
    x, summ = GPO.solve_gp(cen0, pts0, sc.pt_obs_begin, sc.obs_cam, t_obs, cal, GPO.GPOptions())
    np.savez_compressed(os.path.join(HERE, "gp_small.npz"), centers0=cen0, points0=pts0, pt_obs_begin=sc.pt_obs_begin,
                        obs_cam=sc.obs_cam, obs_dir=t_obs, cam_calibrated=cal, centers=x["centers"], points=x["points"],
                        scales=x["scales"], cost=np.array([summ.initial_cost, summ.final_cost]),
                        iterations=np.array([summ.iterations]))


def ra_fixture():
    vg = S.make_random_view_graph(40, 6.0, seed=121, noise_deg=2.0, outlier_ratio=0.1)
    R0 = E.initialize_from_maximum_spanning_tree(vg, None)
    th, info = RO.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
    np.savez_compressed(os.path.join(HERE, "ra_small.npz"), n=np.array([vg.n_images]), ei=vg.ei, ej=vg.ej, R_rel=vg.R_rel,
                        theta0=G.so3_log(R0), theta=th,
                        iterations=np.array([info["l1_iterations"], info["irls_iterations"], info["admm_iterations"]]))


if __name__ == "__main__":
    ba_fixture(); gp_fixture(); ra_fixture()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
