"""b200sfm_cli (the C++ host path: text / flat-binary parsing -> unordered_map world -> shim classes) linked against the
recording test double of the C ABI: checks, without a GPU, what the CLI hands to the solvers -- the relpose text format
of `glomap rotation_averager` (docs/rotation_averager.md:43-69, io/pose_io.cc:8-89), the maximum-spanning-tree
initialisation, and the flat problem reader of the `ba` / `gp` subcommands."""
import os
import subprocess

import numpy as np

from glomap_b200 import estimators as E, geometry as G, synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    lib, cli = tmp_path / "libb200sfm.so", tmp_path / "b200sfm_cli"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O1", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", str(lib),
                    os.path.join(ROOT, "tests", "shim_mock", "mock_b200sfm.c")], check=True, capture_output=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(cli), os.path.join(ROOT, "glomap_b200", "host", "b200sfm_cli.cc"),
                    str(lib), "-Wl,-rpath," + str(tmp_path)], check=True, capture_output=True)
    return cli


def _calls(dump):
    calls, cur = [], None
    for line in open(dump):
        f = line.split()
        if f[0] == "call":
            cur = {"_name": f[1]}
            calls.append(cur)
        else:
            cur[f[0]] = np.array([float(x) for x in f[2:]])
    return calls


def test_cli_rotation_averager_parses_relpose_and_initialises_from_the_tree(tmp_path):
    cli = _build(tmp_path)
    vg = S.make_random_view_graph(30, 5.0, seed=4, noise_deg=0.0)
    rel, out, dump = str(tmp_path / "relpose.txt"), str(tmp_path / "rot.txt"), str(tmp_path / "dump.txt")
    S.write_relpose_file(rel, vg)
    r = subprocess.run([str(cli), "rotation_averager", "--relpose_path", rel, "--output_path", out, "--mst_init", "1"],
                       env=dict(os.environ, MOCK_DUMP=dump), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    (c,) = [c for c in _calls(dump) if c["_name"] == "ra_solve"]
    vg2, names = S.read_relpose_file(rel)
    assert c["dims"].tolist() == [vg2.n_images, vg2.E]
    # same edge set (the CLI orders pairs by colmap pair id), same relative rotations up to the 6+ digits of the file
    got = {(int(a), int(b)): R for a, b, R in zip(c["ei"], c["ej"], c["R_rel"].reshape(-1, 3, 3))}
    for a, b, R in zip(vg2.ei, vg2.ej, vg2.R_rel):
        key = (int(a), int(b)) if (int(a), int(b)) in got else (int(b), int(a))
        Rg = got[key] if key == (int(a), int(b)) else got[key].T
        assert np.abs(Rg - R).max() < 1e-9
    # noise-free graph: the spanning-tree initialisation handed to the solver is already the solution (up to gauge)
    R0 = G.so3_exp(c["theta"].reshape(-1, 3))
    rel_err = np.abs(R0[vg2.ej] @ np.swapaxes(R0[vg2.ei], -1, -2) - vg2.R_rel).max()
    assert rel_err < 1e-5
    # the mock leaves theta untouched: the output file holds the initialisation, one line per image
    lines = open(out).read().strip().splitlines()
    assert len(lines) == vg2.n_images and all(len(l.split()) == 5 for l in lines)


def test_cli_ba_reads_the_flat_problem(tmp_path):
    cli = _build(tmp_path)
    sc = S.make_scene(6, 40, mean_track_len=4, seed=7, model=S.SIMPLE_RADIAL, num_intrinsics=2)
    flat, outp, dump = str(tmp_path / "p.bin"), str(tmp_path / "o.bin"), str(tmp_path / "dump.txt")
    S.write_flat_problem(flat, sc)
    r = subprocess.run([str(cli), "ba", "--problem", flat, "--output", outp, "--fix_rotations", "1"],
                       env=dict(os.environ, MOCK_DUMP=dump), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c, c2 = [c for c in _calls(dump) if c["_name"] == "ba_solve"]
    assert c["dims"].tolist() == [sc.C, sc.P, sc.N, 2]
    # --fix_rotations 1 = the mapper's staged solve: rotations constant first, then free, on the SAME BundleAdjuster
    # through GetOptions() (controllers/global_mapper.cc:204-221)
    assert c["flags"].tolist()[0] == 0 and c2["flags"].tolist()[0] == 1
    assert np.array_equal(c["pt_obs_begin"], sc.pt_obs_begin) and np.array_equal(c["obs_cam"], sc.obs_cam)
    assert np.array_equal(c["obs_xy"].reshape(-1, 2), sc.obs_xy) and np.array_equal(c["cam_intr"], sc.cam_intr)
    assert np.abs(c["quat"].reshape(-1, 4) - sc.quat).max() < 1e-15 and np.array_equal(c["points"].reshape(-1, 3), sc.points)
    back = S.read_flat_problem(outp)                                  # unchanged state written back
    assert np.array_equal(back.points, sc.points) and np.array_equal(back.obs_cam, sc.obs_cam)
