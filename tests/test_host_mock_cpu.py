"""Python host classes and the mapper driver against the recording test double of the C ABI
(tests/shim_mock/mock_b200sfm.c, loaded through B200SFM_LIB in a subprocess): no GPU, no numerics -- what is
checked is the host logic: stage order and option mutations of the mapper (controllers/global_mapper.cc:84-276),
the thresholds handed to the filters, compaction of the observations between stages, the rig paths of
BundleAdjuster / GlobalPositioner and the gravity preparation of RotationEstimator."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mock(tmp_path):
    lib = tmp_path / "libb200sfm_mock.so"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O1", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", str(lib),
                    os.path.join(ROOT, "tests", "shim_mock", "mock_b200sfm.c")], check=True, capture_output=True)
    return lib


def _run(tmp_path, script, **env):
    lib, dump = _mock(tmp_path), tmp_path / "dump.txt"
    if dump.exists():
        dump.unlink()
    e = dict(os.environ, B200SFM_LIB=str(lib), MOCK_DUMP=str(dump), PYTHONPATH=ROOT, **env)
    r = subprocess.run([sys.executable, "-c", script], env=e, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    calls, cur = [], None
    for line in dump.read_text().splitlines():
        f = line.split()
        if f[0] == "call":
            cur = {"_name": f[1]}
            calls.append(cur)
        else:
            cur[f[0]] = np.array([float(x) for x in f[2:]])
    return json.loads(r.stdout.strip().splitlines()[-1]), calls


MAPPER = """
import json, numpy as np
from glomap_b200 import mapper as M, synthetic as S
sc = S.make_scene(10, 200, mean_track_len=5, seed=3)
vg = S.view_graph_from_scene(sc, min_shared=5)
start = sc.copy()
start.trans[:] = 0; start.points[:] = 0
opts = M.GlobalMapperOptions()
m = M.GlobalMapper(opts)
ok, out = m.Solve(vg, start)
print(json.dumps(dict(ok=bool(ok), N0=int(sc.N), N=int(out.N), P=int(out.P), log=m.log)))
"""


def test_mapper_stage_order_without_filtering(tmp_path):
    res, calls = _run(tmp_path, MAPPER)
    names = [c["_name"] for c in calls]
    assert res["ok"] and res["N"] == res["N0"]
    # 3. rotation averaging twice; 5. positioning; three filters; 6. BA: rotations fixed, then free; one reprojection
    # filter per tightening level (nothing is filtered -> the inner loop walks ite to the limit and the loop stops),
    # then the two final filters
    assert names[:2] == ["ra_solve", "ra_solve"] and names[2] == "gp_solve"
    filt = [(c["_name"], float(c["threshold"][0])) for c in calls if c["_name"].startswith("filter")]
    assert filt[:3] == [("filter_angle", 1.0), ("filter_triangulation_angle", 1.0), ("filter_reprojection_normalized", 0.1)]
    ba = [c for c in calls if c["_name"] == "ba_solve"]
    assert len(ba) == 2 and ba[0]["flags"].tolist() == [0, 1, 0] and ba[1]["flags"].tolist() == [1, 1, 0]
    after = filt[3:]
    assert [round(t, 12) for n, t in after if n == "filter_reprojection_normalized"] == [0.03, 0.02, 0.01, 0.01]
    assert after[-1] == ("filter_triangulation_angle", 1.0)
    assert any("fewer than 0.1%" in line for line in res["log"])


def test_mapper_loops_while_tracks_are_filtered(tmp_path):
    res, calls = _run(tmp_path, MAPPER, MOCK_DROP_EVERY="50")
    ba = [c for c in calls if c["_name"] == "ba_solve"]
    assert len(ba) == 6                                           # 3 outer iterations x (rotations fixed, free)
    assert [c["flags"].tolist()[0] for c in ba] == [0, 1, 0, 1, 0, 1]
    # every filter pass compacts the observations: the next problem is created with fewer of them
    nobs = [int(c["nobs"][0]) for c in calls if c["_name"].startswith("filter_reprojection")]
    assert all(b < a for a, b in zip(nobs, nobs[1:])) and res["N"] < nobs[-1] <= res["N0"]
    thr = [round(float(c["threshold"][0]), 12) for c in calls if c["_name"] == "filter_reprojection_normalized"]
    assert thr == [0.1, 0.03, 0.02, 0.01, 0.01]                   # 10x after GP; max(3 - ite, 1) x thr; final


RIGS = """
import json, numpy as np
from glomap_b200 import estimators as E, synthetic as S, geometry as G
rs = S.make_rig_scene(6, 3, 60, seed=2)
ba = E.BundleAdjuster(E.BundleAdjusterOptions())
ok = ba.Solve(rs.copy())
bear = S.bearings_from_scene(rs.images_scene())
prob = E.PositioningProblem(rs.quat, rs.pt_obs_begin, rs.obs_frame, bear, obs_sensor=rs.obs_sensor, sensor_quat=rs.sensor_quat,
                            sensor_trans=rs.sensor_trans, sensor_calibrated=np.array([1, 0, 1], np.uint8))
ok2 = E.GlobalPositioner(E.GlobalPositionerOptions()).Solve(prob)
t_obs, t_rig = E.rig_world_terms(rs.quat, rs.sensor_quat, rs.sensor_trans, bear, rs.obs_frame, rs.obs_sensor)
# gravity preparation: two frames with a gravity prior
vg = S.make_random_view_graph(8, 4.0, seed=5)
g = np.full((8, 3), np.nan); g[2] = [0.1, 1.0, 0.05]; g[5] = [0.0, 1.0, 0.0]
o = E.RotationEstimatorOptions(use_gravity=True)
ok3, R = E.RotationEstimator(o).EstimateRotations(vg, None, 0, g)
print(json.dumps(dict(ok=bool(ok), ok2=bool(ok2), ok3=bool(ok3), N=int(rs.N), t_obs=t_obs.ravel().tolist(), t_rig=t_rig.ravel().tolist(),
                      obs_sensor=rs.obs_sensor.tolist(), obs_frame=rs.obs_frame.tolist())))
"""


def test_python_rig_and_gravity_paths(tmp_path):
    res, calls = _run(tmp_path, RIGS)
    assert res["ok"] and res["ok2"] and res["ok3"]
    names = [c["_name"] for c in calls]
    assert names[:3] == ["ba_problem_create_rig", "ba_problem_set_state", "ba_problem_solve"]
    c = calls[0]
    assert c["dims"].tolist()[:5] == [6, 60, res["N"], 3, 3]
    assert c["obs_sensor"].tolist() == res["obs_sensor"] and c["obs_frame"].tolist() == res["obs_frame"]
    assert c["mask"].tolist() == [3, 0, 0, 0, 0, 0]
    i = names.index("gp_problem_create")
    assert names[i:i + 3] == ["gp_problem_create", "gp_problem_set_rig_terms", "gp_problem_solve"]
    assert np.abs(calls[i]["obs_dir"] - np.array(res["t_obs"])).max() < 1e-15
    assert np.abs(calls[i + 1]["obs_offset"] - np.array(res["t_rig"])).max() < 1e-15
    cal = np.array([1, 0, 1])[np.array(res["obs_sensor"])]
    assert calls[i + 1]["obs_calibrated"].tolist() == cal.tolist()
    ra = calls[names.index("ra_solve_gravity")]
    assert ra["has_gravity"].tolist() == [0, 0, 1, 0, 0, 1, 0, 0]
    assert int(ra["dims"][2]) == 2                                # the first gravity frame is the fixed one (.cc:213-217)
    th = ra["theta"].reshape(-1, 3)
    assert np.abs(th[[2, 5]][:, [0, 2]]).max() == 0.0             # gravity frames are (0, phi, 0)
