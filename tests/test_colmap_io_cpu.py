"""COLMAP sparse-model binary I/O and the flat-scene converters (reference:
glomap/io/colmap_io.cc:8-58, colmap_converter.cc:22-133).  Known-answer bytes for the three files are written by
hand from the published layout; the converters must round-trip a synthetic scene exactly."""
import struct

import numpy as np

from glomap_b200 import colmap_io as CIO, synthetic as S


def test_known_answer_bytes(tmp_path):
    cam = struct.pack("<Q", 1) + struct.pack("<IiQQ", 7, 2, 640, 480) + struct.pack("<4d", 500.0, 320.0, 240.0, 0.01)
    img = (struct.pack("<Q", 1) + struct.pack("<I", 3) + struct.pack("<4d", 1.0, 0.0, 0.0, 0.0) + struct.pack("<3d", 0.1, 0.2, 0.3)
           + struct.pack("<I", 7) + b"a/b.jpg\0" + struct.pack("<Q", 2) + struct.pack("<ddQ", 10.5, 20.5, 42)
           + struct.pack("<ddQ", 1.0, 2.0, 2**64 - 1))
    pts = (struct.pack("<Q", 1) + struct.pack("<Q", 42) + struct.pack("<3d", 1.0, 2.0, 3.0) + bytes([9, 8, 7])
           + struct.pack("<dQ", 0.25, 1) + struct.pack("<II", 3, 0))
    for name, data in (("cameras.bin", cam), ("images.bin", img), ("points3D.bin", pts)):
        (tmp_path / name).write_bytes(data)
    cameras, images, points = CIO.read_model(str(tmp_path))
    assert cameras[7].model_id == 2 and cameras[7].width == 640 and np.allclose(cameras[7].params, [500, 320, 240, 0.01])
    im = images[3]
    assert im.name == "a/b.jpg" and im.camera_id == 7 and np.allclose(im.tvec, [0.1, 0.2, 0.3])
    assert im.xy.shape == (2, 2) and im.point3D_ids[0] == 42 and im.point3D_ids[1] == CIO.INVALID_POINT3D
    p = points[42]
    assert np.allclose(p.xyz, [1, 2, 3]) and list(p.rgb) == [9, 8, 7] and p.error == 0.25 and p.image_ids[0] == 3
    # writing the parsed model back reproduces the bytes
    out = tmp_path / "out"
    CIO.write_model(str(out), cameras, images, points)
    for name, data in (("cameras.bin", cam), ("images.bin", img), ("points3D.bin", pts)):
        assert (out / name).read_bytes() == data


def test_scene_round_trip(tmp_path):
    sc = S.make_scene(12, 200, mean_track_len=5, seed=3, pixel_sigma=0.5, model=S.RADIAL, num_intrinsics=2)
    cameras, images, points = CIO.model_from_scene(sc)
    CIO.write_model(str(tmp_path), cameras, images, points)
    sc2, index = CIO.scene_from_model(*CIO.read_model(str(tmp_path)))
    assert sc2.C == sc.C and sc2.P == sc.P and sc2.N == sc.N
    qn = sc.quat / np.linalg.norm(sc.quat, axis=1, keepdims=True)
    assert np.abs(sc2.quat - qn).max() < 1e-15 and np.array_equal(sc2.trans, sc.trans)
    assert np.array_equal(sc2.points, sc.points) and np.array_equal(sc2.pt_obs_begin, sc.pt_obs_begin)
    assert np.array_equal(sc2.obs_cam, sc.obs_cam) and np.array_equal(sc2.obs_xy, sc.obs_xy)
    assert np.array_equal(sc2.cam_intr, sc.cam_intr) and np.array_equal(sc2.intr_model, sc.intr_model)
    assert np.array_equal(sc2.intr_params[:, :5], sc.intr_params[:, :5])
    # every observed feature carries its point id, the error is the mean reprojection error (0.5 px noise)
    n_set = sum(int((im.point3D_ids != CIO.INVALID_POINT3D).sum()) for im in images.values())
    assert n_set == sc.N
    errs = np.array([p.error for p in points.values()])
    assert 0.2 < errs.mean() < 1.5
    # write the (here unchanged) state back through the index: identical model
    cameras3, images3, points3 = CIO.model_from_scene(sc2, index)
    out = tmp_path / "again"
    CIO.write_model(str(out), cameras3, images3, points3)
    for name in ("cameras.bin", "images.bin"):
        assert (out / name).read_bytes() == (tmp_path / name).read_bytes()
    # points: identical up to the last bits of the recomputed error (the quaternions were re-normalised on the way)
    pa, pb = CIO.read_points3D(str(out / "points3D.bin")), CIO.read_points3D(str(tmp_path / "points3D.bin"))
    assert pa.keys() == pb.keys()
    for k in pa:
        assert np.array_equal(pa[k].xyz, pb[k].xyz) and np.array_equal(pa[k].image_ids, pb[k].image_ids)
        assert np.array_equal(pa[k].point2D_idxs, pb[k].point2D_idxs) and abs(pa[k].error - pb[k].error) < 1e-9


def test_short_tracks_are_dropped_and_missing_images_skipped():
    sc = S.make_scene(8, 60, mean_track_len=4, seed=5)
    # make track 0 a single observation: ConvertGlomapToColmap drops points with < 2 supports
    lens = np.diff(sc.pt_obs_begin)
    keep = np.ones(sc.N, bool)
    keep[1:lens[0]] = False
    sc.obs_cam, sc.obs_xy = sc.obs_cam[keep], sc.obs_xy[keep]
    lens[0] = 1
    sc.pt_obs_begin = np.concatenate([[0], np.cumsum(lens)])
    cameras, images, points = CIO.model_from_scene(sc)
    assert len(points) == sc.P - 1 and 1 not in points
    # an image missing from the model: its track elements are skipped on read (bundle_adjustment.cc:125)
    del images[3]
    sc2, _ = CIO.scene_from_model(cameras, images, points)
    assert sc2.C == sc.C - 1 and sc2.N == int((sc.obs_cam[lens[0]:] != 2).sum())
