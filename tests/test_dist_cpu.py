"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: rank-
independent chunked scene generation, point sharding, id broadcast."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from glomap_b200 import dist as D, synthetic as S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, P, chunk = 40, 4000, 500
    a, b = D.shard_range(P, chunk, rank, world)
    sc = S.make_scene(C, P, 6, seed=1, chunk=chunk, point_range=(a, b))
    init = S.perturb_scene(sc, chunk=chunk, point_offset=a)
    ident = D.broadcast_nccl_id(lambda: b"x" * 128, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(a=a, b=b, N=sc.N, pts=sc.points, cam=sc.obs_cam, init=init.points,
                                          quat=init.quat, ident=ident))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_generation_equals_full_scene():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = S.make_scene(40, 4000, 6, seed=1, chunk=500)
    finit = S.perturb_scene(full, chunk=500)
    assert shards[0]["a"] == 0 and shards[-1]["b"] == 4000 and shards[0]["b"] == shards[1]["a"]
    assert sum(s["N"] for s in shards) == full.N
    assert np.array_equal(np.concatenate([s["pts"] for s in shards]), full.points)
    assert np.array_equal(np.concatenate([s["cam"] for s in shards]), full.obs_cam)
    assert np.array_equal(np.concatenate([s["init"] for s in shards]), finit.points)
    assert all(np.array_equal(s["quat"], finit.quat) for s in shards)       # cameras replicated
    assert all(s["ident"] == b"x" * 128 for s in shards)


def test_shard_scene_partitions_points_and_observations():
    sc = S.make_scene(20, 1000, 6, seed=2)
    parts = [D.shard_scene(sc, r, 3, chunk=64) for r in range(3)]
    assert sum(p[0].P for p in parts) == sc.P and sum(p[0].N for p in parts) == sc.N
    for sh, (a, b) in parts:
        assert sh.pt_obs_begin[0] == 0 and sh.pt_obs_begin[-1] == sh.N
        assert np.array_equal(sh.points, sc.points[a:b])
        assert np.array_equal(np.diff(sh.pt_obs_begin), np.diff(sc.pt_obs_begin[a:b + 1]))


def test_shard_rig_scene_keeps_frames_and_sensors_replicated():
    """Known rigs shard like plain scenes: tracks (with their frame / sensor indices) are split, the frame
    poses and the sensor table are replicated; the shard costs add up to the full cost (oracle)."""
    from oracle import ba_oracle as B
    rs = S.make_rig_scene(8, 3, 300, seed=3, pixel_sigma=0.5)
    parts = [D.shard_scene(rs, r, 2, chunk=32) for r in range(2)]
    assert sum(p[0].P for p in parts) == rs.P and sum(p[0].N for p in parts) == rs.N
    total = 0.0
    for sh, (a, b) in parts:
        assert sh.F == rs.F and sh.S == rs.S and np.array_equal(sh.sensor_quat, rs.sensor_quat)
        o0 = int(rs.pt_obs_begin[a])
        assert np.array_equal(sh.obs_sensor, rs.obs_sensor[o0:o0 + sh.N])
        assert np.array_equal(sh.obs_frame, rs.obs_frame[o0:o0 + sh.N])
        p = B.BAProblem(sh.quat, sh.trans, sh.points, sh.pt_obs_begin, sh.obs_frame, sh.obs_xy, np.zeros(sh.F, np.int32),
                        sh.intr_model, sh.intr_params, B.BAOptions(), rig=sh.rig_dict())
        total += p.evaluate(p.x0, False)[0]
    full = B.BAProblem(rs.quat, rs.trans, rs.points, rs.pt_obs_begin, rs.obs_frame, rs.obs_xy, np.zeros(rs.F, np.int32),
                       rs.intr_model, rs.intr_params, B.BAOptions(), rig=rs.rig_dict())
    assert abs(total - full.evaluate(full.x0, False)[0]) <= 1e-12 * total
