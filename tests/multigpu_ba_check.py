"""Run under torchrun on >= 2 GPUs (not collected by pytest): the sharded BA
must reproduce the single-GPU BA.  Usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_ba_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glomap_b200 import dist as D, estimators as E, geometry as G, synthetic as S  # noqa: E402


def main():
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local, rank, world, D.broadcast_nccl_id(E.Context.nccl_unique_id, rank, world))
    full = S.make_scene(60, 6000, 7, seed=5, pixel_sigma=0.5, chunk=500)
    init = S.perturb_scene(full, chunk=500)
    mask = E.first_frame_mask(full.C)
    opts = E.BundleAdjusterOptions(optimize_intrinsics=False)
    opts.solver_options.pcg_rel_tolerance = 1e-12
    opts.solver_options.pcg_max_iterations = 2000
    # (1) trajectory parity: a FIXED number of LM iterations (no termination test that could flip on a borderline
    #     comparison with the summation order) must give the same cost after every one of them;
    # (2) the natural run (default tolerances) must stop after the same number of iterations.
    traj = []
    for k in (1, 2, 3, 4, 6):
        o2 = E.BundleAdjusterOptions(optimize_intrinsics=False, fixed_num_iterations=k)
        o2.solver_options.pcg_rel_tolerance = 1e-12
        o2.solver_options.pcg_max_iterations = 2000
        sh_k, _ = D.shard_scene(init, rank, world, chunk=500)
        bk = E.BundleAdjuster(o2, ctx)
        assert bk.Solve(sh_k, mask)
        traj.append(bk.summary.final_cost)
    shard, (a, b) = D.shard_scene(init, rank, world, chunk=500)
    ba = E.BundleAdjuster(opts, ctx)
    assert ba.Solve(shard, mask)
    st = ba.summary
    ok = True
    if rank == 0:
        one = E.Context(local)
        traj1 = []
        for k in (1, 2, 3, 4, 6):
            o2 = E.BundleAdjusterOptions(optimize_intrinsics=False, fixed_num_iterations=k)
            o2.solver_options.pcg_rel_tolerance = 1e-12
            o2.solver_options.pcg_max_iterations = 2000
            r_k = init.copy()
            bk = E.BundleAdjuster(o2, one)
            assert bk.Solve(r_k, mask)
            traj1.append(bk.summary.final_cost)
        print("cost after k LM iterations, multi vs single:", list(zip(traj, traj1)))
        ok_traj = all(abs(a_ - b_) <= 1e-9 * b_ for a_, b_ in zip(traj, traj1))
        ref = init.copy()
        ba1 = E.BundleAdjuster(opts, one)
        assert ba1.Solve(ref, mask)
        s1 = ba1.summary
        # BA leaves the global scale free (only the first frame is fixed): compare after the
        # Sim3 alignment on projection centres, as the reference's tests do (global_mapper_test.cc:27-33)
        rot, cen, (sc_, R_, t_) = G.compare_reconstructions(G.quat_xyzw_to_rotmat(shard.quat), shard.trans,
                                                            G.quat_xyzw_to_rotmat(ref.quat), ref.trans)
        pts_al = (sc_ * (R_ @ shard.points.T)).T + t_
        dp = np.abs(pts_al - ref.points[a:b]).max()
        print(f"multi-GPU({world}) vs single: its {st.iterations}/{s1.iterations} cost {st.final_cost:.12e}/{s1.final_cost:.12e} "
              f"after Sim3: rot {rot:.2e} deg centre {cen:.2e} points {dp:.2e} (scale {sc_:.9f})")
        ok = (ok_traj and st.iterations == s1.iterations and abs(st.final_cost - s1.final_cost) <= 1e-9 * s1.final_cost and
              rot < 1e-6 and cen < 1e-6 and dp < 1e-5)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    if flag.item() < 1:
        raise SystemExit("multi-GPU parity FAILED")
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
