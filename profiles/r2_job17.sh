#!/bin/bash
# 8-GPU box: peer-memory all-reduce at 8 ranks -- bench with it (default from 4 ranks) and with NCCL only, 8-rank parity
export PYTHONPATH=$PWD
O=gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 \
  bench.py --gpus 8 > $O/r2_bench_n8_p2p.json 2> $O/r2_bench_n8_p2p.err
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29752 tests/multigpu_ba_check.py > $O/r2_mg8p_ba.log 2>&1
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29753 tests/multigpu_gp_ra_check.py > $O/r2_mg8p_gpra.log 2>&1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29754 \
  bench.py --gpus 4 --e2e-steps 3 > $O/r2_bench_n4_p2p.json 2> $O/r2_bench_n4_p2p.err
cut -c1-230 $O/r2_bench_n8_p2p.json; cut -c1-230 $O/r2_bench_n4_p2p.json; grep -h "multi" $O/r2_mg8p_ba.log $O/r2_mg8p_gpra.log | cut -c1-200; tail -2 $O/r2_bench_n8_p2p.err | cut -c1-300
