#!/bin/bash
# 8-GPU box (expensive: 8x): bench N = 8 (default and split all-reduce), then the 8-rank parity checks
export PYTHONPATH=$PWD
O=gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 \
  bench.py --gpus 8 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
B200SFM_SPLIT_AR=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29709 \
  bench.py --gpus 8 > $O/r2_bench_n8_split.json 2> $O/r2_bench_n8_split.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29710 tests/multigpu_ba_check.py > $O/r2_mg8_ba.log 2>&1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tests/multigpu_gp_ra_check.py > $O/r2_mg8_gpra.log 2>&1
cut -c1-300 $O/r2_bench_n8.json; cut -c1-300 $O/r2_bench_n8_split.json; grep -h "multi" $O/r2_mg8_ba.log $O/r2_mg8_gpra.log | cut -c1-300
