#!/bin/bash
# 2-GPU box: multi-GPU parity with the split all-reduce, bench N = 2 with / without it; RA fused vs unfused on GPU 0
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -15 > $O/r2_t12.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 \
  bench.py --gpus 2 > $O/r2_bench_n2_split.json 2> $O/r2_bench_n2_split.err
B200SFM_SPLIT_AR=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 \
  bench.py --gpus 2 > $O/r2_bench_n2_nosplit.json 2> $O/r2_bench_n2_nosplit.err
python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_fused2.log 2>&1
tail -4 $O/r2_t12.log; cut -c1-200 $O/r2_bench_n2_split.json; cut -c1-200 $O/r2_bench_n2_nosplit.json; tail -1 $O/r2_ra5_fused2.log | cut -c1-500
