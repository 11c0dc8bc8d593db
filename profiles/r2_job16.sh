#!/bin/bash
# 2-GPU box: peer-memory all-reduce -- parity tests, bench N = 2 with it and with NCCL only
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -8 > $O/r2_t16.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
  bench.py --gpus 2 > $O/r2_bench_n2_p2p.json 2> $O/r2_bench_n2_p2p.err
B200SFM_P2P_AR=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 \
  bench.py --gpus 2 > $O/r2_bench_n2_nccl.json 2> $O/r2_bench_n2_nccl.err
tail -3 $O/r2_t16.log; cut -c1-220 $O/r2_bench_n2_p2p.json; cut -c1-220 $O/r2_bench_n2_nccl.json; tail -3 $O/r2_bench_n2_p2p.err | cut -c1-300
