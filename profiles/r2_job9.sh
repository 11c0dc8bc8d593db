#!/bin/bash
# 8-GPU box: multi-GPU parity tests (2 and 8 ranks), bench at N = 8 and 4
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -15 > $O/r2_t9.log
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
    bench.py --gpus $N > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
done
tail -5 $O/r2_t9.log; cat $O/r2_bench_n8.json | cut -c1-600
