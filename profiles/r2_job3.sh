#!/bin/bash
# 2-GPU box: multi-GPU parity inside pytest, 2-GPU bench line, then single-GPU variant sweep and the GP secondary bench
export PYTHONPATH=$PWD
O=gpurun_out
nvidia-smi -L > $O/r2_gpus.log
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -40 > $O/r2_t6.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
B200SFM_PCG_DEPTH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_n2_d1.json 2> $O/r2_bench_n2_d1.err
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
SWEEP_REPS=3 SWEEP_CARVES=75 python profiles/sweep_worker.py > $O/r2_sweep3.log 2>&1
python bench_secondary.py --what gp > $O/r2_gp2d.log 2>&1
tail -5 $O/r2_t6.log; cat $O/r2_sweep3.log | cut -c1-260; tail -c 300 $O/r2_bench_n2.err
