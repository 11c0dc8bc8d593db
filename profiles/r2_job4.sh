#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_ba_check.py > $O/r2_mg_ba.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/multigpu_gp_ra_check.py > $O/r2_mg_gpra.log 2>&1
grep -v "^\s*File\|^\s\s\s\s" $O/r2_mg_ba.log | tail -25 | cut -c1-300
grep -v "^\s*File\|^\s\s\s\s" $O/r2_mg_gpra.log | tail -25 | cut -c1-300
