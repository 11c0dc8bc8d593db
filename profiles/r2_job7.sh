#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
timeout 900 python profiles/r2_gp_diag.py > $O/r2_gp_diag.log 2>&1
python bench.py --no-parity --no-cpu-baseline > $O/r2_bench5.json 2> $O/r2_bench5.err
cat $O/r2_gp_diag.log | cut -c1-300
