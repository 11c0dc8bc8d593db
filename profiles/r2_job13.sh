#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 > $O/r2_t13.log
grep -h "config-2 pipeline" $O/r2_t13.log; tail -12 $O/r2_t13.log | cut -c1-300
