#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2_t10.log
python bench.py --optimize-intrinsics 1 --no-parity --no-cpu-baseline > $O/r2_bench_intr2.json 2> $O/r2_bench_intr2.err
tail -8 $O/r2_t10.log | cut -c1-300; cut -c1-400 $O/r2_bench_intr2.json
