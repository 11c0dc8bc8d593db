#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_ba_gpu.py tests/test_config2_gpu.py tests/test_rig_gpu.py tests/test_gp_gpu.py -q -x 2>&1 | tail -30 > $O/r2_t8.log
timeout 900 python profiles/r2_gp_diag.py > $O/r2_gp_diag2.log 2>&1
python bench.py --optimize-intrinsics 1 --no-parity --no-cpu-baseline > $O/r2_bench_intr.json 2> $O/r2_bench_intr.err
B200SFM_KFAST=0 python bench.py --optimize-intrinsics 1 --no-parity --no-cpu-baseline > $O/r2_bench_intr_mf.json 2> $O/r2_bench_intr_mf.err
tail -12 $O/r2_t8.log | cut -c1-300; grep "tol\|scales:" $O/r2_gp_diag2.log | cut -c1-250
