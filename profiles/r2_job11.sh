#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_ra_gpu.py tests/test_rig_gpu.py tests/test_mapper_gpu.py tests/test_golden_gpu.py -q 2>&1 | tail -12 > $O/r2_t11.log
python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1
python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_fused.log 2>&1
B200SFM_RA_FUSED=0 python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_unfused.log 2>&1
python bench.py --workload config5 > $O/r2_bench_cfg5.json 2> $O/r2_bench_cfg5.err
python bench.py > $O/r2_bench6.json 2> $O/r2_bench6.err
tail -4 $O/r2_t11.log; tail -3 $O/r2_smoke.log | cut -c1-300; tail -1 $O/r2_ra5_fused.log | cut -c1-700; tail -1 $O/r2_ra5_unfused.log | cut -c1-500
