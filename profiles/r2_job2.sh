#!/bin/bash
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_config2_gpu.py tests/test_rig_gpu.py tests/test_ra_gpu.py tests/test_gp_gpu.py tests/test_ba_gpu.py -q 2>&1 | tail -30 > $O/r2_t5.log
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
SWEEP_REPS=3 B200SFM_CARVEOUT=75 python profiles/sweep_worker.py --worker > $O/r2_sweep2.log 2>&1
python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_b.log 2>&1
python bench_secondary.py --what gp > $O/r2_gp2c.log 2>&1
python bench.py > $O/r2_bench3.json 2> $O/r2_bench3.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench3_ref.json 2> $O/r2_bench3_ref.err
tail -4 $O/r2_t5.log; cat $O/r2_sweep2.log | cut -c1-300
