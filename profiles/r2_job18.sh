#!/bin/bash
# final single-GPU validation of round 2: full GPU test suite, the default bench line, linearisation variant sweep
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > $O/r2_t18_full.log
grep -h "config-2 pipeline" $O/r2_t18_full.log | cut -c1-300
tail -6 $O/r2_t18_full.log | cut -c1-300
python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err
cut -c1-300 $O/r2_bench_final.json
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
SWEEP_REPS=3 SWEEP_CARVES=max timeout 150 python profiles/sweep_worker.py > $O/r2_sweep5.log 2>&1
cut -c1-260 $O/r2_sweep5.log
