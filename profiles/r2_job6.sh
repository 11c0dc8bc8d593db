#!/bin/bash
# sweep of the 256-bit-gather build (default: LC 4, variant: LC 3), full single-GPU test suite, bench line
export PYTHONPATH=$PWD
O=gpurun_out
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
SWEEP_REPS=3 SWEEP_CARVES=max python profiles/sweep_worker.py > $O/r2_sweep4.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2_t7.log
python bench.py > $O/r2_bench4.json 2> $O/r2_bench4.err
cat $O/r2_sweep4.log | cut -c1-400; tail -15 $O/r2_t7.log
