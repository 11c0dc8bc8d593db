#!/bin/bash
# ncu: launch list of one warm solve + full captures of the two linearisation kernels of the current build
export PYTHONPATH=$PWD
O=gpurun_out
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
export SWEEP_REPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file $O/r2_launches.csv \
  python profiles/sweep_worker.py --worker > $O/r2_launches.out 2>&1
for k in ba3_linearize_points ba2_linearize_cams; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $O/r2b_prof_$k \
    python profiles/sweep_worker.py --worker > $O/r2b_prof_$k.out 2>&1
done
wc -l $O/r2_launches.csv
