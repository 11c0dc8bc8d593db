#!/bin/bash
# GPU job of round 2 (run under gpurun from the repo root): re-run of the tests adjusted after the first full run, ncu launch
# list + full captures of the four hot BA kernels, -D tunable sweep, secondary RA / GP measurements, the bench line.
export PYTHONPATH=$PWD
O=gpurun_out
python -m pytest tests/test_ba_gpu.py::test_many_per_image_cameras_at_bench_tolerance tests/test_config2_gpu.py \
  tests/test_rig_gpu.py::test_rotation_averaging_with_unknown_cam_from_rig -q 2>&1 | tail -30 > $O/r2_t4.log
python -c "import sys; sys.path.insert(0,'profiles'); import sweep_worker as w; w.gen()" > $O/r2_gen.log 2>&1
export SWEEP_REPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 500 --csv --log-file $O/r2_launches.csv \
  python profiles/sweep_worker.py --worker > $O/r2_launches.out 2>&1
for k in ba3_linearize_points ba2_linearize_cams ba3_pass_a ba2_pass_b; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o $O/r2_prof_$k \
    python profiles/sweep_worker.py --worker > $O/r2_prof_$k.out 2>&1
done
export SWEEP_REPS=3
SWEEP_CARVES=75 python profiles/sweep_worker.py > $O/r2_sweep.log 2>&1
python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_2lvl.log 2>&1
B200SFM_RA_2LVL=0 python bench_secondary.py --what ra --neighbours 100 --pcg-tol 1e-6 > $O/r2_ra5_jacobi.log 2>&1
python bench_secondary.py --what gp > $O/r2_gp2b.log 2>&1
python bench.py > $O/r2_bench2.json 2> $O/r2_bench2.err
tail -3 $O/r2_t4.log
