#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -x -q -m gpu ) > $O/r2d_gpu_tests.log 2>&1
grep -E "passed|failed" $O/r2d_gpu_tests.log
python tools/sweep.py --mb 1024 --reps 9 --configs "first_stage=3;first_stage=3,pf_dist=4;first_stage=3,pf_dist=0;first_stage=3,warps=24;first_stage=3,warps=26" > $O/r2d_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "big_set=0;big_set=1;big_set_classes=1;big_set_classes=2;big_set_classes=8;big_set_classes=4,warps=20;first_stage=1,wide=0,split=0" > $O/r2d_sweep_50k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 5000 --configs "big_set=0;big_set=1;big_set_classes=8" > $O/r2d_sweep_5k.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 48 --configs "wide=1,split=1;wide=0,split=0" > $O/r2d_sweep_teddy48.log 2>&1
cat $O/r2d_sweep_*.log
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2d_ncu_pair.csv python tools/sweep.py --mb 512 --reps 1 --configs "first_stage=3" > $O/r2d_ncu_pair.out 2>&1
ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2d_ncu_50k.csv python tools/sweep.py --mb 256 --reps 1 --lits 50000 --max-len 16 --configs "big_set=1" > $O/r2d_ncu_50k.out 2>&1
