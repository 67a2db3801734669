#!/bin/bash
# Round 2, third GPU call: restructured class-pair loop; byte-table kernels wide+split; full GPU suite.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -x -q -m gpu ) > $O/r2c_gpu_tests.log 2>&1
tail -3 $O/r2c_gpu_tests.log
python tools/sweep.py --mb 512 --reps 9 --configs \
"first_stage=3;first_stage=3,warps=24;first_stage=3,pf_dist=4;first_stage=3,pf_dist=12;first_stage=1" \
  > $O/r2c_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 48 --configs "queue=2;wide=1,split=1;wide=1,split=1,warps=28;wide=1,split=1,warps=32" > $O/r2c_sweep_teddy48.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 8 --configs "queue=2;wide=1,split=1,warps=28" > $O/r2c_sweep_teddy8.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 1 --min-len 6 --max-len 6 --configs "queue=2;wide=1,split=1,warps=28" > $O/r2c_sweep_noodle.log 2>&1
cat $O/r2c_sweep_*.log
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2c_ncu_pair.csv python tools/sweep.py --mb 512 --reps 1 --configs "first_stage=3" > $O/r2c_ncu_pair.out 2>&1
python bench.py --steps 10 --warmup 3 > $O/r2c_bench.json 2> $O/r2c_bench.err
tail -c 1500 $O/r2c_bench.json
