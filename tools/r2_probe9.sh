#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "gram=0,heavy=2;gram=1;gram=1,warps=24;gram=1,warps=20" > $O/r2j_sweep_50k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 5000 --configs "gram=0;gram=2" > $O/r2j_sweep_5k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 20000 --max-len 12 --configs "gram=0;gram=2" > $O/r2j_sweep_20k.log 2>&1
python tools/sweep.py --mb 1024 --reps 7 --configs "gram=0;gram=2" > $O/r2j_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --alphabet "abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn" --configs "gram=0;gram=2" > $O/r2j_sweep_lowercase.log 2>&1
cat $O/r2j_sweep_*.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "4gram or layout or two_level or config5" > $O/r2j_gpu_tests.log 2>&1
grep -E "passed|failed" $O/r2j_gpu_tests.log
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2j_ncu_50k.csv python tools/sweep.py --mb 256 --reps 1 --lits 50000 --max-len 16 --configs "gram=1" > $O/r2j_ncu_50k.out 2>&1
