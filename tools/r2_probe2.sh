#!/bin/bash
# Round 2, second GPU call: the class-pair first stage (FK_PAIR32) against the hash kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -x -q -m gpu ) > $O/r2b_gpu_tests.log 2>&1
tail -3 $O/r2b_gpu_tests.log
python tools/sweep.py --mb 512 --reps 7 --configs \
"first_stage=1;first_stage=3;first_stage=3,warps=24;first_stage=3,warps=20;first_stage=3,warps=16;first_stage=3,pf_dist=4;first_stage=3,pf_dist=16;first_stage=3,pf_dist=0;first_stage=3,tile_bytes=4096;first_stage=3,prefilter=0;first_stage=1,wide=1,split=1,replicas=1;first_stage=1,wide=1,split=1,replicas=1,warps=28;first_stage=1,wide=1,split=1,replicas=2,warps=28" \
  > $O/r2b_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 1000 --min-len 6 --max-len 12 --configs "first_stage=1;first_stage=3" > $O/r2b_sweep_fdr_long.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 5000 --configs "first_stage=1;first_stage=3" > $O/r2b_sweep_5k.log 2>&1
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "first_stage=1;first_stage=3;first_stage=3,warps=16" > $O/r2b_sweep_50k.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --alphabet "abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn" --configs "first_stage=1;first_stage=3" > $O/r2b_sweep_lowercase.log 2>&1
cat $O/r2b_sweep_*.log
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
for cfg in "first_stage=3"; do
  tag=$(echo $cfg | tr ',=' '__')
  ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2b_ncu_$tag.csv python tools/sweep.py --mb 512 --reps 1 --configs "$cfg" > $O/r2b_ncu_$tag.out 2>&1
done
ls -la $O | tail -12
