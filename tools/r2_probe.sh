#!/bin/bash
# First GPU call of the next round: validate the opt-in kernel variants written at the
# end of round 1 (never run on a GPU), time them against the shipped kernels, and take the
# ncu sections that explain the differences.  Everything lands in gpurun_out/r2_*.
#   gpurun --timeout 600 -- 'bash tools/r2_probe.sh'
# Budget note: a gpurun call is charged about 3x the command's run time.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time HSB200_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q ) > $O/r2_exp_tests.log 2>&1
tail -3 $O/r2_exp_tests.log

python tools/sweep.py --mb 512 --reps 7 --configs \
"queue=2;wide=1;wide=1,split=1;wide=1,split=1,warps=32;wide=1,split=1,warps=24;wide=1,domain=12,replicas=8;wide=1,split=1,domain=12,replicas=8;wide=1,split=1,domain=12,replicas=8,warps=32;wide=1,split=1,domain=12,replicas=8,warps=24;wide=1,domain=12,replicas=4;wide=1,replicas=4;wide=1,warps=24;wide=1,pf_dist=4;wide=1,pf_dist=16;queue=1;domain=12,replicas=8,queue=1" \
  > $O/r2_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 48 --configs "queue=2;queue=0;wide=1;wide=1,split=1;wide=1,warps=24;wide=1,split=1,warps=24;wide=1,split=1,warps=32" > $O/r2_sweep_teddy48.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 1 --min-len 6 --max-len 6 --configs "queue=2;queue=0;wide=1" > $O/r2_sweep_noodle.log 2>&1
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "queue=2;queue=1;wide=1;first_stage=2,queue=1;first_stage=2,wide=1" > $O/r2_sweep_50k.log 2>&1
cat $O/r2_sweep_*.log

# ncu: the shipped FDR kernel, the restructured loop + queue (slower, cause unknown), the wide variant
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
for cfg in "queue=2" "queue=1" "wide=1" "wide=1,split=1" "wide=1,domain=12,replicas=8"; do
  tag=$(echo $cfg | tr ',=' '__')
  ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2_ncu_$tag.csv python tools/sweep.py --mb 512 --reps 1 --configs "$cfg" > $O/r2_ncu_$tag.out 2>&1
done
ls -la $O | tail -20
