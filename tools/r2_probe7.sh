#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "heavy=0,big_set=0;heavy=2,big_set=0;heavy=2,big_set=1;heavy=2,big_set_classes=8;heavy=2,big_set_classes=2" > $O/r2g_sweep_50k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 5000 --configs "heavy=0;heavy=2" > $O/r2g_sweep_5k.log 2>&1
python tools/sweep.py --mb 1024 --reps 9 --configs "heavy=0;heavy=2;heavy=1" > $O/r2g_sweep_fdr1000.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --alphabet "abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn" --configs "heavy=0;heavy=2" > $O/r2g_sweep_lowercase.log 2>&1
cat $O/r2g_sweep_*.log
( time python -m pytest tests -x -q -m gpu ) > $O/r2g_gpu_tests.log 2>&1
grep -E "passed|failed" $O/r2g_gpu_tests.log
( time python bench.py --steps 5 --warmup 2 ) > $O/r2g_bench.json 2> $O/r2g_bench.err
grep -v "^$" $O/r2g_bench.err | tail -5
python -c "
import json
d=json.loads(open('$O/r2g_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ['value','ms_per_step','e2e','roofline','verify','secondary']}, indent=1)[:8000])"
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy"
ncu $SECTIONS --clock-control none -k regex:scanKernel -s 3 -c 1 --csv --page raw \
      --log-file $O/r2g_ncu_50k.csv python tools/sweep.py --mb 256 --reps 1 --lits 50000 --max-len 16 --configs "heavy=2,big_set=0" > $O/r2g_ncu_50k.out 2>&1
