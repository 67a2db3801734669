#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2m}
python tools/sweep.py --mb 512 --reps 7 --lits 96 --avx2 --configs "fat_pair=0;fat_pair=1;fat_pair=1,heavy=0" > $O/${P}_sweep_fat96.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 60 --avx2 --configs "fat_pair=0;fat_pair=1" > $O/${P}_sweep_fat60.log 2>&1
cat $O/${P}_sweep_fat*.log
python -m pytest tests -x -q -m gpu -k "fat" > $O/${P}_gpu_tests.log 2>&1
tail -2 $O/${P}_gpu_tests.log
