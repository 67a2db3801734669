#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
P=r2k
python tools/sweep.py --mb 256 --reps 5 --lits 50000 --max-len 16 --configs "gram=0,heavy=2;gram=1;gram=1,warps=24" > $O/${P}_sweep_50k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 5000 --configs "gram=0;gram=0,heavy=2;gram=2" > $O/${P}_sweep_5k.log 2>&1
python tools/sweep.py --mb 512 --reps 5 --lits 20000 --max-len 12 --configs "gram=0;gram=2" > $O/${P}_sweep_20k.log 2>&1
python tools/sweep.py --mb 1024 --reps 7 --configs "gram=0;gram=0,heavy=2" > $O/${P}_sweep_fdr1000.log 2>&1
cat $O/${P}_sweep_*.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "4gram or layout or two_level or config5" > $O/${P}_gpu_tests.log 2>&1
grep -E "passed|failed" $O/${P}_gpu_tests.log
ncu --set full --import-source on --clock-control none -k regex:scanKernelGram -s 3 -c 1 -o $O/${P}_gram50k \
      python tools/sweep.py --mb 256 --reps 1 --lits 50000 --max-len 16 --configs "gram=1" > $O/${P}_ncu_50k.out 2>&1
ls -la $O/${P}_gram50k.ncu-rep
