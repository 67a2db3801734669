#!/bin/bash
# one ncu capture of a wide LimEx kernel (source counters)
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r02g}
ncu --set full --import-source on --clock-control none -k regex:dfaStaged -s 2 -c 1 -o $O/${P}_limex512 \
    python tools/dfa_bench.py --mb 64 --reps 1 --only limex512 > $O/${P}_ncu.out 2>&1
tail -3 $O/${P}_ncu.out
