#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2q}
python -m pytest tests/test_dfa.py tests/test_small_write.py -x -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -2 $O/${P}_gpu_tests.log
HSB200_DFA_ILP=1 python -m pytest tests/test_dfa.py -x -q -m gpu > $O/${P}_gpu_tests_ilp1.log 2>&1
tail -2 $O/${P}_gpu_tests_ilp1.log
for ilp in 1 2; do
python tools/dfa_bench.py --mb 256 --ilp $ilp >> $O/${P}_dfa.log 2>&1
python tools/dfa_bench.py --mb 1024 --ilp $ilp >> $O/${P}_dfa.log 2>&1
python tools/dfa_bench.py --mb 256 --block-len 4096 --ilp $ilp >> $O/${P}_dfa.log 2>&1
done
cat $O/${P}_dfa.log
ncu --set full --import-source on --clock-control none -k regex:dfaStaged -s 4 -c 1 -o $O/${P}_dfa_mcc8 \
   python tools/dfa_bench.py --mb 256 --reps 1 > $O/${P}_ncu1.out 2>&1
ls -la $O/${P}_dfa_*.ncu-rep
