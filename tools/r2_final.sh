#!/bin/bash
# Round-end measurement pass on one B200 (N=1): everything profiles/README.md cites.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r02}
python -m pytest tests -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -2 $O/${P}_gpu_tests.log
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
tail -c 300 $O/${P}_bench_n1.err; echo
python bench.py --impl reference > $O/${P}_bench_reference_arm.json 2> $O/${P}_bench_reference_arm.err
tail -c 600 $O/${P}_bench_reference_arm.json; echo
bash tools/capture_traffic.sh > $O/${P}_traffic.out 2>&1
tail -2 $O/${P}_traffic.out
# every kernel launch of a short bench run with its device time (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --passes-per-step 4 --no-cpu --no-secondary > $O/${P}_launches_bench.out 2>&1
# the headline kernel, the Teddy kernel, the 4-gram kernel, the DFA kernels: one --set full capture each
ncu --set full --import-source on --clock-control none -k regex:scanKernelPair -s 3 -c 1 -o $O/${P}_pair \
    python tools/sweep.py --mb 1024 --reps 1 --configs "gram=1" > $O/${P}_ncu_pair.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:scanKernelWide -s 3 -c 1 -o $O/${P}_teddy \
    python tools/sweep.py --mb 512 --reps 1 --lits 48 --configs "gram=1" > $O/${P}_ncu_teddy.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:scanKernelGram -s 3 -c 1 -o $O/${P}_gram \
    python tools/sweep.py --mb 256 --reps 1 --lits 50000 --max-len 16 --configs "gram=1" > $O/${P}_ncu_gram.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:dfaStaged -s 4 -c 1 -o $O/${P}_dfa_mcc8 \
    python tools/dfa_bench.py --mb 256 --reps 1 > $O/${P}_ncu_dfa1.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:dfaStaged -s 7 -c 1 -o $O/${P}_dfa_sheng \
    python tools/dfa_bench.py --mb 256 --reps 1 > $O/${P}_ncu_dfa2.out 2>&1
python tools/dfa_bench.py --mb 256 > $O/${P}_dfa.log 2>&1
python tools/dfa_bench.py --mb 1024 >> $O/${P}_dfa.log 2>&1
cat $O/${P}_dfa.log
python tools/sweep.py --mb 1024 --reps 7 --configs "gram=1" > $O/${P}_sweep_headline.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 48 --configs "gram=1" >> $O/${P}_sweep_headline.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 8 --configs "gram=1" >> $O/${P}_sweep_headline.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 1 --min-len 6 --max-len 6 --configs "gram=1" >> $O/${P}_sweep_headline.log 2>&1
python tools/sweep.py --mb 512 --reps 7 --lits 5000 --configs "gram=1" >> $O/${P}_sweep_headline.log 2>&1
cut -c1-160 $O/${P}_sweep_headline.log
ls -la $O/${P}_*.ncu-rep
