#!/bin/bash
# final pass of the round: GPU tests, smoke, engine table, bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r02d}
python -m pytest tests -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
python - <<'PY'
import __graft_entry__ as g
g.smoke()
PY
python tools/dfa_bench.py --mb 256 > $O/${P}_dfa.log 2>&1
python tools/dfa_bench.py --mb 1024 >> $O/${P}_dfa.log 2>&1
cat $O/${P}_dfa.log
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
tail -c 300 $O/${P}_bench_n1.err; echo
