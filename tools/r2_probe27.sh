#!/bin/bash
# hsbench-format workload through tools/hsbench_b200.py on a B200
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r02i}
python -m pytest tests/test_hsbench_cli.py -q -m gpu > $O/${P}_cli_tests.log 2>&1
tail -2 $O/${P}_cli_tests.log
python tools/make_corpus.py --literals 1000 --blocks 131072 --block-len 1024 --out /tmp/c2 > $O/${P}_hsbench.log 2>&1
python tools/hsbench_b200.py -e /tmp/c2/sigs -c /tmp/c2/corpus.db -N -n 20 --literal-on --resident >> $O/${P}_hsbench.log 2>&1
python tools/hsbench_b200.py -e /tmp/c2/sigs -c /tmp/c2/corpus.db -n 5 --literal-on >> $O/${P}_hsbench.log 2>&1
cat $O/${P}_hsbench.log
