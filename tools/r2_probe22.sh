#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2v}
python -m pytest tests -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
python - <<'PY'
import __graft_entry__ as g
g.smoke()
PY
