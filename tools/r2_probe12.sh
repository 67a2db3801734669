#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2l}
python -m pytest tests -x -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
tail -c 400 $O/${P}_bench_n1.err
python - <<PY
import json
d = json.loads(open("$O/${P}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "wall_ms_per_step_with_barrier", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
for k, v in d["secondary"].items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("api", "note", "engine")})
PY
