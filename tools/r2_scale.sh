#!/bin/bash
# one 8-GPU box: the scaling series of bench.py (N = 1, 2, 4, 8; p2p exchange) and the NCCL exchange at N = 8
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2s}
run() { # n, extra flags, tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 "$@" > $O/${P}_bench_${tag}.json 2> $O/${P}_bench_${tag}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
        bench.py --gpus $n "$@" > $O/${P}_bench_${tag}.json 2> $O/${P}_bench_${tag}.err
  fi
  echo "== $tag rc=$?"; tail -c 300 $O/${P}_bench_${tag}.err
}
run 8 n8 
run 8 n8_nccl --exchange nccl --no-cpu
run 4 n4 --no-cpu
run 2 n2 --no-cpu
run 1 n1 --no-cpu
for f in $O/${P}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("secondary", {}).get("config5_shape_sharded", {})
    print(sys.argv[1], d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
          "cfg5", round(s.get("value_gbit_s", 0)), s.get("exchange"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
