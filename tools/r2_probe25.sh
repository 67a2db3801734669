#!/bin/bash
# engine table + the regex tests on a B200 (short)
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r02e}
python -m pytest tests/test_limex.py tests/test_regex.py tests/test_outfix.py -q -m gpu > $O/${P}_gpu_tests.log 2>&1
tail -3 $O/${P}_gpu_tests.log
python tools/dfa_bench.py --mb 256 > $O/${P}_dfa.log 2>&1
cat $O/${P}_dfa.log
python - > $O/${P}_regex.log 2>&1 <<'PY'
import sys, json, numpy as np
sys.path.insert(0, ".")
from hyperscan_b200 import capi, synth
pats = [rb"ab+c", rb"[0-9]{2,}\.[0-9]", rb"^GET\s", rb"(foo|bar)x*z", rb"q.{2,4}w$"]
nb, bl = 262144, 1024
data, off, ln, _ = synth.block_corpus(nb, bl, [b"abbbc", b"123.4", b"GET /", b"fooxxz", b"q123w"], plant_per_kb=0.05, seed=96)
corpus = capi.Corpus.upload(data, off, ln)
for dfa in (1, 0):
    capi.set_build_option("regex_dfa", dfa)
    db = capi.compile_multi(pats, [0, 0, 0, capi.HS_FLAG_CASELESS, 0], list(range(1, 6)))
    sc = capi.Scratch(db)
    ms = []
    for i in range(6):
        capi.scan_corpus_async(db, corpus, sc)
        rc, n, _ = capi.scan_corpus_finish(sc)
        if rc == capi.HS_INSUFFICIENT_SPACE:
            continue
        capi._check(rc, "regex")
        if i >= 2:
            ms.append(sc.last_kernel_ms())
    k = float(np.median(ms))
    print(json.dumps({"regex_dfa": dfa, "engine_id": int(db.info().engine_id), "states": int(db.info().num_literals),
                      "kernel_ms": k, "GBps": nb * bl / (k * 1e-3) / 1e9, "matches": int(n)}), flush=True)
    sc.free()
PY
cat $O/${P}_regex.log
