#!/usr/bin/env python
"""Generate tests/golden/literal_cases.json: seeded literal sets + corpora and
the match lists the UNMODIFIED reference runtime (oracle/_ref, built from
/root/reference) delivers for them.  Committed so that parity can be checked
where neither /root/reference nor oracle/_ref exists."""
import base64
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hyperscan_b200 import capi, synth  # noqa: E402
import oracle.ref as ref  # noqa: E402

CASES = [
    ("noodle_caseless", dict(n=1, min_len=6, max_len=6, seed=31, caseless_frac=1.0), -1),
    ("teddy_8", dict(n=8, min_len=2, max_len=9, seed=32, caseless_frac=0.3), -1),
    ("teddy_48_packed", dict(n=48, min_len=3, max_len=8, seed=33, caseless_frac=0.2), -1),
    ("fdr_300", dict(n=300, min_len=4, max_len=12, seed=34, caseless_frac=0.2, singlematch_frac=0.1), -1),
    ("fdr_1000_config2_shape", dict(n=1000, min_len=4, max_len=8, seed=2, caseless_frac=0.1), -1),
]


def main():
    out = []
    for name, kw, engine in CASES:
        n = kw.pop("n")
        lits, flags, ids = synth.literal_set(n, alphabet=b"abcdefghij", **kw)
        db = capi.compile_lit_multi(lits, flags, ids)
        data, off, ln = synth.ragged_corpus([0, 5, 300, 1024, 2500, 63, 4097], lits, seed=len(name),
                                            plant_per_kb=10, alphabet=b"abcdefghijABCDEFGHIJ .")
        want = ref.scan_sorted(db.ptr, data, off, ln)
        info = db.info()
        out.append({
            "name": name,
            "literals": [base64.b64encode(x).decode() for x in lits],
            "flags": flags, "ids": ids,
            "corpus": base64.b64encode(data.tobytes()).decode(),
            "offsets": [int(x) for x in off], "lengths": [int(x) for x in ln],
            "engine": [int(info.hwlm_type), int(info.engine_id), int(info.fdr_domain), int(info.fdr_stride)],
            "matches": [[int(r["id"]), int(r["block"]), int(r["to"])] for r in want],
            "db_crc_serialized_len": len(db.serialize()),
        })
        print(name, "engine", out[-1]["engine"], "matches", len(want))
    with open(os.path.join(ROOT, "tests", "golden", "literal_cases.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_literal_cases.py", "reference": "intel/hyperscan 5.4.2 runtime (oracle/_ref)",
                   "cases": out}, f)


if __name__ == "__main__":
    main()
