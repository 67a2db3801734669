#!/usr/bin/env python
"""DFA engines on a resident corpus of uniform blocks: kernel ms and GB/s per engine kind
(run on the GPU box):  python tools/dfa_bench.py [--mb 256] [--block-len 1024]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperscan_b200 import capi, synth  # noqa: E402

KINDS = {"mcclellan16_2000lits": (2, 2000, 4, 8), "mcclellan8_30lits": (1, 30, 2, 4), "sheng_4lits": (3, 4, 1, 3),
         "limex32_6lits": (-1, 6, 4, 5), "limex128_12lits": (-1, 12, 6, 10), "limex256_20lits": (-1, 20, 8, 12),
         "limex512_40lits": (-1, 40, 8, 12)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--block-len", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ilp", type=int, default=1, help="blocks per lane (runtime option dfa_ilp)")
    ap.add_argument("--only", default="", help="run the engines whose name contains this")
    args = ap.parse_args()
    bl = args.block_len
    nb = (args.mb << 20) // bl
    capi.set_runtime_option("dfa_ilp", args.ilp)
    for name, (kind, nl, lo, hi) in KINDS.items():
        if args.only not in name:
            continue
        alpha = b"abcdefghijklmnopqrstuvwxyz" if nl > 100 else (b"abcdefgh" if nl > 4 else b"abc")
        lits, flags, ids = synth.literal_set(nl, min_len=lo, max_len=hi, seed=nl, caseless_frac=0.0, alphabet=alpha)
        if kind < 0:      # LimEx position automaton: the model that holds the literals' bytes + 1 states
            eng = capi.limex32_from_literals(lits, None, ids)
        else:
            eng = capi.dfa_from_literals(lits, None, ids, kind=kind)
        data, off, ln, _ = synth.block_corpus(nb, bl, lits, plant_per_kb=0.01, seed=3)
        corpus = capi.Corpus.upload(data, off, ln)
        ms = []
        for i in range(2 + args.reps):
            got, kms = capi.nfa_scan_corpus(eng, corpus)
            if i >= 2:
                ms.append(kms)
        k = float(np.median(ms))
        print(json.dumps({"engine": name, "engine_bytes": len(eng), "blocks": nb, "block_len": bl, "ilp": args.ilp, "ms": round(k, 4),
                          "GBps": round(nb * bl / (k * 1e-3) / 1e9, 1), "records": int(got.size)}), flush=True)
        corpus.free()


if __name__ == "__main__":
    main()
