#!/usr/bin/env python
"""The synthetic configurations of BASELINE.json in hsbench's own input formats: a signature file
(`ID:/literal/flags`, metacharacters escaped) and a sqlite corpus (`chunk(id, stream_id, data)`), so that the same
workload can be given to tools/hsbench_b200.py here and to a stock hsbench elsewhere.

  python tools/make_corpus.py --literals 1000 --blocks 65536 --block-len 1024 --out /tmp/c2
  python tools/hsbench_b200.py -e /tmp/c2/sigs -c /tmp/c2/corpus.db -N -n 20 --literal-on"""
import argparse
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hyperscan_b200 import synth  # noqa: E402
import hsbench_b200 as cli  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--literals", type=int, default=1000)
    ap.add_argument("--min-len", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=8)
    ap.add_argument("--blocks", type=int, default=65536)
    ap.add_argument("--block-len", type=int, default=1024)
    ap.add_argument("--streams", type=int, default=0, help="> 0: spread the blocks over this many streams")
    ap.add_argument("--plant-per-kb", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--out", required=True, help="directory for sigs and corpus.db")
    args = ap.parse_args()
    lits, flags, ids = synth.literal_set(args.literals, min_len=args.min_len, max_len=args.max_len, seed=args.seed)
    data, off, ln, _ = synth.block_corpus(args.blocks, args.block_len, lits, plant_per_kb=args.plant_per_kb, seed=args.seed + 5)
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, "sigs"), "wb") as fh:
        for l, f, i in zip(lits, flags, ids):
            fh.write(b"%d:/%s/%s\n" % (i, re.escape(l), b"i" if f & 1 else b""))
    raw = data.tobytes()
    nstreams = args.streams or args.blocks
    cli.write_corpus(os.path.join(args.out, "corpus.db"),
                     ((b % nstreams, raw[int(o):int(o) + int(n)]) for b, (o, n) in enumerate(zip(off, ln))))
    print("wrote %d signatures and %d blocks (%d bytes) to %s" % (len(lits), args.blocks, int(ln.sum()), args.out))


if __name__ == "__main__":
    main()
