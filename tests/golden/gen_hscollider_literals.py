#!/usr/bin/env python
"""Generate tests/golden/hscollider_literals.json from the reference's own
recorded regression vectors (tools/hscollider/test_cases/{pcre,corpora}/*.txt).

Only entries on this path are kept: patterns this library's hs_compile accepts
(expressions that denote a finite set of literals: literal text, groups,
alternation, character classes, bounded repeats) with flags out of
{i, s, m, H, O}, together with the corpora
lines that carry recorded match offsets (`id="data": to1,to2,...`, format per
tools/hscollider/ColliderCorporaParser.rl:100-150; pattern flag letters per
util/ExpressionParser.rl:60-85).  The pattern TEXT and the recorded offsets are
data of the reference's test-suite, not code; they are stored here because
/root/reference does not exist on the GPU box.

Each kept entry is also run through the unmodified reference runtime
(oracle/_ref) at generation time: the recorded offsets and the reference
runtime must agree, or the generator aborts."""
import base64
import glob
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hyperscan_b200 import capi  # noqa: E402
import oracle.ref as ref  # noqa: E402

BASE = "/root/reference/tools/hscollider/test_cases"
FLAG = {"i": capi.HS_FLAG_CASELESS, "s": 2, "m": 4, "H": capi.HS_FLAG_SINGLEMATCH, "O": 0}
SPECIAL = {"0": 0, "a": 7, "e": 27, "f": 12, "n": 10, "v": 11, "r": 13, "t": 9}


def decode_corpus(s):
    """ColliderCorporaParser.rl corpus_new: \\xHH, \\[0aefnvrt], \\<non-alnum>."""
    out = bytearray()
    i = 0
    while i < len(s):
        c = s[i]
        if c == 0x5C and i + 1 < len(s):
            n = chr(s[i + 1])
            if n == "x" and i + 3 < len(s) and re.fullmatch(rb"[0-9a-fA-F]{2}", s[i + 2:i + 4]):
                out.append(int(s[i + 2:i + 4], 16))
                i += 4
                continue
            if n in SPECIAL:
                out.append(SPECIAL[n])
                i += 2
                continue
            if not n.isalnum():
                out.append(s[i + 1])
                i += 2
                continue
        out.append(c)
        i += 1
    return bytes(out)


def main():
    pats = {}
    for f in sorted(glob.glob(BASE + "/pcre/*.txt")):
        for line in open(f, "rb"):
            m = re.match(rb"^(\d+):/(.*)/([a-zA-Z8]*)\s*$", line.rstrip(b"\n"))
            if not m:
                continue
            pid, pat, fl = int(m.group(1)), m.group(2), m.group(3).decode()
            if set(fl) - set(FLAG):
                continue
            flags = 0
            for c in fl:
                flags |= FLAG[c]
            try:
                if capi.compile_multi([pat], [flags], [pid]).info().runtime_impl != 1:
                    continue  # not a finite set of literals: hscollider_regex.json has it
            except capi.HsError:
                continue  # needs the regex back end
            pats[pid] = (pat, fl, os.path.basename(f))
    cases = {}
    for f in sorted(glob.glob(BASE + "/corpora/*.txt")):
        for line in open(f, "rb"):
            m = re.match(rb'^(\d+)="(.*)":\s*([\d, ]*)\s*$', line.rstrip(b"\n"))
            if not m or int(m.group(1)) not in pats:
                continue
            pid = int(m.group(1))
            ends = sorted(int(x) for x in m.group(3).replace(b" ", b"").split(b",") if x)
            cases.setdefault(pid, []).append((decode_corpus(m.group(2)), ends, os.path.basename(f)))
    out = []
    for pid in sorted(cases):
        pat, fl, pfile = pats[pid]
        flags = 0
        for c in fl:
            flags |= FLAG[c]
        db = capi.compile_multi([pat], [flags], [pid])
        corp = []
        for data, ends, cfile in cases[pid]:
            arr = np.frombuffer(data, dtype=np.uint8)
            got = ref.scan_sorted(db.ptr, arr, np.array([0], dtype=np.uint64),
                                  np.array([len(data)], dtype=np.uint32))
            tos = [int(r["to"]) for r in got]
            if "H" in fl:
                # tools/hscollider/main.cpp:522-537: in single-match mode the scanner must
                # return exactly one of the recorded matches (when there are any)
                assert (len(tos) == 1 and tos[0] in ends) if ends else not tos, (pid, pat, ends, tos)
            else:
                assert tos == ends, (pid, pat, data, ends, got)
            corp.append({"data": base64.b64encode(data).decode(), "ends": ends, "file": cfile})
        out.append({"id": pid, "pattern": base64.b64encode(pat).decode(), "flag_letters": fl,
                    "hs_flags": flags, "file": pfile, "corpora": corp})
        print(pid, pat, fl, len(corp), "corpora")
    with open(os.path.join(ROOT, "tests", "golden", "hscollider_literals.json"), "w") as f:
        json.dump({"generator": "tests/golden/gen_hscollider_literals.py",
                   "source": "intel/hyperscan 5.4.2 tools/hscollider/test_cases (recorded offsets)",
                   "cases": out}, f, indent=0)
    print(len(out), "patterns,", sum(len(c["corpora"]) for c in out), "corpora")


if __name__ == "__main__":
    main()
