"""Definition-level oracle for literal pattern sets (SURVEY.md section 8c "O2").
TEST INFRASTRUCTURE ONLY.

A literal L with report id r matches with end offset `to` iff
data[to-len(L):to] == L, comparing ASCII letters case-insensitively iff
HS_FLAG_CASELESS (src/util/compare.h:49-107: only A-Z/a-z fold).  One report
per distinct (r, to) (dedupe, src/report.h:55-119); HS_FLAG_SINGLEMATCH keeps
only the smallest `to` per report id and block (src/report.h:121-147).
"""
import numpy as np

_UP = bytes(c - 32 if 97 <= c <= 122 else c for c in range(256))


def fold(b):
    return bytes(b).translate(_UP)


def scan_block(lits, flags, ids, data):
    data = bytes(data)
    up = None
    res = set()
    first = {}
    for lit, fl, rid in zip(lits, flags, ids):
        lit = bytes(lit)
        if fl & 1:
            if up is None:
                up = fold(data)
            hay, needle = up, fold(lit)
        else:
            hay, needle = data, lit
        pos = hay.find(needle)
        while pos >= 0:
            to = pos + len(needle)
            if fl & 8:
                if rid not in first or to < first[rid]:
                    first[rid] = to
                break
            res.add((rid, to))
            pos = hay.find(needle, pos + 1)
    for rid, to in first.items():
        res.add((rid, to))
    return res


def scan_blocks(lits, flags, ids, data, offsets, lengths):
    """Sorted (id, block, to) structured array over independent blocks."""
    from .ref import REC_DTYPE
    a = np.asarray(data).view(np.uint8).reshape(-1) if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), np.uint8)
    out = []
    for b, (o, n) in enumerate(zip(offsets, lengths)):
        for rid, to in scan_block(lits, flags, ids, a[int(o):int(o) + int(n)].tobytes()):
            out.append((rid, b, to))
    r = np.array(out, dtype=REC_DTYPE) if out else np.zeros(0, dtype=REC_DTYPE)
    return np.sort(r, order=["block", "to", "id"])
