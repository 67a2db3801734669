#!/usr/bin/env python
"""Random regular expressions through hs_compile's NFA route against the definition.

For every case: a random expression (literals, '.', classes, class escapes, groups, alternation,
quantifiers, optional start / end anchors) is compiled here, the UNMODIFIED reference hs_scan
(oracle/_ref) scans random data with the resulting single-outfix database, and the reported
end offsets must be exactly those the definition gives: every e such that some data[s:e] is in
the body's language (re.fullmatch), s restricted by the start anchor, e by the end anchor.
TEST INFRASTRUCTURE; CPU only.   python tests/fuzz_regex.py [--seconds 60] [--seed 1]"""
import argparse
import os
import re
import signal
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperscan_b200 import capi  # noqa: E402
import oracle.ref as ref  # noqa: E402

ALPHA = b"abc1 \n_X"


ANCHORS = False   # main(): anchors as atoms anywhere (most placements are refused, as the reference refuses them)


def gen_atom(rng, depth):
    r = rng.random()
    if ANCHORS and rng.random() < 0.12:
        a = [b"^", b"$", rb"\A", rb"\z", rb"\Z"][int(rng.integers(0, 5))]
        return b"(" + a + b"|" + gen_atom(rng, depth + 1) + b")" if rng.random() < 0.4 else \
            b"(" + a + b")?" if rng.random() < 0.3 else a
    if r < 0.45:
        return re.escape(bytes([ALPHA[rng.integers(0, len(ALPHA))]]))
    if r < 0.55:
        return b"."
    if r < 0.70:
        k = int(rng.integers(1, 4))
        body = b"".join(re.escape(bytes([ALPHA[rng.integers(0, len(ALPHA))]])) for _ in range(k))
        return (b"[^" if rng.random() < 0.3 else b"[") + body + b"]"
    if r < 0.80:
        return [rb"\d", rb"\w", rb"\s", rb"\D", rb"\W", rb"\S"][int(rng.integers(0, 6))]
    if r < 0.86:
        return [rb"\b", rb"\B"][int(rng.integers(0, 2))] + gen_atom(rng, depth + 1) if rng.random() < 0.5 \
            else gen_atom(rng, depth + 1) + [rb"\b", rb"\B"][int(rng.integers(0, 2))]
    if depth >= 2:
        return b"a"
    arms = [gen_seq(rng, depth + 1) for _ in range(int(rng.integers(1, 3)))]
    opener = [b"(?:", b"(", b"(?:", b"(", b"(?i:", b"(?s:", b"(?-i:", b"(?i-s:"][int(rng.integers(0, 8))]
    return opener + b"|".join(arms) + b")"


def gen_seq(rng, depth):
    out = b""
    for _ in range(int(rng.integers(1, 4))):
        a = gen_atom(rng, depth)
        q = rng.random()
        if q < 0.15:
            a += b"?"
        elif q < 0.27:
            a += b"*"
        elif q < 0.40:
            a += b"+"
        elif q < 0.50:
            big = rng.random() < 0.06                      # now and then a repeat that needs the 128- ... 512-state models
            lo = int(rng.integers(20, 90)) if big else int(rng.integers(0, 3))
            hi = lo + int(rng.integers(0, 3))
            a += [b"{%d}" % max(lo, 1), b"{%d,%d}" % (lo, max(hi, 1)), b"{%d,}" % lo][int(rng.integers(0, 3))]
        if q < 0.5 and rng.random() < 0.15:
            a += b"?"       # lazy: same set of ends
        out += a
    return out


EXT = {}   # main(): the hs_expr_ext of the case under test (min_offset, max_offset, min_length)


def _ext_ok(e):
    return e >= EXT.get("min_offset", 0) and e <= EXT.get("max_offset", 1 << 62)


def _starts(e):
    return range(e - max(EXT.get("min_length", 0), 1) + 1)


def definition(body, flags, start, end, data):
    """every end offset e such that the body matches some data[s:e] IN CONTEXT (\\b / \\B look at the bytes around
    the match): for each e the body is followed by a fixed-width look-behind that pins the match end to e"""
    fl = (re.I if flags & 1 else 0) | (re.S if flags & 2 else 0)
    n = len(data)
    ml = bool(flags & 4)
    out = set()
    for e in range(1, n + 1):
        if end == b"":
            ok = True
        elif end == rb"\z":
            ok = e == n
        elif not ml or end == rb"\Z":   # "$" / \Z (which ignores (?m)): at the end or before a final newline
            ok = e == n or (e == n - 1 and data[e:e + 1] == b"\n")
        else:               # "$" under (?m): before any newline or at the end
            ok = e == n or data[e:e + 1] == b"\n"
        if not ok or not _ext_ok(e):
            continue
        rx = re.compile(b"(?:" + body + b")(?<=(?s:\\A.{%d}))" % e, fl)
        for s in _starts(e):
            if start and not (s == 0 or (ml and data[s - 1:s] == b"\n")):
                continue
            if rx.match(data, s):
                out.add(e)
                break
    return sorted(out)


def definition_whole(expr, flags, data):
    """the same for an expression with anchors anywhere: Python's own reading of ^ $ \\A, with PCRE's \\z and \\Z
    spelled its way"""
    fl = (re.I if flags & 1 else 0) | (re.S if flags & 2 else 0) | (re.M if flags & 4 else 0)
    expr = expr.replace(b"\\Z", b"(?=\n?\0)").replace(b"\\z", b"\\Z").replace(b"\0", b"\\Z")
    out = []
    for e in range(1, len(data) + 1):
        rx = re.compile(b"(?:" + expr + b")(?<=(?s:\\A.{%d}))" % e, fl)
        if _ext_ok(e) and any(rx.match(data, s) for s in _starts(e)):
            out.append(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    capi.lib()
    t0, n, refused, literal, skipped, asserted = time.time(), 0, 0, 0, 0, 0

    def on_alarm(signum, frame):
        raise TimeoutError()
    signal.signal(signal.SIGALRM, on_alarm)
    global ANCHORS
    anchored_cases = 0
    while time.time() - t0 < args.seconds:
        ANCHORS = rng.random() < 0.35
        body = b"|".join(gen_seq(rng, 0) for _ in range(int(rng.integers(1, 3))))
        single_arm = b"|" not in body or body.count(b"(") > 0 and False
        start = rng.random() < 0.2
        end = [b"", b"", b"", b"$", rb"\z", rb"\Z"][int(rng.integers(0, 6))]
        flags = int(rng.choice([0, 0, 1, 2, 3, 4, 6]))
        # anchors bind to ONE top-level alternative in the expression: wrap the body to keep the definition simple
        expr = (b"^" if start else b"") + b"(?:" + body + b")" + end
        whole = ANCHORS and any(a in body for a in (b"^", b"$", rb"\A", rb"\z", rb"\Z"))
        if whole:
            expr = body
        try:
            signal.setitimer(signal.ITIMER_REAL, 2.0)      # (Python's matcher can blow up here too)
            if not whole and re.compile(body).fullmatch(b"") is not None:
                continue                                   # matches the empty string: refused by design
        except (re.error, TimeoutError):
            continue
        finally:
            signal.setitimer(signal.ITIMER_REAL, 0)
        if args.verbose:
            print("expr", expr, flags, flush=True)
        try:
            EXT.clear()
            if rng.random() < 0.3:     # extended parameters
                if rng.random() < 0.5:
                    EXT["min_offset"] = int(rng.integers(0, 12))
                if rng.random() < 0.5:
                    EXT["max_offset"] = EXT.get("min_offset", 0) + int(rng.integers(0, 20))
                if rng.random() < 0.5:
                    EXT["min_length"] = min(int(rng.integers(1, 7)), EXT.get("max_offset", 99))
            db = capi.compile_ext_multi([expr], [flags], [5], [dict(EXT) or None])
        except capi.HsError as e:
            refused += 1
            if args.verbose:
                print("refused", expr, str(e)[:80])
            continue
        if db.info().runtime_impl != 2:
            literal += 1
        n += 1
        asserted += (rb"\b" in body) or (rb"\B" in body)
        anchored_cases += whole
        for trial in range(3):
            size = int(rng.integers(1, 28))
            a = np.frombuffer(ALPHA, dtype=np.uint8)
            if re.search(rb"\{[2-9][0-9]", body) and trial:       # a long repeat: long data over few letters
                size = int(rng.integers(60, 160))
                a = np.frombuffer(ALPHA[:2] if trial == 1 else ALPHA[:4], dtype=np.uint8)
            data = a[rng.integers(0, a.size, size=size)].tobytes()
            arr = np.frombuffer(data, dtype=np.uint8)
            got = [int(r["to"]) for r in ref.scan_sorted(db.ptr, arr, np.array([0], np.uint64), np.array([size], np.uint32))]
            try:                                            # Python's backtracking matcher can blow up on nested
                signal.setitimer(signal.ITIMER_REAL, 2.0)   # nullable repeats: such a case is skipped, not judged
                want = definition_whole(expr, flags, data) if whole else definition(body, flags, start, end, data)
            except TimeoutError:
                skipped += 1
                continue
            finally:
                signal.setitimer(signal.ITIMER_REAL, 0)
            if got != want:
                print("MISMATCH expr", expr, "ext", EXT, "flags", flags, "data", data, "got", got, "want", want)
                sys.exit(1)
    print("fuzz regex: %d expressions (%d with \\b / \\B, %d with anchors inside groups, %d through the literal route; %d refused, %d inputs skipped: "
          "definition too slow), all equal to the definition (%.0f s)"
          % (n, asserted, anchored_cases, literal, refused, skipped, time.time() - t0))


if __name__ == "__main__":
    main()
