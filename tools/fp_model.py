#!/usr/bin/env python
"""Offline model of first-stage filter designs: candidate rate per KB for a key
function / slot set / bucket count, on synthetic corpora.  Reads the literal
tails and buckets out of a compiled database (LitInfo chains).  Its prediction
for the shipped design (FDR hash, domain 13, slots 1..4: 0.260 candidates/KB)
matches the device counters (0.271).

  python tools/fp_model.py [designs|extra|slots|classes]
"""
import ctypes as C
import struct
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperscan_b200 import capi, synth


def tails_from_db(db):
    blob = db.serialize()
    bc = blob[32:]
    fm = struct.unpack_from("<I", bc, 96)[0]
    eng = fm + 192
    engineID, size, maxlen, nstr, confOff = struct.unpack_from("<5I", bc, eng)
    conf = eng + confOff
    out = []
    nb = 8
    for b in range(nb):
        cf = struct.unpack_from("<I", bc, conf + 4 * b)[0]
        if not cf:
            continue
        fc = conf + cf
        andmsk, mult, nbits = struct.unpack_from("<QQI", bc, fc)
        seen = set()
        for c in range(1 << nbits):
            st = struct.unpack_from("<I", bc, fc + 32 + 4 * c)[0]
            if not st:
                continue
            li = fc + st
            while True:
                v, msk, groups, lid, sz, fl, nxt = struct.unpack_from("<QQQIBBB", bc, li)
                out.append((b, v, msk, sz))
                if not nxt:
                    break
                li += 32
    return out


def build_table(tails, keyfn, nkeys, slots, bucket_of, nb):
    """possible[key, slot] bitmask of buckets (bit set = possible)."""
    T = np.zeros((nkeys, slots), dtype=np.uint32)
    allb = np.arange(256, dtype=np.uint32)
    for (b, v, msk, sz) in tails:
        bk = bucket_of(b)
        for p in range(slots):
            # char at distance p from end = byte lane 7-p ; next byte (distance p-1) = lane 8-p
            if p >= sz:
                T[:, p] |= 1 << bk
                continue
            c0 = (v >> (8 * (7 - p))) & 0xff
            m0 = (msk >> (8 * (7 - p))) & 0xff
            b0s = allb[(allb & m0) == c0]
            if p == 0:
                b1s = allb
            else:
                c1 = (v >> (8 * (8 - p))) & 0xff
                m1 = (msk >> (8 * (8 - p))) & 0xff
                b1s = allb[(allb & m1) == c1]
            keys = np.unique(keyfn(b0s[:, None], b1s[None, :]).ravel())
            T[keys, p] |= 1 << bk
    return T


def cand_rate(T, keyfn, text, slots, stride=1):
    n = text.size - 1
    k = keyfn(text[:-1].astype(np.uint32), text[1:].astype(np.uint32))
    poss = np.full(n, 0xffffffff, dtype=np.uint32)
    # end position e gets constraint from sample x=e-p slot p
    for p in range(slots):
        e = T[k, p]          # for sample positions x: applies to end x+p
        contrib = np.full(n, 0xffffffff, dtype=np.uint32)
        xs = np.arange(0, n - p)
        if stride > 1:
            sel = (xs % stride) == 0
            contrib[xs[sel] + p] = e[xs[sel]]
        else:
            contrib[p:] = e[: n - p]
        poss &= contrib
    # count candidate (bucket,position) pairs and positions
    pos_c = np.count_nonzero(poss)
    return pos_c / (n / 1024.0)


def main():
    lits, flags, ids = synth.literal_set(1000)
    db = capi.compile_lit_multi(lits, flags, ids)
    tails = tails_from_db(db)
    print(len(tails), "tails")
    data, off, ln, _ = synth.block_corpus(2048, 1024, lits, plant_per_kb=0.0)
    lower = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn", dtype=np.uint8)
    text2 = lower[np.random.default_rng(5).integers(0, lower.size, size=data.size)]
    # alphabet classes (case folded)
    sigma = set()
    for (b, v, msk, sz) in tails:
        for p in range(min(sz, 8)):
            sigma.add(((v >> (8 * (7 - p))) & 0xff) | 0x20)
    cls = np.full(256, 31, dtype=np.uint32)
    for i, c in enumerate(sorted(sigma)):
        cls[c] = min(i, 30)
        if 0x61 <= c <= 0x7a:
            cls[c - 0x20] = min(i, 30)
    designs = {
        "fdr13 (b0, b1&31) 8bk": (lambda a, b: a | ((b & 31) << 8), 1 << 13, 8),
        "k12 (b0&127, b1&31) 8bk": (lambda a, b: (a & 127) | ((b & 31) << 7), 1 << 12, 8),
        "k11 (b0&63, b1&31) 8bk": (lambda a, b: (a & 63) | ((b & 31) << 6), 1 << 11, 8),
        "k11b (b0&31|bit6, b1&31)": (lambda a, b: (a & 31) | ((a >> 1) & 32) | ((b & 31) << 6), 1 << 11, 8),
        "k10 (b0&31, b1&31) 8bk": (lambda a, b: (a & 31) | ((b & 31) << 5), 1 << 10, 8),
        "cls10 (cls b0, cls b1) 8bk": (lambda a, b: cls[a] | (cls[b] << 5), 1 << 10, 8),
        "k15 (b0, b1&127) 8bk": (lambda a, b: a | ((b & 127) << 8), 1 << 15, 8),
        "k11 4 buckets": (lambda a, b: (a & 63) | ((b & 31) << 6), 1 << 11, 4),
    }
    for name, (fn, nk, nb) in designs.items():
        bo = (lambda b: b) if nb == 8 else (lambda b: b // 2)
        for slots in (4, 8):
            T = build_table(tails, fn, nk, slots, bo, nb)
            r1 = cand_rate(T, fn, data, slots, 1)
            r2 = cand_rate(T, fn, data, slots, 2)
            rl = cand_rate(T, fn, text2, slots, 1)
            print("%-28s slots %d: printable s1 %.3f/KB  s2 %.2f/KB | lowercase s1 %.2f/KB" % (name, slots, r1, r2, rl))




def extra():
    lits, flags, ids = synth.literal_set(1000)
    db = capi.compile_lit_multi(lits, flags, ids)
    tails = tails_from_db(db)
    data, off, ln, _ = synth.block_corpus(2048, 1024, lits, plant_per_kb=0.0)
    for d in (9, 10, 11, 12, 13, 14):
        fn = (lambda dd: (lambda a, b: (a | (b << 8)) & ((1 << dd) - 1)))(d)
        T = build_table(tails, fn, 1 << d, 4, lambda b: b, 8)
        print("contiguous domain %d: s1 %.3f/KB" % (d, cand_rate(T, fn, data, 4, 1)))





def build_table_slots(tails, keyfn, nkeys, slot_list, nb=8):
    T = np.zeros((nkeys, len(slot_list)), dtype=np.uint32)
    allb = np.arange(256, dtype=np.uint32)
    for (b, v, msk, sz) in tails:
        for si, p in enumerate(slot_list):
            if p >= sz:
                T[:, si] |= 1 << b
                continue
            c0 = (v >> (8 * (7 - p))) & 0xff
            m0 = (msk >> (8 * (7 - p))) & 0xff
            b0s = allb[(allb & m0) == c0]
            if p == 0:
                b1s = allb
            else:
                c1 = (v >> (8 * (8 - p))) & 0xff
                m1 = (msk >> (8 * (8 - p))) & 0xff
                b1s = allb[(allb & m1) == c1]
            keys = np.unique(keyfn(b0s[:, None], b1s[None, :]).ravel())
            T[keys, si] |= 1 << b
    return T


def cand_rate_slots(T, keyfn, text, slot_list):
    n = text.size - 1
    k = keyfn(text[:-1].astype(np.uint32), text[1:].astype(np.uint32))
    poss = np.full(n, 0xffffffff, dtype=np.uint32)
    for si, p in enumerate(slot_list):
        contrib = np.full(n, 0xffffffff, dtype=np.uint32)
        contrib[p:] = T[k, si][: n - p]
        poss &= contrib
    return np.count_nonzero(poss) / (n / 1024.0)


def slots_experiment():
    lits, flags, ids = synth.literal_set(1000)
    db = capi.compile_lit_multi(lits, flags, ids)
    tails = tails_from_db(db)
    data, off, ln, _ = synth.block_corpus(2048, 1024, lits, plant_per_kb=0.0)
    lower = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn", dtype=np.uint8)
    text2 = lower[np.random.default_rng(5).integers(0, lower.size, size=data.size)]
    for d in (11, 12, 13):
        fn = (lambda dd: (lambda a, b: (a | (b << 8)) & ((1 << dd) - 1)))(d)
        for sl in ([0, 1, 2, 3], [1, 2, 3, 4], [1, 2, 3, 5], [1, 2, 3, 4, 5, 6, 7]):
            T = build_table_slots(tails, fn, 1 << d, sl)
            print("domain %d slots %s: printable %.3f/KB lowercase %.2f/KB" %
                  (d, sl, cand_rate_slots(T, fn, data, sl), cand_rate_slots(T, fn, text2, sl)))




def class_map(tails, nclass=32, slots=(1, 2, 3, 4)):
    """byte -> class (nclass-1 = 'other': not in any literal)."""
    freq = np.zeros(256, dtype=np.int64)
    allb = np.arange(256, dtype=np.uint32)
    for (b, v, msk, sz) in tails:
        for p in set(slots) | {q - 1 for q in slots if q > 0}:
            if p >= sz or p > 7:
                continue
            c = (v >> (8 * (7 - p))) & 0xff
            m = (msk >> (8 * (7 - p))) & 0xff
            freq[allb[(allb & m) == c]] += 1
    used = np.nonzero(freq)[0]
    order = used[np.argsort(-freq[used], kind="stable")]
    cls = np.full(256, nclass - 1, dtype=np.uint32)
    own = nclass - 1
    if len(order) <= own:
        for i, bb in enumerate(order):
            cls[bb] = i
    else:
        shared = max(1, min(8, own // 4))
        for i, bb in enumerate(order):
            if i < own - shared:
                cls[bb] = i
            else:
                cls[bb] = own - shared + (i % shared)
    return cls


def class_experiment():
    lits, flags, ids = synth.literal_set(1000)
    db = capi.compile_lit_multi(lits, flags, ids)
    tails = tails_from_db(db)
    data, off, ln, _ = synth.block_corpus(2048, 1024, lits, plant_per_kb=0.0)
    lower = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz     eeeettaaooiinn", dtype=np.uint8)
    text2 = lower[np.random.default_rng(5).integers(0, lower.size, size=data.size)]
    mixed = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 .,;:-_/", dtype=np.uint8)
    text3 = mixed[np.random.default_rng(6).integers(0, mixed.size, size=data.size)]
    for nclass in (32,):
        cls = class_map(tails, nclass)
        print("classes used", len(set(cls.tolist())))
        fn = lambda a, b: cls[a] | (cls[b] << 5)
        for sl in ([1, 2, 3, 4], [0, 1, 2, 3]):
            T = build_table_slots(tails, fn, 1 << 10, sl)
            print("cls%d slots %s: printable %.3f/KB lowercase %.2f/KB mixed %.2f/KB" %
                  (nclass, sl, cand_rate_slots(T, fn, data, sl), cand_rate_slots(T, fn, text2, sl),
                   cand_rate_slots(T, fn, text3, sl)))
    fn = lambda a, b: (a | (b << 8)) & 0x1fff
    T = build_table_slots(tails, fn, 1 << 13, [1, 2, 3, 4])
    print("fdr13 slots 1-4: printable %.3f lowercase %.2f mixed %.2f" % (
        cand_rate_slots(T, fn, data, [1, 2, 3, 4]), cand_rate_slots(T, fn, text2, [1, 2, 3, 4]),
        cand_rate_slots(T, fn, text3, [1, 2, 3, 4])))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "designs"
    {"designs": main, "extra": extra, "slots": slots_experiment, "classes": class_experiment}[which]()
