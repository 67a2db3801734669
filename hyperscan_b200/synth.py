"""Seeded synthetic workloads of the shapes BASELINE.json names (SURVEY.md
section 8d): literal sets and block corpora.  Used by bench.py and tests/.

There is no network and hsbench's corpora are SQLite files of real traffic, so
every workload is generated: uniform printable ASCII (0x20-0x7e) with pattern
literals planted at a seeded rate.
"""
import numpy as np

HS_FLAG_CASELESS = 1
HS_FLAG_SINGLEMATCH = 8


def literal_set(n, min_len=4, max_len=8, caseless_frac=0.1, seed=2, alphabet=b"abcdefghijklmnopqrstuvwxyz",
                singlematch_frac=0.0):
    """n distinct literals, lengths uniform in [min_len, max_len] over
    `alphabet`; a seeded fraction carries HS_FLAG_CASELESS / SINGLEMATCH.
    Returns (lits, flags, ids)."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    seen = set()
    lits = []
    while len(lits) < n:
        L = int(rng.integers(min_len, max_len + 1))
        s = alpha[rng.integers(0, alpha.size, size=L)].tobytes()
        if s in seen:
            continue
        seen.add(s)
        lits.append(s)
    flags = []
    for _ in range(n):
        f = 0
        if rng.random() < caseless_frac:
            f |= HS_FLAG_CASELESS
        if rng.random() < singlematch_frac:
            f |= HS_FLAG_SINGLEMATCH
        flags.append(f)
    return lits, flags, list(range(n))


def block_corpus(nblocks, block_len, lits=None, plant_per_kb=0.01, seed=7, pitch=None):
    """nblocks blocks of block_len bytes at a fixed 16-byte aligned pitch.
    Returns (data uint8[nblocks*pitch], offsets u64, lengths u32, planted) where
    planted is a list of (block, to, literal index) of the literals written in
    (later plants may overwrite earlier ones; it is a seed of matches, not the
    expected match set)."""
    rng = np.random.default_rng(seed)
    pitch = pitch or ((block_len + 15) // 16) * 16
    total = nblocks * pitch
    data = rng.integers(0x20, 0x7F, size=total, dtype=np.uint8)
    offsets = (np.arange(nblocks, dtype=np.uint64) * np.uint64(pitch))
    lengths = np.full(nblocks, block_len, dtype=np.uint32)
    planted = []
    if lits and plant_per_kb > 0:
        nplant = int(nblocks * block_len / 1024.0 * plant_per_kb)
        which = rng.integers(0, len(lits), size=nplant)
        blk = rng.integers(0, nblocks, size=nplant)
        pos = rng.random(size=nplant)
        for w, b, p in zip(which, blk, pos):
            lit = lits[int(w)]
            if len(lit) > block_len:
                continue
            start = int(p * (block_len - len(lit) + 1))
            o = int(b) * pitch + start
            data[o:o + len(lit)] = np.frombuffer(lit, dtype=np.uint8)
            planted.append((int(b), start + len(lit), int(w)))
    return data, offsets, lengths, planted


def ragged_corpus(lengths, lits=None, plant_per_kb=0.05, seed=11, align=16, alphabet=None):
    """Blocks of the given (ragged, possibly zero) lengths packed at `align`-byte
    aligned starts (align=1: back to back, exercising the unaligned host path)."""
    rng = np.random.default_rng(seed)
    offs = []
    pos = 0
    for n in lengths:
        pos = (pos + align - 1) // align * align
        offs.append(pos)
        pos += int(n)
    total = max(pos, 1)
    if alphabet is None:
        data = rng.integers(0x20, 0x7F, size=total, dtype=np.uint8)
    else:
        al = np.frombuffer(alphabet, dtype=np.uint8)
        data = al[rng.integers(0, al.size, size=total)]
    if lits and plant_per_kb > 0:
        for b, (o, n) in enumerate(zip(offs, lengths)):
            k = rng.poisson(n / 1024.0 * plant_per_kb)
            for _ in range(k):
                lit = lits[int(rng.integers(0, len(lits)))]
                if len(lit) > n:
                    continue
                s = int(rng.integers(0, n - len(lit) + 1))
                data[o + s:o + s + len(lit)] = np.frombuffer(lit, dtype=np.uint8)
    return data, np.array(offs, dtype=np.uint64), np.array(lengths, dtype=np.uint32)
