"""Committed golden vectors (tests/golden/literal_cases.json, produced by
tests/golden/gen_literal_cases.py from the unmodified reference runtime): the C restatement
and -- on the GPU box -- the CUDA path must reproduce them.  These do not need
oracle/_ref or /root/reference at run time."""
import base64
import json
import os

import numpy as np
import pytest

import oracle.brute as brute
import oracle.port as port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "literal_cases.json")) as f:
    CASES = json.load(f)["cases"]


def load(case):
    lits = [base64.b64decode(x) for x in case["literals"]]
    data = np.frombuffer(base64.b64decode(case["corpus"]), dtype=np.uint8)
    off = np.array(case["offsets"], dtype=np.uint64)
    ln = np.array(case["lengths"], dtype=np.uint32)
    want = np.array([tuple(m) for m in case["matches"]],
                    dtype=[("id", "<u4"), ("block", "<u4"), ("to", "<u8")])
    return lits, case["flags"], case["ids"], data, off, ln, want


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracles_reproduce_golden(hs, case):
    lits, flags, ids, data, off, ln, want = load(case)
    db = hs.compile_lit_multi(lits, flags, ids)
    info = db.info()
    # the compiler still makes the engine choice the fixture was recorded with
    assert [info.hwlm_type, info.engine_id, info.fdr_domain, info.fdr_stride] == case["engine"]
    assert np.array_equal(port.scan_sorted(db.ptr, data, off, ln), want)
    assert np.array_equal(brute.scan_blocks(lits, flags, ids, data, off, ln), want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_reproduces_golden(hs, case):
    lits, flags, ids, data, off, ln, want = load(case)
    db = hs.compile_lit_multi(lits, flags, ids)
    scratch = hs.Scratch(db)
    got = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
    assert np.array_equal(got, want)
    corpus = hs.Corpus.upload(data, off, ln)
    got = np.sort(hs.scan_corpus(db, corpus, scratch), order=["block", "to", "id"])
    assert np.array_equal(got, want)


# --- the reference's own recorded hscollider vectors for literal patterns ---------------------------
# tests/golden/hscollider_literals.json (tests/golden/gen_hscollider_literals.py): pattern text and recorded
# end offsets from tools/hscollider/test_cases/{pcre,corpora}; compare rule for single-match patterns
# per tools/hscollider/main.cpp:522-537 (exactly one of the recorded matches).
with open(os.path.join(ROOT, "tests", "golden", "hscollider_literals.json")) as f:
    COLLIDER = json.load(f)["cases"]


def _collider_blocks(case):
    datas = [base64.b64decode(c["data"]) for c in case["corpora"]]
    # every corpus is one block at a 16-byte aligned start
    off, buf = [], bytearray()
    for d in datas:
        while len(buf) % 16:
            buf.append(0)
        off.append(len(buf))
        buf += d
    data = np.frombuffer(bytes(buf) + b"\0" * 16, dtype=np.uint8)
    return (data, np.array(off, dtype=np.uint64), np.array([len(d) for d in datas], dtype=np.uint32),
            [c["ends"] for c in case["corpora"]])


def _check_collider(case, recs, ends):
    for b, want in enumerate(ends):
        tos = [int(r["to"]) for r in recs if r["block"] == b]
        assert all(int(r["id"]) == case["id"] for r in recs)
        if "H" in case["flag_letters"]:
            assert (len(tos) == 1 and tos[0] in want) if want else not tos, (case["id"], b, tos, want)
        else:
            assert tos == want, (case["id"], b, tos, want)


@pytest.mark.parametrize("case", COLLIDER, ids=[str(c["id"]) for c in COLLIDER])
def test_oracle_reproduces_hscollider_vectors(hs, case):
    db = hs.compile_ext_multi([base64.b64decode(case["pattern"])], [case["hs_flags"]], [case["id"]], [case.get("ext")])
    data, off, ln, ends = _collider_blocks(case)
    _check_collider(case, port.scan_sorted(db.ptr, data, off, ln), ends)


# --- ... and for expressions that take the NFA route (regex_nfa.cpp -> LimEx-32 -> single-outfix database) --
# tests/golden/hscollider_regex.json (tests/golden/gen_hscollider_regex.py): 1 128 patterns / 9 838 corpora of the
# same suite that are NOT a finite set of literals and fit the NFA models.  The checker here is the unmodified
# reference runtime scanning the database this compiler emits: with its default engine choice (a McClellan DFA when
# the determinised automaton is small, else LimEx), and -- every third pattern -- with LimEx forced.
with open(os.path.join(ROOT, "tests", "golden", "hscollider_regex.json")) as f:
    COLLIDER_REGEX = json.load(f)["cases"]


@pytest.mark.parametrize("case", COLLIDER_REGEX, ids=[str(c["id"]) for c in COLLIDER_REGEX])
def test_reference_runtime_reproduces_hscollider_regex_vectors(hs, ref, case):
    db = hs.compile_ext_multi([base64.b64decode(case["pattern"])], [case["hs_flags"]], [case["id"]], [case.get("ext")])
    assert db.info().runtime_impl == 2
    data, off, ln, ends = _collider_blocks(case)
    _check_collider(case, ref.scan_sorted(db.ptr, data, off, ln), ends)


@pytest.mark.parametrize("case", COLLIDER_REGEX[::3], ids=[str(c["id"]) for c in COLLIDER_REGEX[::3]])
def test_reference_runtime_reproduces_hscollider_regex_vectors_limex_forced(hs, ref, case):
    hs.set_build_option("regex_dfa", 0)
    try:
        db = hs.compile_ext_multi([base64.b64decode(case["pattern"])], [case["hs_flags"]], [case["id"]], [case.get("ext")])
    finally:
        hs.set_build_option("regex_dfa", 1)
    assert db.info().runtime_impl == 2 and db.info().engine_id <= 5         # a LimEx model
    data, off, ln, ends = _collider_blocks(case)
    _check_collider(case, ref.scan_sorted(db.ptr, data, off, ln), ends)


def _groups(n_groups=120, size=6, seed=5):
    """random groups of recorded expressions (no single-match flag, no extended parameters) for the shared automaton"""
    rng = np.random.default_rng(seed)
    pool = [c for c in COLLIDER_REGEX if "H" not in c["flag_letters"] and not c.get("ext")]
    return [[pool[int(i)] for i in rng.choice(len(pool), size=size, replace=False)] for _ in range(n_groups)]


@pytest.mark.parametrize("gi", range(120))
def test_recorded_expressions_compiled_together(hs, ref, gi):
    """six recorded expressions in ONE database (one shared automaton, start / context helper states included):
    on every corpus of every member, the matches under the member's id are the recorded ones"""
    group = _groups()[gi]
    try:
        db = hs.compile_multi([base64.b64decode(c["pattern"]) for c in group], [c["hs_flags"] for c in group],
                              [1000 + k for k in range(len(group))])
    except hs.HsError as e:
        assert "too large" in str(e)          # the six together exceed the 512-state model
        return
    for k, c in enumerate(group):
        data, off, ln, ends = _collider_blocks(c)
        got = ref.scan_sorted(db.ptr, data, off, ln)
        mine = got[got["id"] == 1000 + k]
        for b, want in enumerate(ends):
            assert [int(r["to"]) for r in mine[mine["block"] == b]] == want, (base64.b64decode(c["pattern"]), b)
