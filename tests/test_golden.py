"""Committed golden vectors (tests/golden/literal_cases.json, produced by
tools/gen_golden.py from the unmodified reference runtime): the C restatement
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
