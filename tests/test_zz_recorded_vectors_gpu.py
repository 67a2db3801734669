"""The reference's recorded hscollider vectors (tests/golden/hscollider_literals.json and
hscollider_regex.json, tests/golden/gen_hscollider_*.py) through the CUDA path.  The CPU half (C oracle
against the same fixture) is in tests/test_golden.py; this file sorts last on
purpose: it is the widest sweep over compiler-accepted expressions (groups,
alternation, classes, bounded repeats -> many literals under one id)."""
import base64

import numpy as np
import pytest

from test_golden import COLLIDER, COLLIDER_REGEX, _check_collider, _collider_blocks


ALL_CASES = COLLIDER + COLLIDER_REGEX   # literal route, then the regex route (single-outfix databases: McClellan / LimEx)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ALL_CASES, ids=[str(c["id"]) for c in ALL_CASES])
def test_cuda_path_reproduces_hscollider_vectors(hs, case):
    db = hs.compile_ext_multi([base64.b64decode(case["pattern"])], [case["hs_flags"]], [case["id"]], [case.get("ext")])
    data, off, ln, ends = _collider_blocks(case)
    scratch = hs.Scratch(db)
    got = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
    _check_collider(case, got, ends)
    # and one hs_scan per corpus, the way hscollider drives the engine
    for c, want in zip(case["corpora"], ends):
        tos = []
        hs.scan(db, base64.b64decode(c["data"]), scratch, on_event=lambda i, frm, to, fl: tos.append(to) or 0)
        if "H" in case["flag_letters"]:
            assert (len(tos) == 1 and tos[0] in want) if want else not tos
        else:
            assert tos == want


@pytest.mark.gpu
@pytest.mark.parametrize("case", COLLIDER_REGEX[1::3], ids=[str(c["id"]) for c in COLLIDER_REGEX[1::3]])
def test_cuda_path_reproduces_hscollider_vectors_limex_forced(hs, case):
    hs.set_build_option("regex_dfa", 0)
    try:
        db = hs.compile_ext_multi([base64.b64decode(case["pattern"])], [case["hs_flags"]], [case["id"]], [case.get("ext")])
    finally:
        hs.set_build_option("regex_dfa", 1)
    assert db.info().engine_id <= 5
    data, off, ln, ends = _collider_blocks(case)
    scratch = hs.Scratch(db)
    got = np.sort(hs.scan_blocks(db, data, off, ln, scratch), order=["block", "to", "id"])
    _check_collider(case, got, ends)


@pytest.mark.gpu
@pytest.mark.parametrize("gi", range(0, 120, 4))
def test_cuda_path_on_recorded_expressions_compiled_together(hs, gi):
    """six recorded expressions in one database (shared automaton: McClellan or LimEx-32 ... -512 by size) through
    the device path: under each member's id, the recorded matches"""
    from test_golden import _groups
    group = _groups()[gi]
    try:
        db = hs.compile_multi([base64.b64decode(c["pattern"]) for c in group], [c["hs_flags"] for c in group],
                              [1000 + k for k in range(len(group))])
    except hs.HsError:
        return
    scratch = hs.Scratch(db)
    for k, c in enumerate(group):
        data, off, ln, ends = _collider_blocks(c)
        got = hs.scan_blocks(db, data, off, ln, scratch)
        mine = got[got["id"] == 1000 + k]
        for b, want in enumerate(ends):
            assert sorted(int(r["to"]) for r in mine[mine["block"] == b]) == want
