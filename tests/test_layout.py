"""Database layout pins: every sizeof/offsetof hyperscan_b200/csrc/ref_layout.h
restates equals the value the reference's own headers give
(tests/golden/ref_layout.json, produced by tests/golden/gen_ref_layout.py from
/root/reference).  The serialized database format is the drop-in boundary."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restated_layouts_match_reference(tmp_path):
    exe = str(tmp_path / "layout_check")
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "hyperscan_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "layout_check.cpp")], check=True)
    ours = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    with open(os.path.join(ROOT, "tests", "golden", "ref_layout.json")) as f:
        gold = json.load(f)
    assert len(ours) > 100
    bad = {k: (v, gold.get(k)) for k, v in ours.items() if gold.get(k) != v}
    assert not bad, bad


def test_golden_matches_live_reference_headers():
    """When the reference runtime is built here, the golden file is current."""
    import pytest
    so = os.path.join(ROOT, "oracle", "_ref", "libhsref_corei7.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    code = "import ctypes,sys; ctypes.CDLL(sys.argv[1]).ref_layout_dump()"
    import sys
    out = subprocess.run([sys.executable, "-c", code, so], capture_output=True, text=True, check=True).stdout
    live = json.loads(out)
    with open(os.path.join(ROOT, "tests", "golden", "ref_layout.json")) as f:
        gold = json.load(f)
    assert live == gold
