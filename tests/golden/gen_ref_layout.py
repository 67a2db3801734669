#!/usr/bin/env python
"""Regenerate tests/golden/ref_layout.json from the reference headers.

Runs ref_layout_dump() (oracle/ref/ref_driver.c), which is compiled against the
reference's own headers under /root/reference, and stores sizeof/offsetof of
every bytecode structure that hyperscan_b200/csrc/ref_layout.h restates.  The
committed JSON lets tests/test_layout.py pin our restated layouts on machines
that have no /root/reference (the GPU box).
"""
import ctypes, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(root, "oracle", "_ref", "libhsref_corei7.so")
code = "import ctypes,sys; ctypes.CDLL(sys.argv[1]).ref_layout_dump()"
out = subprocess.run([sys.executable, "-c", code, so], capture_output=True, text=True, check=True).stdout
data = json.loads(out)
with open(os.path.join(root, "tests", "golden", "ref_layout.json"), "w") as f:
    json.dump(data, f, indent=1, sort_keys=True)
print(len(data), "entries")
