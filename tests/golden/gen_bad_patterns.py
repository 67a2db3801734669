#!/usr/bin/env python
"""tests/golden/bad_patterns.json from the reference's unit/hyperscan/bad_patterns.txt: expressions hs_compile must
refuse, with the reference's error message (`ID:/regex/flags{ext} #message`).  Test data of the reference's suite,
stored because /root/reference does not exist where the tests run."""
import base64
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FL = {"i": 1, "s": 2, "m": 4, "H": 8, "V": 16, "8": 32, "W": 64, "P": 128, "L": 256, "C": 512, "Q": 1024}
out = []
for line in open("/root/reference/unit/hyperscan/bad_patterns.txt", "rb"):
    m = re.match(rb"^(\d+):/(.*)/([a-zA-Z0-9]*)(\{[^}]*\})?\s+#(.*)$", line.rstrip(b"\n"))
    if not m:
        continue
    flags = 0
    for c in m.group(3).decode():
        flags |= FL.get(c, 0)
    ext = None
    if m.group(4):
        ext = {}
        for kv in m.group(4).decode()[1:-1].split(","):
            if "=" in kv:
                k, v = kv.split("=")
                ext[k.strip()] = int(v) if v.strip().lstrip("-").isdigit() else v.strip()
    out.append({"id": int(m.group(1)), "pattern": base64.b64encode(m.group(2)).decode(), "hs_flags": flags, "ext": ext,
                "message": m.group(5).decode("latin1").strip()})
with open(os.path.join(ROOT, "tests", "golden", "bad_patterns.json"), "w") as f:
    json.dump({"generator": "tests/golden/gen_bad_patterns.py", "source": "intel/hyperscan 5.4.2 unit/hyperscan/bad_patterns.txt",
               "cases": out}, f, indent=0)
print(len(out), "bad patterns")
