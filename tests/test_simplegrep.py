"""BASELINE config 1: the reference's simplegrep flow as a plain C program
against include/hs_b200.h + libhs_b200.so (examples/simplegrep_b200.c).  The
header must be valid C99; on the GPU box the program's matches must equal the
reference's for 1 literal over a 1 MB ASCII buffer."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hyperscan_b200", "lib")


def build_example(tmp_path, hs):
    exe = str(tmp_path / "simplegrep_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O2", "-o", exe,
                    os.path.join(ROOT, "examples", "simplegrep_b200.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", LIBDIR, "-lhs_b200", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def test_header_is_c99_and_example_links(tmp_path, hs):
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(ROOT, "include", "hs_b200.h")], check=True)
    exe = build_example(tmp_path, hs)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage" in r.stderr
    # a regex that needs the regex back end is refused at compile time with a message
    f = tmp_path / "in.txt"
    f.write_bytes(b"hello")
    r = subprocess.run([exe, "a.*(?=b)b", str(f)], capture_output=True, text=True)
    assert r.returncode != 0 and "Unable to compile pattern" in r.stderr


@pytest.mark.gpu
def test_simplegrep_config1(tmp_path, hs, ref, real_gpu):
    exe = build_example(tmp_path, hs)
    rng = np.random.default_rng(1)
    data = rng.integers(0x20, 0x7F, size=1 << 20, dtype=np.uint8)
    lit = b"needle"
    for pos in (0, 13, 127, 4093, 65531, (1 << 20) - len(lit)):   # incl. offset 0, the very end, tile edges
        data[pos:pos + len(lit)] = np.frombuffer(lit, dtype=np.uint8)
    f = tmp_path / "corpus.txt"
    f.write_bytes(data.tobytes())
    r = subprocess.run([exe, lit.decode(), str(f)], capture_output=True, text=True, check=True)
    got = [int(line.rsplit(" ", 1)[1]) for line in r.stdout.splitlines() if line.startswith("Match for")]
    db = hs.compile_multi([lit])
    want = [int(x["to"]) for x in ref.scan_sorted(db.ptr, data, [0], [data.size])]
    assert got == want and len(got) >= 6
