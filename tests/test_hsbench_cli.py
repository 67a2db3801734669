"""tools/hsbench_b200.py: the reference benchmark's inputs (signature files `ID:/regex/flags`, the sqlite corpus
schema of tools/hsbench/data_corpus.cpp) and report through this library."""
import importlib.util
import io
import os
import re
from contextlib import redirect_stdout

import numpy as np
import pytest

import oracle.brute as brute

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("hsbench_b200", os.path.join(ROOT, "tools", "hsbench_b200.py"))
cli = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(cli)


def test_signature_lines():
    assert cli.parse_signature_line(b"12:/foo.*bar/is\n") == (12, b"foo.*bar", 3, None)
    assert cli.parse_signature_line(b"7:/a\\/b/H") == (7, b"a\\/b", 8, None)      # the LAST slash ends the regex
    assert cli.parse_signature_line(b"5:/x+y/i{min_offset=3,max_offset=40}") == (5, b"x+y", 1, {"min_offset": 3, "max_offset": 40})
    assert cli.parse_signature_line(b"  # comment\n") is None and cli.parse_signature_line(b"\n") is None
    assert cli.parse_signature_line(b"3:/x/8W")[2] == 32 | 64
    for bad in (b"nope", b"5:/x/k", b"5:/x/i{min_offset=x}", b"5:/x/i{som=3}"):
        with pytest.raises(ValueError):
            cli.parse_signature_line(bad)


def test_corpus_round_trip_and_packing(tmp_path):
    chunks = [(1, b"hello"), (2, b""), (1, b"world!" * 5), (3, bytes(range(256)))]
    path = str(tmp_path / "c.db")
    cli.write_corpus(path, chunks)
    assert cli.load_corpus(path) == chunks
    data, off, ln = cli.pack_blocks([d for _, d in chunks])
    assert ln.tolist() == [5, 0, 30, 256] and all(o % 16 == 0 for o in off.tolist())
    for (_, d), o, n in zip(chunks, off.tolist(), ln.tolist()):
        assert data[o:o + n].tobytes() == d


def _inputs(tmp_path, nstreams=24, per_stream=3, seed=3):
    rng = np.random.default_rng(seed)
    lits = [b"needle", b"hay", b"stack", b"ab"]
    sig = tmp_path / "sigs"
    sig.write_bytes(b"# four literals\n" + b"".join(b"%d:/%s/%s\n" % (10 + i, l, b"i" if i == 1 else b"")
                                                     for i, l in enumerate(lits)))
    chunks = []
    for k in range(per_stream):
        for s in range(nstreams):
            n = int(rng.integers(0, 200))
            d = bytearray(rng.integers(0x61, 0x7b, size=n, dtype=np.uint8).tobytes())
            for _ in range(n // 40):
                l = lits[int(rng.integers(0, len(lits)))]
                p = int(rng.integers(0, max(1, n - len(l))))
                d[p:p + len(l)] = l[:max(0, n - p)]
            chunks.append((s, bytes(d[:n])))
    corpus = tmp_path / "corpus.db"
    cli.write_corpus(str(corpus), chunks)
    return str(sig), str(corpus), lits, chunks


def _run(argv):
    out = io.StringIO()
    with redirect_stdout(out):
        rc = cli.main(argv)
    return rc, out.getvalue()


def _count(text):
    return int(re.search(r"Matches per iteration:\s+(\d+)", text).group(1))


def _brute_total(lits, blocks):
    data, off, ln = cli.pack_blocks(blocks)
    return int(brute.scan_blocks(lits, [0, 1, 0, 0], [10, 11, 12, 13], data, off, ln).size)


@pytest.mark.gpu
def test_block_mode_report(hs, tmp_path):
    sig, corpus, lits, chunks = _inputs(tmp_path)
    want = _brute_total(lits, [d for _, d in chunks])
    for extra in ([], ["--literal-on"], ["--per-call", "-n", "1"], ["--resident", "--per-scan"]):
        rc, text = _run(["-e", sig, "-c", corpus, "-N", "-n", "2"] + extra)
        assert rc == 0, text
        assert _count(text) == want and want > 20
        assert "Mean throughput (overall):" in text and "(%d blocks)" % len(chunks) in text
    rc, text = _run(["-e", sig, "-c", corpus, "-N", "-n", "1", "-z", "12"])          # one signature only
    assert rc == 0 and _count(text) == _brute_total([lits[2]] * 4, [d for _, d in chunks]) // 4


@pytest.mark.gpu
def test_streaming_mode_report(hs, tmp_path):
    sig, corpus, lits, chunks = _inputs(tmp_path)
    streams = {}
    for s, d in chunks:
        streams.setdefault(s, []).append(d)
    want = _brute_total(lits, [b"".join(v) for v in streams.values()])     # a stream's writes are one text
    rc, text = _run(["-e", sig, "-c", corpus, "-n", "2"])
    assert rc == 0, text
    assert _count(text) == want
    assert "(%d blocks in %d streams)" % (len(chunks), len(streams)) in text


@pytest.mark.gpu
def test_saved_database_is_loaded_back(hs, tmp_path):
    sig, corpus, lits, chunks = _inputs(tmp_path, nstreams=4, per_stream=1)
    rc, a = _run(["-e", sig, "-c", corpus, "-N", "-n", "1", "-w", str(tmp_path / "dbs")])
    rc2, b = _run(["-e", sig, "-c", corpus, "-N", "-n", "1", "-i", str(tmp_path / "dbs")])
    assert rc == 0 and rc2 == 0 and _count(a) == _count(b)
