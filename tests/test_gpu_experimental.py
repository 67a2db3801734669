"""Parity of every kernel variant a runtime option can select: the 16-byte-lane
kernels (queued and not), the wide-step kernels with and without the split
confirm, their 768 / 896 / 1024-thread builds, the hash-table and per-byte first
stages for FDR sets (the class-pair kernel is the default and is what
tests/test_gpu_parity.py runs).  Same checks as tests/test_gpu_parity.py::_check_all."""
import numpy as np
import pytest

from hyperscan_b200 import synth
from test_gpu_parity import _check_all

pytestmark = pytest.mark.gpu

DEFAULTS = {"wide": 1, "queue": 2, "first_stage": 3, "replicas": 1, "domain": 0, "tile_bytes": 1024,
            "warps": 0, "split": 1, "initial_ring": 1 << 20}
# every variant pins first_stage: 1 = hash table, 2 = per-byte table (FDR sets only; Teddy and
# noodle sets use their per-byte tables whatever this says)
BASE = {"first_stage": 1, "wide": 0, "split": 0}
VARIANTS = [
    {},                                                 # 16-byte lanes, queue for byte tables only (round-1 default)
    {"wide": 1},
    {"wide": 1, "domain": 12, "replicas": 8},
    {"wide": 1, "tile_bytes": 4096, "warps": 5},
    {"queue": 1},
    {"queue": 1, "first_stage": 2},
    {"wide": 1, "first_stage": 2},
    {"wide": 1, "split": 1},
    {"wide": 1, "split": 1, "domain": 12, "replicas": 8},
    {"wide": 1, "split": 1, "initial_ring": 16},       # candidate list and ring overflow -> grow and rescan
    {"wide": 1, "split": 1, "warps": 24},               # the 768-thread build of the kernel
    {"wide": 1, "split": 1, "warps": 32},               # the 1024-thread build
]


@pytest.mark.parametrize("opts", VARIANTS, ids=[",".join("%s=%d" % kv for kv in v.items()) for v in VARIANTS])
@pytest.mark.parametrize("nlits", [1, 30, 48, 300, 1000])
def test_variant_parity(hs, ref, opts, nlits):
    lits, flags, ids = synth.literal_set(nlits, min_len=2 if nlits < 40 else 4, max_len=12, seed=nlits + 7,
                                         caseless_frac=0.2, alphabet=b"abcdefgh")
    data, off, ln = synth.ragged_corpus([0, 1, 3, 15, 16, 17, 31, 32, 33, 100, 511, 512, 513, 1023, 1024, 1025,
                                         2047, 2048, 2049, 4096, 10000, 65536 + 5, 200000], lits, seed=3,
                                        alphabet=b"abcdefghABCDxy")
    try:
        for k, v in dict(BASE, **opts).items():
            hs.set_runtime_option(k, v)
        _check_all(hs, ref, lits, flags, ids, data, off, ln, use_brute=False)
    finally:
        for k, v in DEFAULTS.items():
            hs.set_runtime_option(k, v)


@pytest.mark.parametrize("opts", VARIANTS[1:4], ids=["wide", "wide-d12x8", "wide-geom"])
def test_variant_config2_sample(hs, ref, opts):
    lits, flags, ids = synth.literal_set(1000)
    data, off, ln, _ = synth.block_corpus(4096, 1024, lits, plant_per_kb=0.05, seed=9)
    try:
        for k, v in dict(BASE, **opts).items():
            hs.set_runtime_option(k, v)
        _check_all(hs, ref, lits, flags, ids, data, off, ln, use_brute=False)
    finally:
        for k, v in DEFAULTS.items():
            hs.set_runtime_option(k, v)
