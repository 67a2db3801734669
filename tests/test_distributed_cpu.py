"""The N>1 path on CPU: world_size-2 gloo run of the sharded-scan exchange
(hyperscan_b200/dist.py): every rank 'scans' its shard (here: the reference
runtime stands in for the kernel, raw records shuffled and duplicated), the
records are all-gathered and merged, and the result equals one scan of the
whole corpus."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hyperscan_b200 import capi, dist as hd, synth
    import oracle.ref as ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lits, flags, ids = synth.literal_set(60, min_len=3, max_len=10, seed=5, alphabet=b"abcdef",
                                         singlematch_frac=0.2)
    db = capi.compile_lit_multi(lits, flags, ids)
    nblocks = 64
    data, off, ln, _ = synth.block_corpus(nblocks * world, 512, lits, plant_per_kb=6, seed=1)
    lo, hi = rank * nblocks, (rank + 1) * nblocks          # contiguous block shards
    # rank-local scan: blocks renumbered from 0, as the device numbers them
    shard = data[int(off[lo]):int(off[hi - 1]) + int(ln[hi - 1])]
    local, _ = ref.scan_collect(db.ptr, shard, off[lo:hi] - off[lo], ln[lo:hi])
    rng = np.random.default_rng(rank)
    raw = np.concatenate([local, local[: local.size // 3]])   # duplicates, like un-deduped records
    raw = raw[rng.permutation(raw.size)]
    words = hd.records_to_words(raw)
    counts, gathered = hd.all_gather_records(words, raw.size, pad_to=64)
    merged = hd.merge_gathered(counts, gathered, [r * nblocks for r in range(world)])
    final = capi.postprocess_matches(db, merged)
    # the fused single-collective form (what bench.py uses over NCCL)
    import torch
    buf = torch.zeros((raw.size + 8 + 1, 2), dtype=torch.int64)
    buf[1:raw.size + 1] = words
    res = hd.all_gather_records_fused(buf, raw.size)
    assert res is not None
    merged2 = hd.merge_gathered(res[0], res[1], [r * nblocks for r in range(world)])
    assert np.array_equal(capi.postprocess_matches(db, merged2), final)
    small = torch.zeros((2, 2), dtype=torch.int64)
    assert hd.all_gather_records_fused(small, raw.size) is None   # overflow is reported
    if rank == 0:
        want = ref.scan_sorted(db.ptr, data, off, ln)
        q.put((bool(np.array_equal(final, want)), int(final.size), counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_merge(hs, ref):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n, counts = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok and n > 50 and len(counts) == 2 and min(counts) > 0
