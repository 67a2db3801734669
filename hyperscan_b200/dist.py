"""Multi-GPU plumbing for the sharded block scan: blocks
shard by rank, the database is replicated, and the only exchange is one
all-gather of per-rank match records at the end of a scan.

Records travel as int64 pairs (the 16-byte hs_b200_match_t viewed as two
little-endian words: word0 = id | block << 32, word1 = to).  Works on any
torch.distributed backend: NCCL on the GPUs (bench.py), gloo on CPU (tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .capi import MATCH_DTYPE


def records_to_words(recs):
    """MATCH_DTYPE array -> int64 tensor [n, 2]."""
    a = np.ascontiguousarray(recs, dtype=MATCH_DTYPE)
    return torch.from_numpy(a.view(np.int64).reshape(-1, 2).copy())


def words_to_records(words):
    a = words.detach().cpu().contiguous().numpy().reshape(-1, 2)
    return a.view(MATCH_DTYPE).reshape(-1)


def all_gather_records(local_words, n_local, group=None, pad_to=4096):
    """All-gather a variable number of records per rank.

    local_words: int64 tensor [cap, 2] on the backend's device whose first
    n_local rows are valid.  Returns (counts [world] python ints, gathered
    int64 tensor [world, maxn, 2]); two collectives: counts, then records
    padded to the max count rounded up to `pad_to` (buffers are reused across
    steps by the caller when the rounded size is stable)."""
    world = dist.get_world_size(group)
    dev = local_words.device
    cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts_h = [int(x) for x in counts.cpu().tolist()]
    maxn = max(max(counts_h), 1)
    maxn = (maxn + pad_to - 1) // pad_to * pad_to
    mine = torch.zeros((maxn, 2), dtype=torch.int64, device=dev)
    k = min(n_local, local_words.shape[0])
    mine[:k] = local_words[:k]
    out = torch.empty((world, maxn, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    return counts_h, out


def all_gather_records_fused(buf, n_local, group=None):
    """One collective per scan: `buf` is an int64 tensor [cap + 1, 2] whose rows
    1.. hold this rank's records; row 0 is set to (n_local, 0) here and travels
    with them.  Returns (counts, gathered [world, cap, 2]) or None when some
    rank's count exceeds cap (the caller grows `buf` on every rank and repeats)."""
    out = all_gather_records_fused_async(buf, n_local, group)
    return fused_result(out, buf.shape[0] - 1)


def all_gather_records_fused_async(buf, n_local, group=None):
    """Enqueue the fused all-gather and return the gathered tensor without
    waiting for it (no host synchronisation); pass it to fused_result() later."""
    world = dist.get_world_size(group)
    buf[0, 0] = n_local
    out = torch.empty((world,) + tuple(buf.shape), dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(out.view(-1), buf.view(-1), group=group)
    return out


def fused_result(out, cap):
    """(counts, records [world, cap, 2]) of a fused all-gather, or None if some
    rank had more than `cap` records (its tail was cut: grow and redo)."""
    counts = [int(x) for x in out[:, 0, 0].cpu().tolist()]
    if max(counts) > cap:
        return None
    return counts, out[:, 1:, :]


def merge_gathered(counts, gathered, block_base):
    """Concatenate every rank's valid records, renumbering rank-local block
    indices into the global numbering (block_base[r] = first global block of
    rank r).  Returns a MATCH_DTYPE array (unsorted)."""
    parts = []
    for r, n in enumerate(counts):
        if n == 0:
            continue
        rec = words_to_records(gathered[r, :n]).copy()
        rec["block"] += np.uint32(block_base[r])
        parts.append(rec)
    if not parts:
        return np.zeros(0, dtype=MATCH_DTYPE)
    return np.concatenate(parts)


class PeerExchange:
    """Fused scan + all-gather over NVLink peer memory (include/hs_b200.h,
    hs_b200_set_peer_exchange): every rank owns a buffer of world x (cap + 1)
    records; the scan kernel of rank r stores its records into slot row r of
    every rank's buffer as it finds them.  Handles travel once through the
    process group at set-up; the data path has no collective."""

    def __init__(self, cap, group=None):
        import ctypes as C
        from . import capi
        self.capi = capi
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap = int(cap)
        self.bytes = self.world * (self.cap + 1) * 16
        L = capi.lib()
        self.mine = C.c_void_p()
        handle = C.create_string_buffer(64)
        capi._check(L.hs_b200_peer_buffer_alloc(self.bytes, C.byref(self.mine), handle), "peer_buffer_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.bases = (C.c_void_p * self.world)()
        self.opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases[r] = self.mine.value
            else:
                p = C.c_void_p()
                capi._check(L.hs_b200_peer_buffer_open(h, C.byref(p)), "peer_buffer_open rank %d" % r)
                self.bases[r] = p.value
                self.opened.append(p)

    def attach(self, scratch, block_base):
        self.capi._check(self.capi.lib().hs_b200_set_peer_exchange(
            scratch.ptr, self.world, self.rank, self.bases, self.cap, int(block_base)), "set_peer_exchange")

    def detach(self, scratch):
        self.capi.lib().hs_b200_set_peer_exchange(scratch.ptr, 0, 0, None, 0, 0)

    def read(self):
        """(counts per rank, merged MATCH_DTYPE array of the valid records) from
        THIS rank's buffer (block indices are already global)."""
        raw = np.zeros(self.world * (self.cap + 1), dtype=MATCH_DTYPE)
        self.capi._check(self.capi.lib().hs_b200_peer_buffer_read(self.mine, raw.ctypes.data, self.bytes))
        raw = raw.reshape(self.world, self.cap + 1)
        counts = [int(raw[r, 0]["id"]) for r in range(self.world)]
        if max(counts) > self.cap:
            raise RuntimeError("peer exchange buffer overflowed: a rank published %d records, capacity %d"
                               % (max(counts), self.cap))
        parts = [raw[r, 1:1 + counts[r]] for r in range(self.world)]
        return counts, np.concatenate(parts) if parts else np.zeros(0, dtype=MATCH_DTYPE)

    def close(self):
        L = self.capi.lib()
        for p in self.opened:
            L.hs_b200_peer_buffer_close(p, 1)
        self.opened = []
        if self.mine:
            L.hs_b200_peer_buffer_close(self.mine, 0)
            self.mine = None
