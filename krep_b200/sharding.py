"""Multi-GPU plumbing for the scan path (one process per GPU): shard geometry and the single gather of per-shard
counts + occurrence keys to rank 0, followed by the key merge.

This is the analogue of krep's chunker + merge (krep.c:2816-2905, 2928-3004) with two differences that make
the result equal to the reference's single-chunk run instead of its multi-thread artefacts:
a match belongs to the shard that contains its START, and -w context bytes come from the neighbouring shards.

Order of the gathered lists.  Literal keys are ordered AND owned by start offset, so the per-rank lists concatenate
into a globally ascending list.  Pattern-set keys are ordered by END offset (aho_corasick_search's emission order,
aho_corasick.c:353-431) but still owned by start offset: a long match that starts just before a cut belongs to the
earlier rank yet ends after a short match owned by the later rank.  Rank 0 therefore always MERGES the per-rank lists
by key (krep_b200_merge_keys, C); the disorder is confined to max_pattern_len bytes around each cut, so the merge is
one linear pass per cut.

Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib


def shard_bounds(total_len, world, rank, halo, align=16):
    """-> (begin, own_len, avail_len): the shard owns [begin, begin+own_len) and can read avail_len bytes from begin.

    own ranges tile [0, total_len) exactly; every shard but the last reads `halo` bytes past its owned range
    (longest pattern + 1, so an occurrence starting on the last owned byte and the byte after it are visible)."""
    per = -(-total_len // world)
    per = -(-per // align) * align
    begin = min(rank * per, total_len)
    end = min(begin + per, total_len)
    avail_end = min(end + halo, total_len)
    return begin, end - begin, avail_end - begin


def merge_rows(rows, counts, out=None):
    """rows: 2-D int64 CPU tensor, row r = [count_r, key_0 .. ] (contiguous); counts: list of valid keys per row.
    -> 1-D int64 tensor with all keys in ascending key order (krep_b200_merge_keys)."""
    L = lib.load()
    world = rows.shape[0]
    total = int(sum(counts))
    if out is None or out.numel() < max(total, 1):
        out = torch.empty(max(total, 1), dtype=torch.int64)
    stride = rows.stride(0) * 8
    base = rows.data_ptr()
    lists = (C.c_void_p * world)(*[base + r * stride + 8 for r in range(world)])
    cnts = (C.c_uint64 * world)(*[int(c) for c in counts])
    n = L.krep_b200_merge_keys(lists, cnts, world, C.c_void_p(out.data_ptr()))
    assert n == total
    return out[:total]


def gather_keys(local_keys, world, rank, device):
    """One all_gather of per-rank counts + one gather of the (padded) sorted key lists to rank 0, merged by key.

    local_keys: 1-D int64 tensor on `device` (sorted, global offsets).  Returns the merged, globally ascending key
    tensor on rank 0 (CPU), None elsewhere, plus the list of per-rank counts."""
    cnt = torch.tensor([local_keys.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    mine = torch.zeros(mx + 1, dtype=torch.int64, device=device)
    mine[0] = local_keys.numel()
    mine[1: 1 + local_keys.numel()] = local_keys
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank != 0:
        return None, counts
    rows = torch.stack([g.cpu() for g in gathered]).contiguous()
    return merge_rows(rows, counts).clone(), counts


class KeyGatherer:
    """The same exchange with persistent buffers and ONE collective per step, for a steady state of many steps.

    Every rank owns a fixed-capacity row [count, key_0 .. key_{cap-1}] on its device (krep_b200_export_packed fills it
    in one device-to-device copy); post() gathers the rows to rank 0 — a gather, not an all_gather: ranks other than 0
    only send, so nothing on their GPUs waits for rank 0 — and rank 0 starts an asynchronous copy of the gathered rows
    into one of two pinned host buffers.  fetch() (rank 0) waits for that copy and merges the rows by key.  The two
    host buffers let rank 0 post step i+1 (and run its scan) before it fetches and replays step i.

    Capacity: negotiate() is the checked form used while warming up — every rank learns the largest count through a
    MAX all_reduce and all ranks grow their buffers in lockstep when it does not fit.  In the unchecked steady state a
    row that does not fit is truncated but still carries its exact count, so rank 0 notices; it does not raise in the
    middle of the run (the other ranks would hang in the next collective) but sets `overflowed`, which the caller
    turns into a collective error after the loop."""

    def __init__(self, world, rank, device, capacity=1 << 14):
        self.world, self.rank, self.device = world, rank, device
        self.overflowed = False
        self._alloc(capacity)

    def _alloc(self, capacity):
        self.cap = int(capacity)
        pin = self.device != "cpu" and torch.cuda.is_available()
        self.row = torch.zeros(self.cap + 1, dtype=torch.int64, device=self.device)
        if self.rank == 0:
            self.rows = torch.zeros((self.world, self.cap + 1), dtype=torch.int64, device=self.device)
            self.host = [torch.zeros((self.world, self.cap + 1), dtype=torch.int64, pin_memory=pin) for _ in range(2)]
            self.merged = torch.empty(self.world * self.cap + 1, dtype=torch.int64)
            self.events = [torch.cuda.Event() if pin else None for _ in range(2)]
        self._mx = torch.zeros(1, dtype=torch.int64, device=self.device)

    def row_ptr(self):
        """Device pointer of this rank's row (what krep_b200_export_packed writes: count + up to `cap` keys)."""
        return self.row.data_ptr()

    def negotiate(self, count):
        """Checked form: True if every rank's count fits, else grows the buffers on all ranks and returns False."""
        self._mx[0] = int(count)
        dist.all_reduce(self._mx, op=dist.ReduceOp.MAX)
        mx = int(self._mx.item())
        if mx <= self.cap:
            return True
        cap = self.cap
        while cap < mx:
            cap *= 2
        self._alloc(cap * 2)
        return False

    def post(self, slot=0):
        """Gather the rows to rank 0; rank 0 also starts the device-to-host copy into host buffer `slot`."""
        if self.rank == 0:
            dist.gather(self.row, list(self.rows.unbind(0)), dst=0)
            self.host[slot].copy_(self.rows, non_blocking=True)
            if self.events[slot] is not None:
                self.events[slot].record()
        else:
            dist.gather(self.row, None, dst=0)

    def fetch(self, slot=0):
        """Rank 0: -> (merged keys as a 1-D int64 CPU tensor, per-rank counts)."""
        assert self.rank == 0
        if self.events[slot] is not None:
            self.events[slot].synchronize()
        rows = self.host[slot]
        counts = [int(c) for c in rows[:, 0].tolist()]
        if max(counts) > self.cap:
            self.overflowed = True
            counts = [min(c, self.cap) for c in counts]
        return merge_rows(rows, counts, self.merged), counts
