"""Multi-GPU plumbing for the scan path: shard geometry and the single gather of counts/offsets.

This is the analogue of krep's chunker + merge (krep.c:2816-2905, 2928-3004) with two differences that make
the result equal to the reference's single-chunk run instead of its multi-thread artefacts:
a match belongs to the shard that contains its START, and -w context bytes come from the neighbouring shards.
Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


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


def gather_keys(local_keys, world, rank, device):
    """One all_gather of per-rank counts + one gather of the (padded) sorted key lists to rank 0.

    local_keys: 1-D int64 tensor on `device` (sorted, global offsets).  Returns the concatenated, globally
    ascending key tensor on rank 0 (CPU), None elsewhere, plus the list of per-rank counts."""
    cnt = torch.tensor([local_keys.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    mine = torch.zeros(mx, dtype=torch.int64, device=device)
    mine[: local_keys.numel()] = local_keys
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank != 0:
        return None, counts
    # shards are disjoint and ordered by rank, so concatenation in rank order is globally sorted
    return torch.cat([g[:c] for g, c in zip(gathered, counts)]).cpu(), counts


class KeyGatherer:
    """The same exchange with persistent buffers and ONE collective per step, for the benchmark's steady state:
    every rank all_gathers a fixed-capacity row [count, key_0 .. key_{cap-1}]; rank 0 reads the rows back in one
    device-to-host copy and concatenates the valid prefixes (rank order = global order, see gather_keys).  If any
    rank's count exceeds the capacity every rank sees it in the gathered counts, grows its buffers to the same new
    capacity and the step is repeated — capacity is a pure function of the gathered counts, so ranks never disagree."""

    def __init__(self, world, rank, device, capacity=1 << 14):
        self.world, self.rank, self.device = world, rank, device
        self._alloc(capacity)

    def _alloc(self, capacity):
        self.cap = int(capacity)
        self.row = torch.zeros(self.cap + 1, dtype=torch.int64, device=self.device)
        self.all = torch.zeros((self.world, self.cap + 1), dtype=torch.int64, device=self.device)
        pin = self.device != "cpu" and torch.cuda.is_available()
        self.host = torch.zeros((self.world, self.cap + 1), dtype=torch.int64, pin_memory=pin) if self.rank == 0 else None
        self.host_counts = torch.zeros(self.world, dtype=torch.int64, pin_memory=pin)

    def key_buffer(self):
        """Device buffer the rank's sorted keys are written into (row[1:], at most `cap` keys)."""
        return self.row[1:]

    def exchange(self, count, check=True):
        """`count` keys are in key_buffer() (or count > cap and the caller will be told to retry).
        -> (keys on host as a 1-D int64 tensor on rank 0 else None, counts list or None, retry flag).

        check=True: every rank reads the gathered counts back (one small synchronising copy) so that all ranks agree
        on growing the capacity.  check=False is for a steady state whose counts are known to fit (e.g. after warm-up
        steps ran with check=True): ranks other than 0 then issue the collective and return without synchronising, and
        rank 0 raises if a count does not fit after all."""
        self.row[0] = int(count)
        dist.all_gather(list(self.all.unbind(0)), self.row)
        if check:
            self.host_counts.copy_(self.all[:, 0], non_blocking=False)
            counts = [int(c) for c in self.host_counts.tolist()]
            if max(counts) > self.cap:
                cap = self.cap
                while cap < max(counts):
                    cap *= 2
                self._alloc(cap * 2)
                return None, counts, True
            if self.rank != 0:
                return None, counts, False
            self.host.copy_(self.all, non_blocking=False)
        else:
            if self.rank != 0:
                return None, None, False
            self.host.copy_(self.all, non_blocking=False)   # one copy brings counts and keys
            counts = [int(c) for c in self.host[:, 0].tolist()]
            if max(counts) > self.cap:
                raise RuntimeError(f"KeyGatherer: count {max(counts)} exceeds capacity {self.cap} in an unchecked exchange")
        return torch.cat([self.host[r, 1:1 + c] for r, c in enumerate(counts)]), counts, False
