"""CPU, world_size 2 and 3 over gloo: the N>1 host path — shard geometry (ownership by match start, halo,
-w context from the neighbour), the count all_gather + key gather to rank 0, and the policy replay over the
concatenated list — gives exactly the oracle's single-chunk answer, i.e. none of the reference's multi-thread
artefacts.  Per-shard occurrence keys are produced in Python exactly as the device emits them."""
import ctypes as C
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util as ou
from krep_b200 import lib, sharding
from krep_b200.abi import ALGO_AC, ALGO_BMH, ALGO_SSE42, Params
from test_replay import device_like_keys


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# the verdict's repro: a long pattern that starts just before a shard cut ends AFTER a short pattern that starts behind
# the cut -> per-rank lists concatenated in rank order are not in aho_corasick_search's emission order (end ascending)
STRADDLE = b"abcdefghij"


def straddle_text(n, world):
    """n bytes of filler with 'abcdefghij' planted so that it starts 2 bytes before every shard cut."""
    t = bytearray(b"xy z\n" * (n // 5 + 1))[:n]
    for r in range(1, world):
        cut, _, _ = sharding.shard_bounds(n, world, r, 11)
        for s in (cut - 2, cut - 9):           # straddling the cut; and wholly in the earlier rank's halo region
            if 0 <= s and s + len(STRADDLE) <= n:
                t[s:s + len(STRADDLE)] = STRADDLE
    return bytes(t)


CASES = [
    ("sse42", ALGO_SSE42, [b"needle"], dict()),
    ("boyer_moore", ALGO_BMH, [b"abab"], dict()),                       # overlapping occurrences across the cut
    ("sse42", ALGO_SSE42, [b"abab"], dict()),                           # greedy non-overlap must be global, not per shard
    ("boyer_moore", ALGO_BMH, [b"needle"], dict(whole_word=True)),      # -w context across the cut
    ("boyer_moore", ALGO_BMH, [b"NeEdLe"], dict(case_sensitive=False, max_count=3)),
    ("aho_corasick", ALGO_AC, [b"ab", b"abcdefgh", b"needle", b"dle x"], dict()),   # short pattern inside the halo: no duplicate
    # nested short pattern behind the cut, long pattern across it (merge by key, not concatenation)
    ("aho_corasick", ALGO_AC, [STRADDLE, b"cde"], dict(text="straddle")),
    ("aho_corasick", ALGO_AC, [STRADDLE, b"cde"], dict(text="straddle", max_count=1)),      # -m 1 keeps the FIRST emission
    ("aho_corasick", ALGO_AC, [STRADDLE, b"cde", b"hij", b"j"], dict(text="straddle", max_count=3)),
    ("aho_corasick", ALGO_AC, [STRADDLE, b"cde", b"cde", b"efg"], dict(text="straddle")),   # duplicate patterns: one emission per index
    ("aho_corasick", ALGO_AC, [STRADDLE, b"cde"], dict(text="straddle", count=True)),       # -c
    ("aho_corasick", ALGO_AC, [b"ABCDEFGHIJ", b"Cde", b"z"], dict(text="straddle", case_sensitive=False, whole_word=True)),
]


def make_text(seed, n):
    rng = random.Random(seed)
    words = [b"needle", b"xneedle", b"needlex", b"abababab", b"abcdefgh", b"the", b"NEEDLE", b" "]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + rng.choice([b" ", b"", b"\n", b"_"])
    return bytes(out[:n])


def shard_keys(func, pats, opts, text, begin, own_len, avail_len):
    """Keys a rank's device scan would report: occurrences that START in the owned range, found in the bytes the
    shard can see, -w judged against the global neighbours (prev/next byte passed as shard context)."""
    view = text[begin:begin + avail_len]
    prev_b = text[begin - 1:begin] if begin > 0 else b""
    next_b = text[begin + avail_len:begin + avail_len + 1]
    padded = prev_b + view + next_b                 # context bytes only influence the -w test
    off = len(prev_b)
    keys = device_like_keys(func, pats, padded, opts.get("case_sensitive", True), opts.get("whole_word", False), False)
    out = []
    for k in keys:
        if func == "aho_corasick":
            end = (k >> 24) - off
            ln = 1024 - ((k >> 14) & 1023)
            start = end - ln
            if 0 <= start < own_len and end <= avail_len:
                out.append(((end + begin) << 24) | (k & 0xFFFFFF))
        else:
            start = (k >> 3) - off
            m = len(pats[0])
            if 0 <= start < own_len and start + m <= avail_len:
                out.append(((start + begin) << 3) | (k & 7))
    return out


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = lib.load()
    ok = True
    gatherer = sharding.KeyGatherer(world, rank, "cpu", capacity=4)  # tiny: the grow-and-retry path runs too
    try:
        for ci, (func, algo, pats, opts) in enumerate(CASES):
            opts = dict(opts)
            kind = opts.pop("text", "words")
            for n in (1000, 4099):
                text = straddle_text(n, world) if kind == "straddle" else make_text(100 + ci, n)
                halo = max(map(len, pats)) + 1
                begin, own, avail = sharding.shard_bounds(n, world, rank, halo)
                keys = shard_keys(func, pats, opts, text, begin, own, avail)
                allk, counts = sharding.gather_keys(torch.tensor(keys, dtype=torch.int64), world, rank, "cpu")
                # the single-collective exchange used by bench.py must deliver the same list
                while not gatherer.negotiate(len(keys)):
                    pass
                gatherer.row[0] = len(keys)
                gatherer.row[1:1 + len(keys)] = torch.tensor(keys, dtype=torch.int64)
                gatherer.post(ci & 1)
                if rank == 0:
                    allk2, counts2 = gatherer.fetch(ci & 1)
                    if counts2 != counts or not torch.equal(allk2, allk):
                        ok = False
                        q.put(("gatherer mismatch", func, n, world))
                    if allk.numel() > 1 and not bool((allk[1:] >= allk[:-1]).all()):
                        ok = False
                        q.put(("merged list not ascending", func, n, world))
                    p = Params(pats, **opts)
                    if func == "aho_corasick":
                        p.struct.ac_trie = 1
                    arr = (C.c_uint64 * max(allk.numel(), 1))(*[int(x) & 0xFFFFFFFFFFFFFFFF for x in allk.tolist()])
                    res = L.krep_b200_match_result_init(16)
                    cnt = L.krep_b200_replay(algo, p.ref(), False, arr, allk.numel(), text, len(text), res)
                    r = res.contents
                    got = (int(cnt), [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)])
                    L.krep_b200_match_result_free(res)
                    want = ou.port().run(func, Params(pats, **opts), text)
                    if got != want:
                        ok = False
                        q.put(("mismatch", func, pats, opts, n, world, got[0], want[0]))
                    if kind == "straddle" and set(opts) <= {"case_sensitive"} and want[0] < 2:
                        ok = False
                        q.put(("straddle case did not produce the nested pair", pats, n, world, want))
    finally:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        q.put(("done", ok))


def test_merge_keys_restores_emission_order():
    """The repro from the round-1 review, without any process group: two shards of a 64-byte text."""
    text = bytearray(b"." * 64)
    text[30:40] = STRADDLE                      # starts 2 bytes before the cut at 32
    text = bytes(text)
    pats = [STRADDLE, b"cde"]
    lists = []
    for rank in range(2):
        b, own, avail = sharding.shard_bounds(64, 2, rank, 11)
        lists.append(shard_keys("aho_corasick", pats, {}, text, b, own, avail))
    assert lists[0] and lists[1] and lists[0][-1] > lists[1][0]          # concatenation is NOT ascending
    cap = max(map(len, lists))
    rows = torch.zeros((2, cap + 1), dtype=torch.int64)
    for r, ks in enumerate(lists):
        rows[r, 0] = len(ks)
        rows[r, 1:1 + len(ks)] = torch.tensor(ks, dtype=torch.int64)
    merged = sharding.merge_rows(rows, [len(x) for x in lists]).tolist()
    assert merged == sorted(lists[0] + lists[1])
    L = lib.load()
    for maxc in (2 ** 64 - 1, 1):
        p = Params(pats, max_count=maxc)
        p.struct.ac_trie = 1
        arr = (C.c_uint64 * len(merged))(*merged)
        res = L.krep_b200_match_result_init(16)
        cnt = L.krep_b200_replay(ALGO_AC, p.ref(), False, arr, len(merged), text, len(text), res)
        got = (int(cnt), [(res.contents.positions[i].start_offset, res.contents.positions[i].end_offset)
                          for i in range(res.contents.count)])
        L.krep_b200_match_result_free(res)
        assert got == ou.port().run("aho_corasick", Params(pats, max_count=maxc), text)
        assert got[1][0] == (32, 35)            # the short match ends first


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_replay_equals_single_chunk(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    msgs = []
    while not q.empty():
        msgs.append(q.get())
    assert ("done", True) in msgs, msgs


def test_shard_bounds_tile_exactly():
    for n in (0, 1, 15, 16, 17, 1000, 10 * (1 << 30) + 5):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                b, own, avail = sharding.shard_bounds(n, world, r, halo=9)
                assert b == pos and own <= avail <= own + 9 and b % 16 == 0 or own == 0
                pos = b + own
            assert pos == n
