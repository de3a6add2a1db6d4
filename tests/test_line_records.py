"""CPU: the monoid that folds per-shard -c records (csrc/scan_count.cu append_rec, exported as
krep_b200_combine_line_counts).  Shard records are computed here by brute force from the text exactly as the device
defines them; folding them in text order must give the oracle's -c count whatever the cuts."""
import ctypes as C
import random

import oracle_util as ou
from krep_b200 import lib
from krep_b200.abi import Params, SIZE_MAX

HAS_HIT, FIRST_OPEN, LAST_PENDING, HAS_NL = 1, 2, 4, 8


class LineCount(C.Structure):
    _fields_ = [("lines", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


def shard_record(text, pat, b, e):
    """(lines, flags) of the shard that owns starts and newlines in [b, e)."""
    hits = []
    s = text.find(pat, b)
    while 0 <= s < e:
        hits.append(s)
        s = text.find(pat, s + 1)
    nls = [i for i in range(b, e) if text[i] == 0x0A]
    if not hits:
        return 0, (HAS_NL if nls else 0)
    lines, last_line = 0, None
    for h in hits:
        line = sum(1 for q in nls if q < h)         # index of the line inside the shard
        if line != last_line:
            lines += 1
            last_line = line
    flags = HAS_HIT | HAS_NL
    if not any(q < hits[0] for q in nls):
        flags |= FIRST_OPEN
    if not any(q >= hits[-1] for q in nls):
        flags |= LAST_PENDING
    return lines, flags


def test_fold_of_shard_records_equals_single_chunk_count():
    L = lib.load()
    L.krep_b200_combine_line_counts.argtypes = [C.POINTER(LineCount), C.c_size_t, C.c_size_t]
    L.krep_b200_combine_line_counts.restype = C.c_uint64
    rng = random.Random(2)
    words = [b"the", b"then", b"x", b"other", b"quick"]
    for trial in range(300):
        nl = rng.choice([0.0, 0.02, 0.2, 0.6])
        t = bytearray()
        n = rng.choice([1, 5, 40, 300, 2000])
        while len(t) < n:
            t += rng.choice(words) + (b"\n" if rng.random() < nl else rng.choice([b" ", b""]))
        text = bytes(t[:n])
        want = ou.port().run("boyer_moore", Params(b"the", count=True), text)[0]
        for nsh in (1, 2, 3, 7, 16):
            cuts = sorted(rng.randrange(0, n + 1) for _ in range(nsh - 1))
            cuts = [0] + cuts + [n]
            recs = (LineCount * nsh)()
            for i in range(nsh):
                recs[i].lines, recs[i].flags = shard_record(text, b"the", cuts[i], cuts[i + 1])
            got = L.krep_b200_combine_line_counts(recs, nsh, SIZE_MAX)
            assert got == want, (text, cuts, [(r.lines, r.flags) for r in recs], got, want)
            assert L.krep_b200_combine_line_counts(recs, nsh, 2) == min(want, 2)


def test_merge_keys_is_a_stable_k_way_merge():
    """krep_b200_merge_keys on random ascending lists (empty lists, duplicates across lists, aliasing dst with the first
    list) == sorted concatenation."""
    import ctypes as C
    L = lib.load()
    rng = random.Random(9)
    for trial in range(300):
        k = rng.randint(1, 9)
        lists = [sorted(rng.randrange(0, 1 << rng.choice([8, 40, 62])) for _ in range(rng.choice([0, 0, 1, 5, 40, 300]))) for _ in range(k)]
        total = sum(map(len, lists))
        bufs = [(C.c_uint64 * max(len(x), 1))(*x) for x in lists]
        ptrs = (C.c_void_p * k)(*[C.addressof(b) for b in bufs])
        cnts = (C.c_uint64 * k)(*[len(x) for x in lists])
        dst = (C.c_uint64 * max(total, 1))()
        n = L.krep_b200_merge_keys(ptrs, cnts, k, C.cast(dst, C.c_void_p))
        assert n == total and list(dst[:total]) == sorted(sum(lists, []))
    # rows of one matrix, dst aliasing nothing: what sharding.merge_rows does
    import torch
    from krep_b200 import sharding
    rows = torch.zeros((3, 6), dtype=torch.int64)
    data = [[5, 9, 30], [1, 2], [7, 8, 100, 200]]
    for r, xs in enumerate(data):
        rows[r, 0] = len(xs)
        rows[r, 1:1 + len(xs)] = torch.tensor(xs)
    assert sharding.merge_rows(rows, [len(x) for x in data]).tolist() == sorted(sum(data, []))
