"""-m gpu: HBM-resident shard API — device corpus == host twin, shard ownership/halo, size-independent
properties at larger sizes."""
import ctypes as C

import pytest
import torch

import gpu_util as gu
import oracle_util as ou
from krep_b200 import lib
from krep_b200.abi import (ALGO_AC, ALGO_BMH, ALGO_SSE42, CORPUS_EMBED_HALF, CORPUS_RANDOM_CASE, Params)

pytestmark = pytest.mark.gpu
NEEDLE = b"qzXv9Kpw"


def checker():
    return ou.reference() or ou.port()


def test_device_corpus_equals_host_twin():
    for flags, needle in [(0, NEEDLE), (CORPUS_RANDOM_CASE, b"QzXv"), (CORPUS_EMBED_HALF, b"needleneedle0016")]:
        spec = lib.make_spec(0x5EED0001, 0x5EED0002, 4096, needle, flags)
        for off, n in [(0, 100_000), (4096 * 7 + 16, 33_333), (1 << 33, 70_001)]:
            dev = gu.device_corpus(spec, off, n)
            assert bytes(dev[:n].cpu().numpy()) == lib.corpus_host(spec, off, n)


@pytest.mark.parametrize("pat,algo,func,opts,flags", [
    (NEEDLE, ALGO_SSE42, "sse42", {}, 0),
    (b"QzXv", ALGO_BMH, "boyer_moore", dict(case_sensitive=False), CORPUS_RANDOM_CASE),
    (b"needleneedle0016", ALGO_SSE42, "sse42", dict(whole_word=True), CORPUS_EMBED_HALF),
    (b"the", ALGO_BMH, "boyer_moore", {}, 0),
])
def test_whole_shard_matches_oracle_on_corpus(pat, algo, func, opts, flags):
    L = lib.load()
    n = 32 * (1 << 20) + 77
    spec = lib.make_spec(7, 8, 1 << 16, pat if len(pat) > 3 else NEEDLE, flags)
    dev = gu.device_corpus(spec, 0, n)
    host = bytes(dev[:n].cpu().numpy())
    p = Params(pat, **opts)
    plan = L.krep_b200_plan_create(p.ref(), algo)
    lib.check(L)
    try:
        out = gu.scan(plan, dev, n)
        got = gu.collect(plan, p, out)
        want = checker().run(func, Params(pat, **opts), host)
        assert got == want
        assert got[0] >= n // (1 << 16) // 2 - 2
        # count-only launch gives the same count
        out2 = gu.scan(plan, dev, n, want_positions=False)
        assert out2.count == out.count
    finally:
        L.krep_b200_plan_destroy(plan)


def test_shards_with_halo_union_equals_whole():
    """each shard owns matches by start offset, reads a halo, and takes -w context from its neighbours."""
    L = lib.load()
    n = 8 * (1 << 20) + 1234
    pat = b"needleneedle0016"
    spec = lib.make_spec(21, 22, 1 << 14, pat, CORPUS_EMBED_HALF)
    whole = gu.device_corpus(spec, 0, n)
    host = bytes(whole[:n].cpu().numpy())
    for opts, algo in [(dict(whole_word=True), ALGO_SSE42), (dict(), ALGO_BMH)]:
        p = Params(pat, **opts)
        plan = L.krep_b200_plan_create(p.ref(), algo)
        try:
            ref = gu.collect(plan, p, gu.scan(plan, whole, n))
            for nshards in (2, 3, 8):
                S = ((n + nshards - 1) // nshards + 15) // 16 * 16
                merged = []
                total = 0
                for g in range(nshards):
                    b, e = g * S, min((g + 1) * S, n)
                    avail = min(e + len(pat), n)          # halo: pattern_len bytes (occurrence + following byte)
                    shard = gu.device_corpus(spec, b, avail - b)
                    out = gu.scan(plan, shard, avail - b, own_begin=0, own_end=e - b, global_offset=b,
                                  prev_byte=host[b - 1] if b else -1, next_byte=host[avail] if avail < n else -1)
                    cnt, pos = gu.collect(plan, p, out)
                    total += cnt
                    merged += pos
                assert (total, merged) == ref, (opts, nshards)
        finally:
            L.krep_b200_plan_destroy(plan)


def test_properties_at_1gib():
    """Size-independent checks at a size the oracle would not finish quickly: sortedness, every reported
    offset really holds the needle, every intact plant is reported, count-only == list length,
    and the shard decomposition agrees with the single-shard result."""
    L = lib.load()
    n = 1 << 30
    period = 1 << 20
    spec = lib.make_spec(0x5EED0001, 0x5EED0002, period, NEEDLE)
    dev = gu.device_corpus(spec, 0, n)
    p = Params(NEEDLE)
    plan = L.krep_b200_plan_create(p.ref(), ALGO_SSE42)
    try:
        out = gu.scan(plan, dev, n)
        cnt, pos = gu.collect(plan, p, out)
        assert cnt == len(pos) >= n // period - 8
        starts = torch.tensor([s for s, _ in pos], dtype=torch.int64, device="cuda")
        assert bool((starts[1:] > starts[:-1]).all())
        idx = starts[:, None] + torch.arange(len(NEEDLE), device="cuda")[None, :]
        needle_t = torch.tensor(list(NEEDLE), dtype=torch.uint8, device="cuda")
        assert bool((dev[idx] == needle_t[None, :]).all())
        assert gu.scan(plan, dev, n, want_positions=False).count == cnt
        # two shards
        half = n // 2
        a = gu.scan(plan, dev, half + 16, own_begin=0, own_end=half, next_byte=-1)
        ca, pa = gu.collect(plan, p, a)
        b = gu.scan(plan, dev, n, own_begin=half, own_end=n)
        cb, pb = gu.collect(plan, p, b)
        assert pa + pb == pos and ca + cb == cnt
    finally:
        L.krep_b200_plan_destroy(plan)


def _texts_for_line_counting():
    import random
    rng = random.Random(31)
    words = [b"needle", b"the", b"quick", b"ab", b"abab", b"NEEDLE", b"x", b"haystack", b"aaa"]
    out = []
    for n, nl in [(200, 0.3), (5_000, 0.1), (300_000, 0.02), (300_000, 0.0005), (50_000, 0.0)]:
        t = bytearray()
        while len(t) < n:
            t += rng.choice(words) + (b"\n" if rng.random() < nl else rng.choice([b" ", b"", b"_", b", "]))
        out.append(bytes(t[:n]))
    out.append(b"\n" * 40 + b"needle\n\nneedle needle\n" + b"x" * 300 + b"needle")
    return out


@pytest.mark.parametrize("func,algo,pats,opts", [
    ("boyer_moore", ALGO_BMH, [b"needle"], dict()),
    ("boyer_moore", ALGO_BMH, [b"ab"], dict(whole_word=True)),
    ("boyer_moore", ALGO_BMH, [b"aaa"], dict(max_count=7)),
    ("sse42", ALGO_SSE42, [b"abab"], dict()),
    ("sse42", ALGO_SSE42, [b"the quick"], dict()),
    ("kmp", 1, [b"abab"], dict()),
    ("memchr", 2, [b"x"], dict()),
    ("memchr_short", 3, [b"ab"], dict(case_sensitive=False)),
    ("avx2", 5, [b"needle the quick ab"], dict()),
    ("aho_corasick", ALGO_AC, [b"needle", b"quick", b"ab", b"haystack"], dict()),
    ("aho_corasick", ALGO_AC, [b"needle", b"haystack", b"quick the"], dict(case_sensitive=False, max_count=5)),
])
def test_count_lines_on_the_device(func, algo, pats, opts):
    """-c for an HBM-resident shard: the scan computes every occurrence's line bounds on the GPU (k_line_bounds) and the
    replay counts lines from them — no host copy of the text is involved.  Must equal the reference's -c result."""
    L = lib.load()
    for text in _texts_for_line_counting():
        p = Params(pats, count=True, **opts)
        if func == "aho_corasick":
            p.struct.ac_trie = 1
        plan = L.krep_b200_plan_create(p.ref(), algo)
        lib.check(L)
        try:
            dev = gu.to_device(text)
            out = gu.scan(plan, dev, len(text))
            got = gu.collect(plan, p, out)
            p.struct.ac_trie = None
            want = checker().run(func, Params(pats, count=True, **opts), text)
            assert got == want, (func, pats, opts, len(text), got, want)
        finally:
            p.struct.ac_trie = None
            L.krep_b200_plan_destroy(plan)


def test_properties_at_baseline_size_10gib():
    """BASELINE configs[1] and [3] at their full single-GPU size (10 GiB resident): size-independent properties.
    Literal: strict sortedness, every reported offset holds the needle, every intact plant is reported, count-only ==
    list length, 4-shard decomposition == single shard.  Pattern set: every reported (start, end) spells a pattern of
    the set, and the count equals the SUM of the single-literal counts of all 1000 patterns (the reference's own
    AC == sum-of-BMH check, test/test_multiple_patterns.c:345-466, at full size)."""
    import bench
    L = lib.load()
    n = 10 * (1 << 30)
    wl = bench.WORKLOADS["literal8"]
    spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, wl["period"], wl["needle"], wl["flags"])
    dev = gu.device_corpus(spec, 0, n)
    needle = wl["needle"]
    p = Params(needle)
    plan = L.krep_b200_plan_create(p.ref(), ALGO_SSE42)
    try:
        out = gu.scan(plan, dev, n)
        cnt, pos = gu.collect(plan, p, out)
        assert cnt == len(pos) == n // wl["period"]          # one plant per period, none lost, no accidental 8-byte hit
        starts = torch.tensor([s for s, _ in pos], dtype=torch.int64, device="cuda")
        assert bool((starts[1:] > starts[:-1]).all())
        idx = starts[:, None] + torch.arange(len(needle), device="cuda")[None, :]
        assert bool((dev[idx] == torch.tensor(list(needle), dtype=torch.uint8, device="cuda")[None, :]).all())
        assert gu.scan(plan, dev, n, want_positions=False).count == cnt
        merged, q = [], n // 4
        for g in range(4):
            b, e = g * q, (g + 1) * q if g < 3 else n
            o = gu.scan(plan, dev, min(e + 16, n), own_begin=b, own_end=e)
            c, ps = gu.collect(plan, p, o)
            merged += ps
        assert merged == pos
    finally:
        L.krep_b200_plan_destroy(plan)
    del dev
    torch.cuda.empty_cache()

    wl = bench.WORKLOADS["multi1000"]
    pats = bench.multi_patterns(wl["multi"], wl["needle"])
    spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, wl["period"], wl["needle"], wl["flags"])
    dev = gu.device_corpus(spec, 0, n)
    pm = Params(pats)
    pm.struct.ac_trie = 1
    plan = L.krep_b200_plan_create(pm.ref(), ALGO_AC)
    try:
        out = gu.scan(plan, dev, n)
        cnt, pos = gu.collect(plan, pm, out)
        assert cnt == len(pos) >= n // wl["period"]
        pset = set(pats)
        starts = torch.tensor([s for s, _ in pos], dtype=torch.int64, device="cuda")
        win = dev[starts[:, None] + torch.arange(12, device="cuda")[None, :]].cpu().numpy()
        for (s, e), row in zip(pos, win):
            assert bytes(row[: e - s]) in pset, (s, e)
        ends = [e for _, e in pos]
        assert ends == sorted(ends)                           # emission order: ascending end offset
        total = 0
        for k, pat in enumerate(pats):
            pk = Params(pat)
            plk = L.krep_b200_plan_create(pk.ref(), ALGO_BMH)
            total += gu.scan(plk, dev, n, want_positions=False).count
            L.krep_b200_plan_destroy(plk)
        assert total == cnt                                   # AC == sum over patterns of the single-literal counts
    finally:
        pm.struct.ac_trie = None
        L.krep_b200_plan_destroy(plan)


def test_shard_larger_than_32gib_pattern_set_and_literal():
    """A 34 GiB resident shard: offsets beyond 2^35, and the pattern-set kernel's queue entries hold 32-bit relative
    group indices, so the host splits such a shard into launches of at most 2^31 groups (32 GiB).  The single-shard
    result must equal the concatenation of three sub-shards' results, and every match must spell a pattern."""
    import bench
    L = lib.load()
    n = 34 * (1 << 30) + 48
    wl = bench.WORKLOADS["multi1000"]
    pats = bench.multi_patterns(wl["multi"], wl["needle"])
    spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, wl["period"], wl["needle"], wl["flags"])
    dev = gu.device_corpus(spec, 0, n)
    pm = Params(pats)
    pm.struct.ac_trie = 1
    plan = L.krep_b200_plan_create(pm.ref(), ALGO_AC)
    pl = Params(wl["needle"])
    plan_lit = L.krep_b200_plan_create(pl.ref(), ALGO_SSE42)
    try:
        cnt, pos = gu.collect(plan, pm, gu.scan(plan, dev, n))
        assert cnt == len(pos) >= n // wl["period"]
        assert pos[-1][0] > 33 * (1 << 30)
        merged = []
        cuts = [0, 12 * (1 << 30), 24 * (1 << 30) + 16, n]
        for b, e in zip(cuts[:-1], cuts[1:]):
            o = gu.scan(plan, dev, min(e + 32, n), own_begin=b, own_end=e)
            merged += gu.collect(plan, pm, o)[1]
        assert merged == pos
        pset = set(pats)
        tail = [p for p in pos if p[0] > 31 * (1 << 30)][:2000]     # beyond the 2^31-group launch boundary
        starts = torch.tensor([s for s, _ in tail], dtype=torch.int64, device="cuda")
        win = dev[starts[:, None] + torch.arange(12, device="cuda")[None, :]].cpu().numpy()
        for (s, e), row in zip(tail, win):
            assert bytes(row[: e - s]) in pset, (s, e)
        # the planted needle through the literal kernel: same occurrences as the pattern set reports for it
        cl, posl = gu.collect(plan_lit, pl, gu.scan(plan_lit, dev, n))
        assert posl == [p for p in pos if p[1] - p[0] == len(wl["needle"]) and p in set(posl)] and cl >= n // wl["period"]
    finally:
        pm.struct.ac_trie = None
        L.krep_b200_plan_destroy(plan)
        L.krep_b200_plan_destroy(plan_lit)
