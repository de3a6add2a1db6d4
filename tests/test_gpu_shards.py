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


# ------------------------------------------------------------------------------------------------
# round 2: cross-shard merge of pattern-set results with the REAL kernels per shard
# ------------------------------------------------------------------------------------------------
def _packed_keys(plan, out, cap=1 << 16):
    """[count, sorted keys] of a shard result, via krep_b200_export_packed (what a multi-GPU host gathers)."""
    L = lib.load()
    row = torch.zeros(cap + 1, dtype=torch.int64, device=f"cuda:{out.device}")
    assert L.krep_b200_export_packed(C.byref(out), row.data_ptr(), cap, None) == 0
    lib.check(L)
    host = row.cpu()
    cnt = int(host[0])
    assert cnt == out.count and cnt <= cap
    return host[:cnt + 1].contiguous()


def _merge_and_replay(algo, params, rows, text, only_matching=False):
    from krep_b200 import sharding
    L = lib.load()
    cap = max(r.numel() for r in rows)
    mat = torch.zeros((len(rows), cap), dtype=torch.int64)
    counts = []
    for i, r in enumerate(rows):
        mat[i, : r.numel()] = r
        counts.append(int(r[0]))
    merged = sharding.merge_rows(mat, counts)
    arr = C.cast(merged.data_ptr(), C.POINTER(C.c_uint64))
    res = L.krep_b200_match_result_init(16)
    try:
        cnt = L.krep_b200_replay(algo, params.ref(), only_matching, arr, merged.numel(), text, len(text), res)
        r = res.contents
        return int(cnt), [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)]
    finally:
        L.krep_b200_match_result_free(res)


def _bench_pattern_set_with_nested_pairs():
    """bench.py's 1000-pattern set (config 4) plus, for a few of its long patterns, an inner 6-byte substring as an
    extra pattern: a long match that starts before a cut then ends AFTER a short match that starts behind the cut."""
    import bench
    wl = bench.WORKLOADS["multi1000"]
    pats = bench.multi_patterns(wl["multi"], wl["needle"])
    longs = [p for p in pats if len(p) >= 11][:6]
    extra = [p[3:9] for p in longs] + [longs[0][5:11], longs[0]]          # nested, nested-at-the-end, and a duplicate
    return wl, pats + extra, longs


@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_pattern_set_shards_merge_equals_reference(nshards):
    """One corpus cut into 2 / 3 / 8 shards, the real multi-pattern kernel per shard (own buffer, global offset, context
    bytes), lists merged by key (krep_b200_merge_keys), replayed — must equal the compiled reference's single-chunk
    aho_corasick_search position for position, in order, with -m 1, -c and -w as well.  Adversarial plants: a long
    pattern starting 1..9 bytes before every cut with a nested short pattern behind the cut."""
    L = lib.load()
    wl, pats, longs = _bench_pattern_set_with_nested_pairs()
    n = 6 * (1 << 20) + 4242
    spec = lib.make_spec(77, 78, 1 << 16, wl["needle"], wl["flags"])
    host = bytearray(lib.corpus_host(spec, 0, n))
    S = ((n + nshards - 1) // nshards + 15) // 16 * 16
    for g in range(1, nshards):
        for k, delta in enumerate((1, 4, 9)):
            p = longs[(g + k) % len(longs)]
            s = g * S - delta - k * 40
            host[s - 1:s + len(p) + 1] = b" " + p + b" "
    host = bytes(host)
    halo = max(map(len, pats)) + 1
    for opts in (dict(), dict(max_count=1), dict(count=True), dict(whole_word=True), dict(case_sensitive=False, max_count=40)):
        p = Params(pats, **opts)
        p.struct.ac_trie = 1
        plan = L.krep_b200_plan_create(p.ref(), ALGO_AC)
        lib.check(L)
        try:
            rows, keep = [], []
            for g in range(nshards):
                b, e = g * S, min((g + 1) * S, n)
                avail = min(e + halo, n)
                shard = gu.to_device(host[b:avail])
                keep.append(shard)
                out = gu.scan(plan, shard, avail - b, own_begin=0, own_end=e - b, global_offset=b,
                              prev_byte=host[b - 1] if b else -1, next_byte=host[avail] if avail < n else -1)
                rows.append(_packed_keys(plan, out))
            cat = torch.cat([r[1:] for r in rows])
            if not opts:
                assert not bool((cat[1:] >= cat[:-1]).all()), "plants did not produce an out-of-order concatenation"
            got = _merge_and_replay(ALGO_AC, p, rows, host)
            p.struct.ac_trie = None
            want = checker().run("aho_corasick", Params(pats, **opts), host)
            assert got == want, (nshards, opts, got[0], want[0])
            assert want[0] > 0
        finally:
            p.struct.ac_trie = None
            L.krep_b200_plan_destroy(plan)


def test_host_search_cut_into_ranges_equals_reference(monkeypatch):
    """The search_func_t entry points on host text with the text cut into several ranges (one per device, or — here, on
    one GPU — KREP_B200_RANGES ranges taken one after the other): per-range lists merged by key inside the library."""
    wl, pats, longs = _bench_pattern_set_with_nested_pairs()
    n = 7 * (1 << 20) + 999
    spec = lib.make_spec(79, 80, 1 << 15, wl["needle"], wl["flags"])
    host = bytearray(lib.corpus_host(spec, 0, n))
    mb = 1 << 20
    for c in range(1, 7):
        for k, delta in enumerate((1, 5, 8)):
            p = longs[(c + k) % len(longs)]
            s = c * mb - delta - 30 * k
            host[s - 1:s + len(p) + 1] = b" " + p + b" "
    host = bytes(host)
    monkeypatch.setenv("KREP_B200_STAGE_MB", "1")
    monkeypatch.setenv("KREP_B200_CHUNK_MB", "1")
    for ranges in ("1", "3", "7"):
        monkeypatch.setenv("KREP_B200_RANGES", ranges)
        for opts in (dict(), dict(max_count=1), dict(count=True), dict(whole_word=True, case_sensitive=False)):
            got = lib.search("aho_corasick", Params(pats, **opts), host)
            want = checker().run("aho_corasick", Params(pats, **opts), host)
            assert got == want, (ranges, opts, got[0], want[0])
        for func, pat, opts in (("sse42", wl["needle"], {}), ("boyer_moore", b"the", dict(count=True)),
                                ("boyer_moore", b"et", dict(whole_word=True))):
            got = lib.search(func, Params(pat, **opts), host)
            want = checker().run(func, Params(pat, **opts), host)
            assert got == want, (ranges, func, opts)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_host_search_on_several_devices_equals_reference():
    """krep_b200_set_devices: one search call, the text spread over every visible GPU (single process)."""
    L = lib.load()
    wl, pats, longs = _bench_pattern_set_with_nested_pairs()
    n = 300 * (1 << 20) + 12345
    spec = lib.make_spec(81, 82, 1 << 18, wl["needle"], wl["flags"])
    host = lib.corpus_host(spec, 0, n)
    ndev = torch.cuda.device_count()
    devs = (C.c_int * ndev)(*range(ndev))
    try:
        L.krep_b200_set_devices(devs, ndev)
        for func, pp, opts in (("aho_corasick", pats, {}), ("sse42", wl["needle"], {}), ("boyer_moore", b"the", dict(count=True))):
            got = lib.search(func, Params(pp, **opts), host)
            L.krep_b200_set_devices(None, 0)
            one = lib.search(func, Params(pp, **opts), host)
            L.krep_b200_set_devices(devs, ndev)
            assert got == one and got[0] > 0, (func, got[0], one[0])
        small = host[: 9 * (1 << 20)]
        got = lib.search("aho_corasick", Params(pats), small)
        assert got == checker().run("aho_corasick", Params(pats), small)
    finally:
        L.krep_b200_set_devices(None, 0)


def test_scan_begin_end_on_a_user_stream_then_collect():
    """krep_b200_scan_shard_begin / _end, on a non-default stream: collect must be ordered after the device sort
    whatever stream the scan ran on (round-1 advisor finding), also for lists longer than the packed read-back."""
    from krep_b200.abi import DeviceResult, Shard
    L = lib.load()
    n = 24 * (1 << 20)
    spec = lib.make_spec(5, 6, 1 << 9, b"the", 0)          # one plant per 512 B -> ~49 k occurrences (> PACK_KEYS)
    dev = gu.device_corpus(spec, 0, n)
    host = bytes(dev[:n].cpu().numpy())
    st = torch.cuda.Stream()
    for pat, func, algo in ((b"the", "boyer_moore", ALGO_BMH), (NEEDLE, "sse42", ALGO_SSE42)):
        p = Params(pat)
        plan = L.krep_b200_plan_create(p.ref(), algo)
        try:
            sh = Shard(dev.data_ptr(), n, 0, n, 0, -1, -1)
            ticket = C.c_int(-1)
            assert L.krep_b200_scan_shard_begin(plan, C.byref(sh), 1, C.c_void_p(st.cuda_stream), C.byref(ticket)) == 0
            out = DeviceResult()
            assert L.krep_b200_scan_shard_end(ticket.value, C.byref(out)) == 0
            lib.check(L)
            got = gu.collect(plan, p, out)
            want = checker().run(func, Params(pat), host)
            assert got == want and (got[0] > 16384 or pat == NEEDLE)
        finally:
            L.krep_b200_plan_destroy(plan)


def test_count_lines_on_newline_aligned_shards():
    """-c on the device for shards that begin right after a newline and end right before one (prev_byte / next_byte are
    newlines): no line is cut, so every shard resolves its own line bounds and the line counts add up."""
    L = lib.load()
    rng_text = _texts_for_line_counting()[2]
    cuts = [0]
    for target in (len(rng_text) // 3, 2 * len(rng_text) // 3):
        cuts.append(rng_text.index(b"\n", target) + 1)
    cuts.append(len(rng_text))
    for pats, func, algo in (([b"needle"], "boyer_moore", ALGO_BMH), ([b"needle", b"quick", b"haystack"], "aho_corasick", ALGO_AC)):
        p = Params(pats, count=True)
        if func == "aho_corasick":
            p.struct.ac_trie = 1
        plan = L.krep_b200_plan_create(p.ref(), algo)
        try:
            total = 0
            for b, e in zip(cuts[:-1], cuts[1:]):
                piece = rng_text[b:e - 1] if e < len(rng_text) else rng_text[b:e]   # the shard stops before its final newline
                devt = gu.to_device(piece)
                out = gu.scan(plan, devt, len(piece), global_offset=(b + 15) // 16 * 16,
                              prev_byte=0x0A if b else -1, next_byte=0x0A if e < len(rng_text) else -1)
                total += gu.collect(plan, p, out)[0]
            p.struct.ac_trie = None
            want = checker().run(func, Params(pats, count=True), rng_text)
            assert total == want[0] > 3, (func, total, want[0])
        finally:
            p.struct.ac_trie = None
            L.krep_b200_plan_destroy(plan)


def test_search_shards_one_call_over_resident_shards():
    """krep_b200_search_shards: several resident shards (here on one GPU; on a multi-GPU box spread over the devices),
    one merged, globally replayed answer — literal with overlap policy and -m, pattern set with straddling plants, -c."""
    L = lib.load()
    wl, pats, longs = _bench_pattern_set_with_nested_pairs()
    n = 5 * (1 << 20) + 321
    spec = lib.make_spec(91, 92, 1 << 15, wl["needle"], wl["flags"])
    host = bytearray(lib.corpus_host(spec, 0, n))
    ndev = torch.cuda.device_count()
    for nsh in (1, 3, 8):
        S = ((n + nsh - 1) // nsh + 15) // 16 * 16
        for g in range(1, nsh):
            p = longs[g % len(longs)]
            host[g * S - 3:g * S - 3 + len(p)] = p
            host[g * S - 40:g * S - 36] = b"abab"
    host = bytes(host)
    from krep_b200.abi import Shard
    cases = [("aho_corasick", ALGO_AC, pats, dict()), ("aho_corasick", ALGO_AC, pats, dict(max_count=2)),
             ("sse42", ALGO_SSE42, [b"abab"], dict()), ("boyer_moore", ALGO_BMH, [b"abab"], dict(max_count=5)),
             ("boyer_moore", ALGO_BMH, [b"the"], dict(count=True)), ("boyer_moore", ALGO_BMH, [b"et"], dict(whole_word=True, count=True)),
             ("boyer_moore", ALGO_BMH, [wl["needle"]], dict(track_positions=False))]
    for func, algo, pp, opts in cases:
        p = Params(pp, **opts)
        if func == "aho_corasick":
            p.struct.ac_trie = 1
        plan = L.krep_b200_plan_create(p.ref(), algo)
        lib.check(L)
        try:
            halo = max(map(len, pp)) + 1
            for nsh in (1, 3, 8):
                S = ((n + nsh - 1) // nsh + 15) // 16 * 16
                shards = (Shard * nsh)()
                keep = []
                for g in range(nsh):
                    b, e = g * S, min((g + 1) * S, n)
                    avail = min(e + halo, n)
                    with torch.cuda.device(g % ndev):
                        buf = torch.empty(avail - b + 64, dtype=torch.uint8, device="cuda")
                        buf[: avail - b] = torch.frombuffer(bytearray(host[b:avail]), dtype=torch.uint8).cuda()
                        torch.cuda.synchronize()
                    keep.append(buf)
                    shards[g] = Shard(buf.data_ptr(), avail - b, 0, e - b, b, host[b - 1] if b else -1, host[avail] if avail < n else -1)
                res = L.krep_b200_match_result_init(16)
                cnt = L.krep_b200_search_shards(plan, p.ref(), shards, nsh, res)
                lib.check(L)
                r = res.contents
                got = (int(cnt), [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)])
                L.krep_b200_match_result_free(res)
                p.struct.ac_trie = None
                want = checker().run(func, Params(pp, **opts), host)
                if func == "aho_corasick":
                    p.struct.ac_trie = 1
                if opts.get("track_positions") is False:
                    want = (want[0], [])
                assert got == want, (func, opts, nsh, got[0], want[0])
        finally:
            p.struct.ac_trie = None
            L.krep_b200_plan_destroy(plan)


def _reference_on_tensor(func, params, host_tensor, n, with_result=True):
    """The compiled reference's kernel function called in-process on a torch CPU tensor (no bytes copy)."""
    chk = checker()
    chk._set_o(bool(params.only_matching))
    trie = None
    if func == "aho_corasick":
        trie = chk._acb(params.ref())
        params.struct.ac_trie = trie
    res = chk._new(16) if with_result else None
    try:
        cnt = chk.fn[func](params.ref(), C.cast(host_tensor.data_ptr(), C.c_char_p), n, res)
        pos = []
        if res:
            r = res.contents
            pos = [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)]
        return int(cnt), pos
    finally:
        if res:
            chk._free(res)
        if trie:
            chk._acf(trie)
            params.struct.ac_trie = None
        chk._set_o(False)


@pytest.mark.parametrize("name,gib,func,algo", [
    ("literal8", 4.0, "sse42", ALGO_SSE42),
    ("word16", 1.0, "sse42", ALGO_SSE42),
    ("icase4", 2.0, "boyer_moore", ALGO_BMH),
    ("multi1000", 0.25, "aho_corasick", ALGO_AC),
])
def test_bench_workloads_equal_the_reference_on_the_synthetic_corpus(name, gib, func, algo):
    """The bench workloads themselves (same corpus generator, seeds, needles, pattern set) against the compiled
    reference's single-chunk run, position for position, at sizes the reference still finishes in seconds."""
    import bench
    L = lib.load()
    wl = bench.WORKLOADS[name]
    n = int(gib * (1 << 30)) + 4096 + 5
    pats = bench.multi_patterns(wl["multi"], wl["needle"]) if wl.get("multi") else [wl["needle"]]
    spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, wl["period"], wl["needle"], wl["flags"])
    dev = gu.device_corpus(spec, 0, n)
    host = dev[:n].cpu()
    p = Params(pats if wl.get("multi") else wl["needle"], **wl["opts"])
    if func == "aho_corasick":
        p.struct.ac_trie = 1
    plan = L.krep_b200_plan_create(p.ref(), algo)
    lib.check(L)
    try:
        got = gu.collect(plan, p, gu.scan(plan, dev, n))
        p.struct.ac_trie = None
        want = _reference_on_tensor(func, Params(pats if wl.get("multi") else wl["needle"], **wl["opts"]), host, n)
        assert got == want, (name, got[0], want[0])
        assert got[0] >= n // wl["period"] // 2 - 2
    finally:
        p.struct.ac_trie = None
        L.krep_b200_plan_destroy(plan)


def test_two_scans_in_flight_and_async_export():
    """Two scans in flight on one device (krep_b200_scan_shard_begin twice, then _end in order): each has its own list and
    counter, each _end waits for its own finish kernel only; krep_b200_export_packed_async enqueues the row [count, keys]
    before the host knows the count.  Results must equal the one-at-a-time results."""
    from krep_b200.abi import DeviceResult, Shard
    L = lib.load()
    n = 48 * (1 << 20)
    spec = lib.make_spec(15, 16, 1 << 14, NEEDLE, 0)          # ~3000 occurrences
    dev = gu.device_corpus(spec, 0, n)
    p = Params(NEEDLE)
    plan = L.krep_b200_plan_create(p.ref(), ALGO_SSE42)
    lib.check(L)
    try:
        half = n // 2
        shards = [Shard(dev.data_ptr(), half + 16, 0, half, 0, -1, -1), Shard(dev.data_ptr(), n, half, n, 0, -1, -1)]
        want = [gu.collect(plan, p, gu.scan(plan, dev, s.avail_len, own_begin=s.own_begin, own_end=s.own_end)) for s in shards]
        rows = [torch.zeros(16385, dtype=torch.int64, device="cuda") for _ in range(2)]
        for rep in range(3):
            tickets = [C.c_int(-1), C.c_int(-1)]
            for i in range(2):
                assert L.krep_b200_scan_shard_begin(plan, C.byref(shards[i]), 1, None, C.byref(tickets[i])) == 0
                assert L.krep_b200_export_packed_async(tickets[i].value, rows[i].data_ptr(), 16384) == 0
            third = C.c_int(-1)
            assert L.krep_b200_scan_shard_begin(plan, C.byref(shards[0]), 1, None, C.byref(third)) != 0   # only two in flight
            assert L.krep_b200_last_error() != 0
            for i in range(2):
                out = DeviceResult()
                assert L.krep_b200_scan_shard_end(tickets[i].value, C.byref(out)) == 0
                lib.check(L)
                got = gu.collect(plan, p, out)
                assert got == want[i], (rep, i, got[0], want[i][0])
            torch.cuda.synchronize()
            for i in range(2):
                host = rows[i].cpu()
                cnt = int(host[0])
                assert cnt == want[i][0]
                assert [int(k) >> 3 for k in host[1:1 + cnt].tolist()] == [s for s, _ in want[i][1]]
    finally:
        L.krep_b200_plan_destroy(plan)
