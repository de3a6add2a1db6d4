"""-m gpu: fused -c (csrc/scan_count.cu) — the scan counts matching lines itself; only a count leaves the GPU.
Must equal the reference's -c result (count_lines_mode, krep.c:1331-1351 and the other literal kernels' branches)
for every eligible kernel, across staging chunks, ranges and resident shards that cut lines."""
import ctypes as C
import random

import pytest
import torch

import gpu_util as gu
import oracle_util as ou
from krep_b200 import lib
from krep_b200.abi import ALGO_BMH, ALGO_KMP, ALGO_MEMCHR, ALGO_MEMCHR_SHORT, ALGO_SSE42, Params, Shard, SIZE_MAX

pytestmark = pytest.mark.gpu


def checker():
    return ou.reference() or ou.port()


class LineCount(C.Structure):
    _fields_ = [("lines", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


def _text(rng, n, nl_rate, words):
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words)
        r = rng.random()
        out += b"\n" if r < nl_rate else rng.choice([b" ", b"", b"_", b", ", b"."])
    return bytes(out[:n])


WORDS = [b"the", b"quick", b"Brown", b"fox_1", b"needle", b"NEEDLE", b"ab", b"abab", b"aaa", b"x", b"then", b"other",
         b"needleneedle0016", b"haystack in a long line with the needle"]


@pytest.mark.parametrize("func,pat", [
    ("boyer_moore", b"the"), ("boyer_moore", b"needle"), ("boyer_moore", b"aaa"), ("boyer_moore", b"needleneedle0016"),
    ("boyer_moore", b"haystack in a long line with the needle"), ("sse42", b"the"), ("sse42", b"abab"), ("sse42", b"fox_1"),
    ("sse42", b"needle"), ("sse42", b"quick Br"), ("kmp", b"abab"), ("kmp", b"needle"), ("memchr", b"x"),
    ("memchr_short", b"ab"), ("memchr_short", b"the"), ("avx2", b"th"), ("avx2", b"needle_"),
])
@pytest.mark.parametrize("opts", [dict(), dict(case_sensitive=False), dict(whole_word=True), dict(max_count=7),
                                  dict(whole_word=True, case_sensitive=False, max_count=2)])
def test_count_lines_host_path(func, pat, opts, monkeypatch):
    """All eligible kernels x options, texts with many / few / no newlines, cut into small staging chunks and ranges so
    that lines straddle every kind of edge."""
    rng = random.Random(len(pat) * 131 + len(func))
    chk = checker()
    monkeypatch.setenv("KREP_B200_STAGE_MB", "1")
    monkeypatch.setenv("KREP_B200_CHUNK_MB", "1")
    monkeypatch.setenv("KREP_B200_COUNT_PART_KB", "8")       # many partitions per chunk
    for n, nl_rate, ranges in [(300_000, 0.2, "1"), (3 * (1 << 20) + 777, 0.02, "3"), (2 * (1 << 20) + 5, 0.0005, "2"),
                               (1_500_000, 0.0, "1"), (40, 0.3, "1"), (5000, 0.5, "1")]:
        monkeypatch.setenv("KREP_B200_RANGES", ranges)
        text = _text(rng, n, nl_rate, WORDS)
        got = lib.search(func, Params(pat, count=True, **opts), text)
        want = chk.run(func, Params(pat, count=True, **opts), text)
        assert got == want, (func, pat, opts, n, nl_rate, got[0], want[0])
        # the list path (k_lit_* + host replay) is the same function of the text
        monkeypatch.setenv("KREP_B200_NO_FUSED_COUNT", "1")
        assert lib.search(func, Params(pat, count=True, **opts), text) == want
        monkeypatch.delenv("KREP_B200_NO_FUSED_COUNT")


def test_count_lines_edges_and_dense_hits(monkeypatch):
    """Hand-made texts: a hit on every line, several hits per line, hits at the very start / end, lines longer than a
    partition, no trailing newline, newline-only text, and every buffer length around the vector / tail boundary."""
    chk = checker()
    monkeypatch.setenv("KREP_B200_COUNT_PART_KB", "2")
    cases = [b"the\n" * 5000, b"the the the\n" * 3000, b"the" * 20000, b"\n" * 5000, b"x" * 70000 + b"the" + b"y" * 70000,
             (b"a" * 5000 + b"the" + b"b" * 5000 + b"\n") * 20, b"the", b"\nthe", b"the\n", b"xthe\nthe", b""]
    for t in cases:
        for pat, func in ((b"the", "boyer_moore"), (b"the", "sse42"), (b"t", "memchr"), (b"th", "memchr_short")):
            for opts in (dict(), dict(whole_word=True)):
                got = lib.search(func, Params(pat, count=True, **opts), t)
                want = chk.run(func, Params(pat, count=True, **opts), t)
                assert got == want, (t[:30], len(t), pat, func, opts, got[0], want[0])
    base = (b"needle7 and more text\nsecond line\n" * 8)
    for n in range(len(base) - 70, len(base)):
        for pat in (b"needle7", b"line", b"text\nsec"[:4]):
            t = base[:n]
            got = lib.search("boyer_moore", Params(pat, count=True), t)
            want = chk.run("boyer_moore", Params(pat, count=True), t)
            assert got == want, (n, pat, got[0], want[0])


@pytest.mark.parametrize("pat,algo,func,opts", [
    (b"the", ALGO_BMH, "boyer_moore", dict()),
    (b"needle", ALGO_SSE42, "sse42", dict(whole_word=True)),
    (b"NEEDLEneedle0016", ALGO_BMH, "boyer_moore", dict(case_sensitive=False)),
    (b"x", ALGO_MEMCHR, "memchr", dict()),
])
def test_count_lines_resident_shards_combine(pat, algo, func, opts):
    """krep_b200_count_lines_shard on 1 / 2 / 5 / 8 shards that cut lines anywhere + krep_b200_combine_line_counts."""
    L = lib.load()
    L.krep_b200_count_lines_shard.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Shard), C.c_void_p, C.POINTER(LineCount)]
    L.krep_b200_count_lines_shard.restype = C.c_int
    L.krep_b200_combine_line_counts.argtypes = [C.POINTER(LineCount), C.c_size_t, C.c_size_t]
    L.krep_b200_combine_line_counts.restype = C.c_uint64
    rng = random.Random(len(pat))
    for n, nl_rate in [(2_000_000, 0.05), (700_000, 0.001), (300_000, 0.0)]:
        text = _text(rng, n, nl_rate, WORDS)
        want = checker().run(func, Params(pat, count=True, **opts), text)[0]
        p = Params(pat, count=True, **opts)
        plan = L.krep_b200_plan_create(p.ref(), algo)
        lib.check(L)
        try:
            whole = gu.to_device(text)
            for nsh in (1, 2, 5, 8):
                S = ((n + nsh - 1) // nsh + 15) // 16 * 16
                recs = (LineCount * nsh)()
                keep = []
                for g in range(nsh):
                    b, e = g * S, min((g + 1) * S, n)
                    if g % 2 == 0:      # the shard as a view into the whole text ...
                        sh = Shard(whole.data_ptr(), min(e + len(pat) + 1, n), b, e, 0, -1, -1)
                    else:               # ... or in a buffer of its own, with halo and context bytes
                        avail = min(e + len(pat) + 1, n)
                        buf = gu.to_device(text[b:avail])
                        keep.append(buf)
                        sh = Shard(buf.data_ptr(), avail - b, 0, e - b, b, text[b - 1] if b else -1, text[avail] if avail < n else -1)
                    rc = L.krep_b200_count_lines_shard(plan, p.ref(), C.byref(sh), None, C.byref(recs[g]))
                    lib.check(L)
                    assert rc == 0
                got = L.krep_b200_combine_line_counts(recs, nsh, SIZE_MAX)
                assert got == want, (pat, opts, n, nl_rate, nsh, got, want, [(r.lines, r.flags) for r in recs])
                assert L.krep_b200_combine_line_counts(recs, nsh, 3) == min(want, 3)
        finally:
            L.krep_b200_plan_destroy(plan)


def test_count_lines_on_the_bench_corpus_matches_reference():
    """`-c the` (config 1's command) on 64 MiB of the synthetic corpus with `the` planted every 1 KiB and every 64 B."""
    import bench
    for period in (1 << 10, 1 << 6):
        spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, period, b"the", 0)
        n = 64 * (1 << 20) + 321
        dev = gu.device_corpus(spec, 0, n)
        host = bytes(dev[:n].cpu().numpy())
        want = checker().run("boyer_moore", Params(b"the", count=True), host)
        got = lib.search("boyer_moore", Params(b"the", count=True), host)
        assert got == want and got[0] > n // period // 2, (period, got[0], want[0])
