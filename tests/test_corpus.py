"""CPU: the synthetic corpus is a pure function of (spec, position)."""
from krep_b200 import lib
from krep_b200.abi import CORPUS_EMBED_HALF, CORPUS_RANDOM_CASE


def test_position_addressable():
    s = lib.make_spec(0x5EED0001, 0x5EED0002, 4096, b"qzXv9Kpw")
    whole = lib.corpus_host(s, 0, 40000)
    for off, ln in [(0, 1), (16, 100), (37, 4099), (4090, 30), (12345, 20000), (39999, 1)]:
        assert lib.corpus_host(s, off, ln) == whole[off:off + ln]


def test_planted_needles_and_lines():
    s = lib.make_spec(1, 2, 4096, b"qzXv9Kpw")
    t = lib.corpus_host(s, 0, 1 << 18)
    assert t.count(b"qzXv9Kpw") >= (1 << 18) // 4096 - 2   # plants may overwrite each other at block edges
    lines = t.split(b"\n")
    # one '\n' per 96-byte segment, except where a planted needle overwrote it
    assert max(map(len, lines)) <= 3 * 96 and (1 << 18) // 96 - 64 <= len(lines) <= (1 << 18) // 96 + 1
    assert set(t) <= set(b"abcdefghijklmnopqrstuvwxyz ETAOIN0123,\n" + b"qzXv9Kpw")


def test_flags():
    s = lib.make_spec(3, 4, 8192, b"QzXv", flags=CORPUS_RANDOM_CASE)
    t = lib.corpus_host(s, 0, 1 << 18)
    assert t.lower().count(b"qzxv") >= 30 and t.count(b"QzXv") < t.lower().count(b"qzxv")
    s2 = lib.make_spec(3, 4, 8192, b"needleneedle0016", flags=CORPUS_EMBED_HALF)
    t2 = lib.corpus_host(s2, 0, 1 << 18)
    assert t2.count(b"xneedleneedle0016x") >= 10 and t2.count(b" needleneedle0016 ") >= 10
