#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small invocations of every kernel family.
O=gpurun_out; mkdir -p $O
cat > /tmp/san_case.py <<'PY'
import random, sys
sys.path.insert(0, "tests")
import oracle_util as ou
from krep_b200 import lib
from krep_b200.abi import Params
rng = random.Random(5)
words = [b"needle", b"the", b"quick", b"ab", b"abab", b"NEEDLE", b"x", b"haystack", b"aaa", b"needle_7", b"fox_1"]
t = bytearray()
while len(t) < 300_000:
    t += rng.choice(words) + rng.choice([b" ", b"\n", b"", b"_", b", "])
text = bytes(t[:300_000])
chk = ou.reference() or ou.port()
cases = [("sse42", [b"needle"], {}), ("boyer_moore", [b"NeEdLe"], dict(case_sensitive=False)), ("boyer_moore", [b"ab"], dict(whole_word=True)),
         ("memchr", [b"x"], {}), ("memchr_short", [b"ab"], dict(case_sensitive=False)), ("kmp", [b"abab"], {}),
         ("avx2", [b"needle the quick ab"], {}), ("boyer_moore", [b"the"], dict(count=True)),              # fused -c, window filter
         ("boyer_moore", [b"needle_7"], dict(count=True, whole_word=True)), ("memchr", [b"x"], dict(count=True)),  # fused -c, aligned filter / 1 byte
         ("sse42", [b"ab"], dict(only_matching=True)),                                                       # masked window kernel: warp-cooperative emission
         ("aho_corasick", [b"needle", b"haystack", b"quick the", b"fox_1 "], {}),                    # shortest 6: tri
         ("aho_corasick", [b"needle_7", b"haystack", b"quick the"], dict(case_sensitive=False)),      # shortest 8: quad, fold
         ("aho_corasick", [b"ab", b"needle", b"x"], {}), ("aho_corasick", [b"abab", b"quick"], {}),   # short patterns: k_ac_scan s=1, s=2
         ]
for func, pats, opts in cases:
    got = lib.search(func, Params(pats, **opts), text)
    want = chk.run(func, Params(pats, **opts), text)
    assert got == want, (func, pats, opts, got[0], want[0])
    print("ok", func, pats[0], got[0], flush=True)
PY
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/san_case.py > $O/sanitize_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" $O/sanitize_$tool.log | tail -20
done
