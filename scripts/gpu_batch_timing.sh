#!/bin/bash
# small files: one call per text vs krep_b200_search_batch (2000 texts x 64 KiB); C calls only are timed
python - <<'PY'
import ctypes as C, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from krep_b200 import lib
from krep_b200.abi import Params, MatchResult
L = lib.load(); assert L.krep_b200_init(0) == 0
spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, 1 << 16, b"qzXv9Kpw", 0)
N = 2000
texts = [lib.corpus_host(spec, i * (1 << 16), 1 << 16) for i in range(N)]
p = Params(b"qzXv9Kpw")
bufs = [C.create_string_buffer(t, len(t)) for t in texts]
tarr = (C.c_char_p * N)(*[C.cast(b, C.c_char_p) for b in bufs]); larr = (C.c_size_t * N)(*[len(t) for t in texts])
counts = (C.c_uint64 * N)()
res = [L.krep_b200_match_result_init(16) for _ in range(N)]; rarr = (C.POINTER(MatchResult) * N)(*res)
entry = C.cast(L.krep_b200_simd_avx2_search, C.c_void_p)
fn = L.krep_b200_simd_avx2_search
def per_call():
    tot = 0
    for i in range(N):
        res[i].contents.count = 0
        tot += fn(p.ref(), C.cast(bufs[i], C.c_void_p), len(texts[i]), res[i])
    return tot
def batch():
    for r in res: r.contents.count = 0
    assert L.krep_b200_search_batch(entry, p.ref(), tarr, larr, N, counts, rarr) == 0
    return sum(counts)
per_call(); batch()
t0 = time.perf_counter(); a = per_call(); t1 = time.perf_counter(); b = batch(); t2 = time.perf_counter()
assert a == b, (a, b)
n = sum(map(len, texts))
print(f"{N} x 64 KiB ({n/1e6:.0f} MB), {a} matches: one call per text {1e3*(t1-t0):.1f} ms ({n/(t1-t0)/1e9:.2f} GB/s), one batch {1e3*(t2-t1):.1f} ms ({n/(t2-t1)/1e9:.2f} GB/s)")
PY
