"""ctypes binding of libkrep_b200.so — plumbing for tests and bench.py, not a second implementation.

The library is loaded from krep_b200/libkrep_b200.so (built in-tree by krep_b200/build.py).  Loading
never falls back to anything else: if the .so is missing this raises, and if no sm_100 device is
usable every search call reports an error through krep_b200_last_error().
"""
import ctypes as C
import os

from .abi import (CorpusSpec, DeviceResult, MatchResult, Params, SearchParams, Shard, SEARCH_FUNC, SIZE_MAX)  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KREP_B200_LIB") or os.path.join(_HERE, "libkrep_b200.so")  # override: kernel-variant builds

SEARCH_ENTRIES = {
    "boyer_moore": "krep_b200_boyer_moore_search",
    "kmp": "krep_b200_kmp_search",
    "memchr": "krep_b200_memchr_search",
    "memchr_short": "krep_b200_memchr_short_search",
    "sse42": "krep_b200_simd_sse42_search",
    "avx2": "krep_b200_simd_avx2_search",
    "avx512": "krep_b200_simd_avx512_search",
    "aho_corasick": "krep_b200_aho_corasick_search",
    "neon": "krep_b200_neon_search",
}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    sig = [C.POINTER(SearchParams), C.c_void_p, C.c_size_t, C.POINTER(MatchResult)]
    for name in SEARCH_ENTRIES.values():
        f = getattr(L, name)
        f.argtypes = sig
        f.restype = C.c_uint64
    L.krep_b200_init.argtypes = [C.c_int]
    L.krep_b200_init.restype = C.c_int
    L.krep_b200_last_error.restype = C.c_int
    L.krep_b200_last_error_string.restype = C.c_char_p
    L.krep_b200_version.restype = C.c_char_p
    L.krep_b200_set_only_matching.argtypes = [C.c_bool]
    L.krep_b200_get_only_matching.restype = C.c_bool
    L.krep_b200_set_force_no_simd.argtypes = [C.c_bool]
    L.krep_b200_set_algo_override.argtypes = [C.c_char_p]
    L.krep_b200_select_search_algorithm.argtypes = [C.POINTER(SearchParams)]
    L.krep_b200_select_search_algorithm.restype = C.c_void_p
    L.krep_b200_get_algorithm_name.argtypes = [C.c_void_p]
    L.krep_b200_get_algorithm_name.restype = C.c_char_p
    L.krep_b200_ac_trie_build.argtypes = [C.POINTER(SearchParams)]
    L.krep_b200_ac_trie_build.restype = C.c_void_p
    L.krep_b200_ac_trie_free.argtypes = [C.c_void_p]
    L.krep_b200_ac_trie_root_has_outputs.argtypes = [C.c_void_p]
    L.krep_b200_ac_trie_root_has_outputs.restype = C.c_bool
    L.krep_b200_match_result_init.argtypes = [C.c_uint64]
    L.krep_b200_match_result_init.restype = C.POINTER(MatchResult)
    L.krep_b200_match_result_add.argtypes = [C.POINTER(MatchResult), C.c_size_t, C.c_size_t]
    L.krep_b200_match_result_add.restype = C.c_bool
    L.krep_b200_match_result_free.argtypes = [C.POINTER(MatchResult)]
    L.krep_b200_match_result_merge.argtypes = [C.POINTER(MatchResult), C.POINTER(MatchResult), C.c_size_t]
    L.krep_b200_match_result_merge.restype = C.c_bool
    L.krep_b200_plan_create.argtypes = [C.POINTER(SearchParams), C.c_int]
    L.krep_b200_plan_create.restype = C.c_void_p
    L.krep_b200_plan_destroy.argtypes = [C.c_void_p]
    L.krep_b200_plan_filter_name.argtypes = [C.c_void_p]
    L.krep_b200_plan_filter_name.restype = C.c_char_p
    L.krep_b200_scan_shard.argtypes = [C.c_void_p, C.POINTER(Shard), C.c_int, C.c_void_p, C.POINTER(DeviceResult)]
    L.krep_b200_scan_shard.restype = C.c_int
    L.krep_b200_collect.argtypes = [C.c_void_p, C.POINTER(SearchParams), C.POINTER(DeviceResult), C.POINTER(MatchResult)]
    L.krep_b200_collect.restype = C.c_uint64
    L.krep_b200_replay.argtypes = [C.c_int, C.POINTER(SearchParams), C.c_bool, C.POINTER(C.c_uint64), C.c_uint64,
                                   C.c_void_p, C.c_size_t, C.POINTER(MatchResult)]
    L.krep_b200_replay.restype = C.c_uint64
    L.krep_b200_replay_lines.argtypes = [C.c_int, C.POINTER(SearchParams), C.c_bool, C.POINTER(C.c_uint64), C.c_uint64,
                                         C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(MatchResult)]
    L.krep_b200_replay_lines.restype = C.c_uint64
    L.krep_b200_search_batch.argtypes = [C.c_void_p, C.POINTER(SearchParams), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t,
                                         C.POINTER(C.c_uint64), C.POINTER(C.POINTER(MatchResult))]
    L.krep_b200_search_batch.restype = C.c_int
    L.krep_b200_scan_shard_begin.argtypes = [C.c_void_p, C.POINTER(Shard), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.krep_b200_scan_shard_begin.restype = C.c_int
    L.krep_b200_scan_shard_end.argtypes = [C.c_int, C.POINTER(DeviceResult)]
    L.krep_b200_scan_shard_end.restype = C.c_int
    L.krep_b200_export_packed.argtypes = [C.POINTER(DeviceResult), C.c_void_p, C.c_uint64, C.c_void_p]
    L.krep_b200_export_packed.restype = C.c_int
    L.krep_b200_export_packed_async.argtypes = [C.c_int, C.c_void_p, C.c_uint64]
    L.krep_b200_export_packed_async.restype = C.c_int
    L.krep_b200_merge_keys.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p]
    L.krep_b200_merge_keys.restype = C.c_uint64
    L.krep_b200_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.krep_b200_device_count.restype = C.c_int
    L.krep_b200_search_shards.argtypes = [C.c_void_p, C.POINTER(SearchParams), C.POINTER(Shard), C.c_uint32, C.POINTER(MatchResult)]
    L.krep_b200_search_shards.restype = C.c_uint64
    L.krep_b200_export_keys.argtypes = [C.POINTER(DeviceResult), C.c_void_p, C.c_uint64, C.c_void_p]
    L.krep_b200_export_keys.restype = C.c_int
    L.krep_b200_last_kernel_ms.restype = C.c_float
    L.krep_b200_launch_count.restype = C.c_uint64
    for n in ("krep_b200_ac_key_end", "krep_b200_ac_key_start"):
        getattr(L, n).argtypes = [C.c_uint64]
        getattr(L, n).restype = C.c_uint64
    L.krep_b200_ac_key_pattern.argtypes = [C.c_uint64]
    L.krep_b200_ac_key_pattern.restype = C.c_uint32
    L.krep_b200_corpus_generate.argtypes = [C.POINTER(CorpusSpec), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    L.krep_b200_corpus_generate.restype = C.c_int
    L.krep_b200_corpus_generate_host.argtypes = [C.POINTER(CorpusSpec), C.c_void_p, C.c_uint64, C.c_uint64]
    L.krep_b200_corpus_generate_host.restype = C.c_int
    _lib = L
    return L


def check(L=None):
    L = L or load()
    if L.krep_b200_last_error() != 0:
        raise RuntimeError("krep_b200: " + L.krep_b200_last_error_string().decode())


def search(func, params, text, with_result=True, text_ptr=None, text_len=None):
    """Calls one search_func_t entry point on host text. -> (count, [(start, end), ...]).

    `text` is bytes (or pass text_ptr/text_len for a raw host buffer, e.g. pinned memory).
    Raises if the library reported an error (missing GPU, CUDA failure): there is no fallback.
    """
    L = load()
    L.krep_b200_set_only_matching(bool(params.only_matching))
    own_trie = False
    if func == "aho_corasick" and not params.struct.ac_trie:
        params.struct.ac_trie = L.krep_b200_ac_trie_build(params.ref())
        own_trie = True
    res = L.krep_b200_match_result_init(16) if with_result else None
    try:
        if text_ptr is None:
            buf = C.cast(C.c_char_p(text), C.c_void_p)
            n = len(text)
        else:
            buf, n = C.c_void_p(text_ptr), text_len
        cnt = getattr(L, SEARCH_ENTRIES[func])(params.ref(), buf, n, res)
        check(L)
        pos = []
        if res:
            r = res.contents
            pos = [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)]
        return int(cnt), pos
    finally:
        if res:
            L.krep_b200_match_result_free(res)
        if own_trie:
            L.krep_b200_ac_trie_free(params.struct.ac_trie)
            params.struct.ac_trie = None
        L.krep_b200_set_only_matching(False)


def make_spec(seed, plant_seed=0, plant_period=0, needle=b"", flags=0):
    s = CorpusSpec()
    s.seed = seed
    s.plant_seed = plant_seed
    s.plant_period = plant_period
    s._needle_keepalive = C.create_string_buffer(needle, max(len(needle), 1))
    s.needle = C.cast(s._needle_keepalive, C.c_char_p)
    s.needle_len = len(needle)
    s.flags = flags
    return s


def corpus_host(spec, offset, length):
    L = load()
    buf = C.create_string_buffer(length)
    rc = L.krep_b200_corpus_generate_host(C.byref(spec), buf, offset, length)
    if rc != 0:
        raise RuntimeError("corpus_generate_host failed")
    return buf.raw


def search_batch(func, params, texts, with_result=True):
    """krep_b200_search_batch on a list of bytes objects. -> [(count, [(start, end), ...]), ...]"""
    L = load()
    L.krep_b200_set_only_matching(bool(params.only_matching))
    own_trie = False
    if func == "aho_corasick" and not params.struct.ac_trie:
        params.struct.ac_trie = L.krep_b200_ac_trie_build(params.ref())
        own_trie = True
    n = len(texts)
    bufs = [C.create_string_buffer(t, max(len(t), 1)) for t in texts]
    tarr = (C.c_char_p * max(n, 1))(*[C.cast(b, C.c_char_p) for b in bufs])
    larr = (C.c_size_t * max(n, 1))(*[len(t) for t in texts])
    counts = (C.c_uint64 * max(n, 1))()
    res = [L.krep_b200_match_result_init(16) for _ in range(n)] if with_result else []
    rarr = (C.POINTER(MatchResult) * max(n, 1))(*res) if with_result else None
    try:
        entry = C.cast(getattr(L, SEARCH_ENTRIES[func]), C.c_void_p)
        rc = L.krep_b200_search_batch(entry, params.ref(), tarr, larr, n, counts, rarr)
        check(L)
        assert rc == 0, rc
        out = []
        for i in range(n):
            pos = []
            if with_result:
                r = res[i].contents
                pos = [(r.positions[k].start_offset, r.positions[k].end_offset) for k in range(r.count)]
            out.append((int(counts[i]), pos))
        return out
    finally:
        for r in res:
            L.krep_b200_match_result_free(r)
        if own_trie:
            L.krep_b200_ac_trie_free(params.struct.ac_trie)
            params.struct.ac_trie = None
        L.krep_b200_set_only_matching(False)
