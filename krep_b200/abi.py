"""ctypes mirror of include/krep_b200.h (the types restated from krep.h:49-101).

Plumbing only: tests and bench.py use these to call the C ABI of libkrep_b200.so,
the oracle port and the compiled reference with the very same structs.
"""
import ctypes as C

SIZE_MAX = (1 << 64) - 1


class MatchPosition(C.Structure):  # krep.h:49-53
    _fields_ = [("start_offset", C.c_size_t), ("end_offset", C.c_size_t)]


class MatchResult(C.Structure):  # krep.h:55-60
    _fields_ = [("positions", C.POINTER(MatchPosition)), ("count", C.c_uint64), ("capacity", C.c_uint64)]


class SearchParams(C.Structure):  # krep.h:65-94
    _fields_ = [
        ("pattern", C.c_char_p),
        ("pattern_len", C.c_size_t),
        ("patterns", C.POINTER(C.c_char_p)),
        ("pattern_lens", C.POINTER(C.c_size_t)),
        ("num_patterns", C.c_size_t),
        ("case_sensitive", C.c_bool),
        ("use_regex", C.c_bool),
        ("count_lines_mode", C.c_bool),
        ("count_matches_mode", C.c_bool),
        ("track_positions", C.c_bool),
        ("whole_word", C.c_bool),
        ("compiled_regex", C.c_void_p),
        ("ac_trie", C.c_void_p),
        ("max_count", C.c_size_t),
    ]


SEARCH_FUNC = C.CFUNCTYPE(C.c_uint64, C.POINTER(SearchParams), C.c_void_p, C.c_size_t, C.POINTER(MatchResult))


class Shard(C.Structure):  # krep_b200_shard_t
    _fields_ = [
        ("d_text", C.c_void_p),
        ("avail_len", C.c_uint64),
        ("own_begin", C.c_uint64),
        ("own_end", C.c_uint64),
        ("global_offset", C.c_uint64),
        ("prev_byte", C.c_int32),
        ("next_byte", C.c_int32),
    ]


class DeviceResult(C.Structure):  # krep_b200_device_result_t
    _fields_ = [
        ("count", C.c_uint64),
        ("stored", C.c_uint64),
        ("d_keys", C.c_void_p),
        ("overflow", C.c_int),
        ("text_len", C.c_uint64),
        ("d_line_bounds", C.c_void_p),
        ("device", C.c_int32),
        ("slot", C.c_int32),
        ("serial", C.c_uint64),
    ]


class CorpusSpec(C.Structure):  # krep_b200_corpus_spec_t
    _fields_ = [
        ("seed", C.c_uint64),
        ("plant_seed", C.c_uint64),
        ("plant_period", C.c_uint64),
        ("needle", C.c_char_p),
        ("needle_len", C.c_uint32),
        ("flags", C.c_uint32),
    ]


CORPUS_RANDOM_CASE = 1
CORPUS_EMBED_HALF = 2

ALGO_BMH, ALGO_KMP, ALGO_MEMCHR, ALGO_MEMCHR_SHORT, ALGO_SSE42, ALGO_AVX2, ALGO_AVX512, ALGO_AC, ALGO_NEON = range(9)


class Params:
    """Owns the Python-side buffers behind one search_params_t.

    Mirrors the reference tests' create_literal_params (test/test_krep.c:208-249):
    count_lines_mode = -c && !-o, count_matches_mode = -c && -o,
    track_positions = !(-c && !-o)   (krep.c:3811-3814).
    """

    def __init__(self, patterns, case_sensitive=True, count=False, only_matching=False,
                 whole_word=False, max_count=SIZE_MAX, track_positions=None):
        if isinstance(patterns, (bytes, bytearray)):
            patterns = [bytes(patterns)]
        self.patterns = [bytes(p) for p in patterns]
        n = len(self.patterns)
        # c_char_p would stop at NUL bytes when read back, but the struct only stores pointers.
        self._bufs = [C.create_string_buffer(p, len(p) + 16) for p in self.patterns]  # +16: krep.c:4725 over-read
        self._arr = (C.c_char_p * max(n, 1))(*[C.cast(b, C.c_char_p) for b in self._bufs])
        self._lens = (C.c_size_t * max(n, 1))(*[len(p) for p in self.patterns])
        s = SearchParams()
        s.patterns = C.cast(self._arr, C.POINTER(C.c_char_p))
        s.pattern_lens = C.cast(self._lens, C.POINTER(C.c_size_t))
        s.num_patterns = n
        if n:
            s.pattern = C.cast(self._bufs[0], C.c_char_p)
            s.pattern_len = len(self.patterns[0])
        s.case_sensitive = case_sensitive
        s.use_regex = False
        s.count_lines_mode = bool(count and not only_matching)
        s.count_matches_mode = bool(count and only_matching)
        s.track_positions = (not (count and not only_matching)) if track_positions is None else track_positions
        s.whole_word = whole_word
        s.compiled_regex = None
        s.ac_trie = None
        s.max_count = max_count
        self.only_matching = only_matching
        self.struct = s

    def ref(self):
        return C.byref(self.struct)
