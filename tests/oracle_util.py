"""Loads the parity checkers (oracle port + compiled reference) for tests. Test infrastructure only."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from krep_b200.abi import MatchResult, SearchParams, Params, SIZE_MAX  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_oracle  # noqa: E402

_SIG = [C.POINTER(SearchParams), C.c_char_p, C.c_size_t, C.POINTER(MatchResult)]

# oracle function name -> reference symbol
FUNCS = {
    "boyer_moore": ("oracle_boyer_moore_search", "boyer_moore_search"),
    "kmp": ("oracle_kmp_search", "kmp_search"),
    "memchr": ("oracle_memchr_search", "memchr_search"),
    "memchr_short": ("oracle_memchr_short_search", "memchr_short_search"),
    "sse42": ("oracle_sse42_search", "simd_sse42_search"),
    "aho_corasick": ("oracle_aho_corasick_search", "aho_corasick_search"),
    "avx2": ("oracle_avx2_search", "simd_avx2_search"),
}
# only present in an AVX-512 build of the reference (oracle/_ref/libkrep_ref512.so)
FUNCS_512 = {"avx512": ("oracle_avx512_search", "simd_avx512_search")}
# only present in the reference's ARM build (oracle/_ref/libkrep_refneon.so: compiled here against a scalar arm_neon.h)
FUNCS_NEON = {"neon": ("oracle_neon_search", "neon_search")}


class _Checker:
    def __init__(self, lib, kind, funcs=None):
        self.lib = lib
        self.kind = kind  # "port" | "reference"
        idx = 0 if kind == "port" else 1
        self.fn = {}
        for k, names in (funcs or FUNCS).items():
            f = getattr(lib, names[idx])
            f.argtypes = _SIG
            f.restype = C.c_uint64
            self.fn[k] = f
        if kind == "port":
            self._set_o = lib.oracle_set_only_matching
            self._new = lib.oracle_result_new
            self._free = lib.oracle_result_free
            self._acb = lib.oracle_ac_build
            self._acf = lib.oracle_ac_free
        else:
            self._set_o = lib.krep_ref_set_only_matching
            self._new = lib.match_result_init
            self._free = lib.match_result_free
            self._acb = lib.ac_trie_build
            self._acf = lib.ac_trie_free
        self._set_o.argtypes = [C.c_bool]
        self._new.argtypes = [C.c_uint64]
        self._new.restype = C.POINTER(MatchResult)
        self._free.argtypes = [C.POINTER(MatchResult)]
        self._acb.argtypes = [C.POINTER(SearchParams)]
        self._acb.restype = C.c_void_p
        self._acf.argtypes = [C.c_void_p]

    def run(self, func, params, text, with_result=True):
        """-> (count, [(start, end), ...])"""
        self._set_o(bool(params.only_matching))
        trie = None
        if func == "aho_corasick":
            trie = self._acb(params.ref())
            params.struct.ac_trie = trie
        res = self._new(16) if with_result else None
        try:
            cnt = self.fn[func](params.ref(), text, len(text), res)
            pos = []
            if res:
                r = res.contents
                pos = [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)]
            return int(cnt), pos
        finally:
            if res:
                self._free(res)
            if trie:
                self._acf(trie)
                params.struct.ac_trie = None
            self._set_o(False)


_cache = {}


def port():
    if "port" not in _cache:
        _cache["port"] = _Checker(C.CDLL(build_oracle.build_port()), "port", {**FUNCS, **FUNCS_512, **FUNCS_NEON})
    return _cache["port"]


def reference():
    """The compiled unmodified reference, or None when neither sources nor a prebuilt .so exist."""
    if "ref" not in _cache:
        lib, _ = build_oracle.build_ref()
        _cache["ref"] = _Checker(C.CDLL(lib), "reference") if lib else None
    return _cache["ref"]


def reference512():
    """The reference compiled as its AVX-512 build (simd_avx512_search exists only there), or None when it is not
    available or this CPU cannot execute AVX-512BW."""
    if "ref512" not in _cache:
        lib = build_oracle.build_ref512()
        ok = False
        if lib:
            try:
                with open("/proc/cpuinfo") as f:
                    ok = "avx512bw" in f.read()
            except OSError:
                ok = False
        _cache["ref512"] = _Checker(C.CDLL(lib), "reference", {**FUNCS, **FUNCS_512}) if ok else None
    return _cache["ref512"]


def reference_neon():
    """The reference's NEON kernel compiled on x86 (see build_oracle.build_refneon), or None."""
    if "refneon" not in _cache:
        lib = build_oracle.build_refneon()
        base = {k: v for k, v in FUNCS.items() if k not in ("sse42", "avx2")}  # no x86 SIMD kernels in that build
        _cache["refneon"] = _Checker(C.CDLL(lib), "reference", {**base, **FUNCS_NEON}) if lib else None
    return _cache["refneon"]


def ref_cli():
    return build_oracle.build_ref()[1]
