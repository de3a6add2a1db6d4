"""CPU: the C-ABI library loads, exports every symbol include/krep_b200.h declares, the restated
structs have krep.h's layout, and the product fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from krep_b200 import lib
from krep_b200.abi import MatchPosition, MatchResult, Params, SearchParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "krep_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(krep_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    L = lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/krep_b200.h but not exported"


def test_struct_layout_matches_krep_h():
    # krep.h:49-60, 65-94 on LP64
    assert C.sizeof(MatchPosition) == 16
    assert C.sizeof(MatchResult) == 24
    assert SearchParams.pattern_len.offset == 8
    assert SearchParams.patterns.offset == 16
    assert SearchParams.num_patterns.offset == 32
    assert SearchParams.case_sensitive.offset == 40
    assert SearchParams.whole_word.offset == 45
    assert SearchParams.compiled_regex.offset == 48
    assert SearchParams.ac_trie.offset == 56
    assert SearchParams.max_count.offset == 64
    assert C.sizeof(SearchParams) == 72


def test_struct_layout_matches_compiled_reference():
    """Same params object drives the compiled reference and gives the expected answer -> layouts agree."""
    import oracle_util as ou
    ref = ou.reference()
    if ref is None:
        pytest.skip("compiled reference not available")
    p = Params(b"cat", whole_word=True)
    assert ref.run("boyer_moore", p, b"cat scatter catalog cat catapult cat") == \
        (3, [(0, 3), (20, 23), (33, 36)])


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        lib.search("boyer_moore", Params(b"aba"), b"abababa")
    L = lib.load()
    assert L.krep_b200_last_error() != 0


def test_dispatch_mirrors_select_search_algorithm():
    # krep.c:1771-1870 for the AVX2 build
    L = lib.load()

    def pick(pat, **kw):
        p = Params(pat, **kw)
        return L.krep_b200_get_algorithm_name(L.krep_b200_select_search_algorithm(p.ref())).decode()

    assert pick(b"a") == "memchr"
    assert pick(b"ab") == "AVX2"
    assert pick(b"ab", case_sensitive=False) == "memchr-short"
    assert pick(b"abcdefgh") == "AVX2"
    assert pick(b"abcd", case_sensitive=False) == "AVX2"      # falls back to BMH inside (krep.c:4883)
    assert pick(b"x" * 33) == "Boyer-Moore-Horspool"
    assert pick([b"a", b"b"]) == "Aho-Corasick"
    L.krep_b200_set_force_no_simd(True)
    try:
        assert pick(b"abcdefgh") == "Boyer-Moore-Horspool"
        assert pick(b"aaaa") == "Knuth-Morris-Pratt"          # repetitive and < 8 (krep.c:1862)
        assert pick(b"ab") == "memchr-short"
    finally:
        L.krep_b200_set_force_no_simd(False)
    L.krep_b200_set_algo_override(b"kmp")
    try:
        assert pick(b"abcdefgh") == "Knuth-Morris-Pratt"
    finally:
        L.krep_b200_set_algo_override(None)
