/* krep_oracle.c — CPU restatement of krep's byte-scanning kernels.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it;
 * the product library (krep_b200/csrc) never links or calls anything here.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   (1) the known-answer vectors of the reference's own tests
 *       (test/test_krep.c, test/test_multiple_patterns.c — see tests/golden/),
 *   (2) oracle/_ref/libkrep_ref.so — the unmodified reference sources compiled
 *       in place by oracle/build_oracle.py — on seeded random inputs, and
 *   (3) fixtures generated from (2) and committed under tests/golden/.
 *
 * Each function is a plain, scalar restatement (no SIMD, no tables beyond what
 * the algorithm needs) of the reference function cited above it, written so
 * that the *observable* behaviour — return value, appended positions and their
 * order, reaction to -w / -c / -m / -o — is identical.  Scan order and skip
 * heuristics are not reproduced where they cannot be observed.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <stdio.h>
#include "../include/krep_b200.h"

/* ---- state mirrored from krep.c's file-static globals (krep.c:117) ------- */
static bool g_only_matching = false;
void oracle_set_only_matching(bool on) { g_only_matching = on; }
bool oracle_get_only_matching(void) { return g_only_matching; }

/* ---- C-locale helpers (krep.c:125-134, krep.h:298-319) ------------------- */
static inline unsigned char lc(unsigned char c) { return (c >= 'A' && c <= 'Z') ? (unsigned char)(c + 32) : c; }
static inline bool wordc(unsigned char c)
{
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}
static inline bool whole_word(const char *t, size_t n, size_t s, size_t e)
{
    if (s > 0 && wordc((unsigned char)t[s - 1])) return false;
    if (e < n && wordc((unsigned char)t[e])) return false;
    return true;
}
/* krep.c:363-398: byte after the last '\n' in [0,pos) */
static size_t line_start(const char *t, size_t n, size_t pos)
{
    if (pos > n) pos = n;
    while (pos > 0 && t[pos - 1] != '\n') pos--;
    return pos;
}
/* krep.c:401-408: index of the first '\n' at or after pos, else n */
static size_t line_end(const char *t, size_t n, size_t pos)
{
    while (pos < n && t[pos] != '\n') pos++;
    return pos < n ? pos : n;
}

/* ---- result vector (krep.c:139-251) -------------------------------------- */
match_result_t *oracle_result_new(uint64_t cap)
{
    match_result_t *r = (match_result_t *)malloc(sizeof *r);
    if (!r) return NULL;
    if (cap == 0) cap = 16;
    r->positions = (match_position_t *)malloc(cap * sizeof(match_position_t));
    r->count = 0;
    r->capacity = cap;
    return r;
}
void oracle_result_free(match_result_t *r)
{
    if (!r) return;
    free(r->positions);
    free(r);
}
static bool push(match_result_t *r, size_t s, size_t e)
{
    if (!r) return false;
    if (r->count >= r->capacity)
    {
        uint64_t nc = r->capacity ? r->capacity * 2 : 16;
        match_position_t *np = (match_position_t *)realloc(r->positions, nc * sizeof *np);
        if (!np) return false;
        r->positions = np;
        r->capacity = nc;
    }
    r->positions[r->count].start_offset = s;
    r->positions[r->count].end_offset = e;
    r->count++;
    return true;
}

/* does the literal occur at t+i ? (memcmp / memory_equals_case_insensitive, krep.c:1198) */
static inline bool occurs(const unsigned char *t, const unsigned char *p, size_t m, bool cs)
{
    if (cs) return memcmp(t, p, m) == 0;
    for (size_t k = 0; k < m; k++)
        if (lc(t[k]) != lc(p[k])) return false;
    return true;
}
/* first occurrence at or after `from`, or SIZE_MAX */
static size_t next_occ(const unsigned char *t, size_t n, const unsigned char *p, size_t m, bool cs, size_t from)
{
    if (m == 0 || n < m) return SIZE_MAX;
    for (size_t i = from; i + m <= n; i++)
        if (occurs(t + i, p, m, cs)) return i;
    return SIZE_MAX;
}

/* ==========================================================================
 * boyer_moore_search — krep.c:1260-1385.
 * Horspool's shift never skips an occurrence, so the scan visits every
 * occurrence in ascending order; the only observable shift is the one taken
 * after a hit: bad-char shift (>=1, never past the next occurrence) by
 * default, pattern_len when only_matching && !count_lines (krep.c:1371).
 * ========================================================================== */
uint64_t oracle_boyer_moore_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return 0; /* krep.c:1266 */
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    size_t m = P->pattern_len;
    if (m == 0 || n < m) return 0; /* krep.c:1278 */
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, i = 0;
    for (;;)
    {
        size_t s = next_occ(t, n, p, m, P->case_sensitive, i);
        if (s == SIZE_MAX) break;
        if (P->whole_word && !whole_word(text, n, s, s + m)) { i = s + 1; continue; } /* krep.c:1323 */
        bool bumped = false;
        if (P->count_lines_mode)
        {
            size_t ls = line_start(text, n, s);
            if (ls != last_line)
            {
                cnt++; last_line = ls; bumped = true;
                if (cnt >= P->max_count) break;                      /* krep.c:1340 */
                size_t le = line_end(text, n, ls);
                size_t nx = le < n ? le + 1 : n;
                if (nx > s) { i = nx; continue; }                    /* krep.c:1346 */
            }
        }
        else
        {
            cnt++; bumped = true;
            if (P->track_positions && res && cnt <= P->max_count) push(res, s, s + m);
        }
        if (bumped && cnt >= P->max_count) break;                    /* krep.c:1366 */
        i = (g_only_matching && !P->count_lines_mode) ? s + m : s + 1;
    }
    return cnt;
}

/* ==========================================================================
 * kmp_search — krep.c:1628-1767.  Non-overlapping: after any full match the
 * automaton restarts at the match end, also when -w rejected it (krep.c:1686).
 * On the limit it appends one extra position before stopping (krep.c:1719).
 * ========================================================================== */
uint64_t oracle_kmp_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    if (P->max_count == 0) return 0; /* krep.c:1634 */
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    size_t m = P->pattern_len;
    if (m == 0 || n < m) return 0;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, i = 0;
    for (;;)
    {
        size_t s = next_occ(t, n, p, m, P->case_sensitive, i);
        if (s == SIZE_MAX) break;
        i = s + m; /* krep.c:1741 / 1686: restart after the match */
        if (P->whole_word && !whole_word(text, n, s, s + m)) continue;
        if (P->count_lines_mode)
        {
            size_t ls = line_start(text, n, s);
            if (ls != last_line)
            {
                if (P->max_count != SIZE_MAX && cnt >= P->max_count) break;
                cnt++; last_line = ls;
                size_t le = line_end(text, n, ls);
                i = le < n ? le + 1 : n; /* krep.c:1707 */
            }
        }
        else
        {
            if (P->max_count != SIZE_MAX && cnt >= P->max_count)
            {
                if (P->track_positions && res) push(res, s, s + m); /* the extra one */
                break;
            }
            cnt++;
            if (P->track_positions && res) push(res, s, s + m);
        }
    }
    return cnt;
}

/* ==========================================================================
 * memchr_search — krep.c:3891-4041 (pattern_len 1; reads pattern[0] only).
 * Positions pass through a 4096-entry local buffer whose final flush is
 * clipped against max_count (krep.c:4026-4038); the clipping and the direct
 * add when the buffer is full at the limit are reproduced.
 * ========================================================================== */
uint64_t oracle_memchr_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    if (P->max_count == 0) return 0;
    const unsigned char *t = (const unsigned char *)text;
    unsigned char a = (unsigned char)P->pattern[0], b = a;
    if (!P->case_sensitive)
    {
        if (a >= 'a' && a <= 'z') b = (unsigned char)(a - 32);
        else if (a >= 'A' && a <= 'Z') b = (unsigned char)(a + 32);
    }
    enum { BUF = 4096 };
    match_position_t *buf = (match_position_t *)malloc(BUF * sizeof *buf);
    size_t nb = 0;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, pos = 0;
    while (pos < n)
    {
        size_t s = pos;
        while (s < n && t[s] != a && t[s] != b) s++;
        if (s >= n) break;
        if (P->whole_word && !whole_word(text, n, s, s + 1)) { pos = s + 1; continue; }
        if (P->count_lines_mode)
        {
            size_t ls = line_start(text, n, s);
            if (ls != last_line)
            {
                if (P->max_count != SIZE_MAX && cnt >= P->max_count) break;
                cnt++; last_line = ls;
                size_t le = line_end(text, n, ls);
                pos = le < n ? le + 1 : n;
            }
            else pos = s + 1;
        }
        else
        {
            bool tracking = P->track_positions && res;
            if (P->max_count != SIZE_MAX && cnt >= P->max_count)
            {
                if (tracking)
                {
                    if (nb < BUF) { buf[nb].start_offset = s; buf[nb].end_offset = s + 1; nb++; }
                    else push(res, s, s + 1);
                }
                break;
            }
            cnt++;
            if (tracking)
            {
                if (nb >= BUF)
                {
                    for (size_t k = 0; k < nb; k++) push(res, buf[k].start_offset, buf[k].end_offset);
                    nb = 0;
                }
                buf[nb].start_offset = s; buf[nb].end_offset = s + 1; nb++;
            }
            pos = s + 1;
        }
    }
    if (P->track_positions && res && nb > 0)
    {
        uint64_t have = res->count;
        uint64_t room = (P->max_count == SIZE_MAX) ? nb : (have >= P->max_count ? 0 : P->max_count - have);
        size_t lim = nb < room ? nb : (size_t)room;
        for (size_t k = 0; k < lim; k++) push(res, buf[k].start_offset, buf[k].end_offset);
    }
    free(buf);
    return cnt;
}

/* ==========================================================================
 * memchr_short_search — krep.c:4371-4503 (pattern_len 2..3).
 * Walks first-byte hits.  With only_matching the advance after ANY first-byte
 * hit (full match or not) is pattern_len (krep.c:4495), so occurrences that
 * start inside the skipped bytes are not seen; reproduced as is.
 * ========================================================================== */
uint64_t oracle_memchr_short_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return 0;
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    size_t m = P->pattern_len;
    bool cs = P->case_sensitive;
    if (m < 2 || m > 3 || n < m) return 0;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, cur = 0;
    while (n - cur >= m)
    {
        size_t h = cur, lim = n - m; /* candidates are cur..lim inclusive */
        while (h <= lim && !(cs ? t[h] == p[0] : lc(t[h]) == lc(p[0]))) h++;
        if (h > lim) break;
        if (occurs(t + h + 1, p + 1, m - 1, cs))
        {
            if (P->whole_word && !whole_word(text, n, h, h + m)) { cur = h + 1; continue; }
            bool bumped = false;
            if (P->count_lines_mode)
            {
                size_t ls = line_start(text, n, h);
                if (ls != last_line)
                {
                    cnt++; last_line = ls; bumped = true;
                    if (cnt >= P->max_count) break;
                    size_t le = line_end(text, n, ls);
                    size_t nx = le < n ? le + 1 : n;
                    if (nx > cur) { cur = nx; continue; }
                }
            }
            else
            {
                cnt++; bumped = true;
                if (P->track_positions && res && cnt <= P->max_count) push(res, h, h + m);
            }
            if (bumped && cnt >= P->max_count) break;
        }
        size_t adv = (h - cur) + (g_only_matching ? m : 1);
        if (adv > n - cur) break;
        cur += adv;
    }
    return cnt;
}

/* ==========================================================================
 * simd_sse42_search — krep.c:4702-4869 (case-sensitive, pattern_len <= 16).
 * _mm_cmpestri(EQUAL_ORDERED) over a window of min(16, remaining) bytes returns
 * the first index at which the pattern matches fully or as a prefix cut by the
 * window end; a hit is accepted only if it is a full match (krep.c:4761) —
 * full matches have smaller indices than cut ones, so that is "the first full
 * match inside the window" — otherwise the window slides by chunk-m+1
 * (krep.c:4858).  After an accepted OR -w-rejected occurrence the cursor moves
 * to match+m (default) or match+1 (only_matching) — krep.c:4839-4848.
 * The window position is observable in -c mode: after counting a line the
 * cursor advances by (line_end+1 - match_start) FROM THE WINDOW START
 * (krep.c:4791-4795: the advance is computed from the match offset but added to
 * current_pos), i.e. it lands `index` bytes before the next line, so the windows
 * are restated literally.
 * Any other precondition falls back to boyer_moore_search (krep.c:4708).
 * ========================================================================== */
uint64_t oracle_sse42_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    size_t m = P->pattern_len;
    if (m == 0 || m > 16 || !P->case_sensitive || n < m) return oracle_boyer_moore_search(P, text, n, res);
    if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return 0;
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, cur = 0, rem = n;
    while (rem >= m)
    {
        const size_t chunk = rem < 16 ? rem : 16;
        size_t idx = SIZE_MAX;
        for (size_t i = 0; i + m <= chunk; i++)
            if (occurs(t + cur + i, p, m, true)) { idx = i; break; }
        if (idx == SIZE_MAX)
        {
            size_t adv = chunk > m ? chunk - m + 1 : 1; /* krep.c:4858 */
            if (adv > rem) adv = rem;
            cur += adv; rem -= adv;
            continue;
        }
        const size_t s = cur + idx;
        if (!P->whole_word || whole_word(text, n, s, s + m))
        {
            bool bumped = false;
            if (P->count_lines_mode)
            {
                size_t ls = line_start(text, n, s);
                if (ls != last_line)
                {
                    if (cnt >= P->max_count) break;
                    cnt++; last_line = ls; bumped = true;
                    size_t le = line_end(text, n, ls);
                    if (le < n)
                    {
                        size_t adv = (le + 1) - s; /* > 0: no newline in [line_start, s) */
                        cur += adv; rem -= adv;     /* krep.c:4795: added to the window start, not to the match */
                        continue;
                    }
                }
            }
            else
            {
                if (cnt >= P->max_count) break;
                cnt++; bumped = true;
                if (P->track_positions && res && cnt <= P->max_count) push(res, s, s + m);
            }
            if (bumped && cnt >= P->max_count) break;
        }
        size_t adv = idx + 1;
        if (!g_only_matching)
        {
            adv = idx + m;
            if (adv > rem) adv = rem;
        }
        cur += adv; rem -= adv;
    }
    return cnt;
}

/* ==========================================================================
 * simd_avx2_search (krep.c:4877-5101, 17..32-byte needles) and
 * simd_avx512_search (krep.c:5108-5286, 33..64-byte needles).
 * Both walk W-byte windows (W = 32 / 64) from a cursor that starts at 0 and
 * advances by W — or, in -c mode, jumps to the start of the next line after a
 * counted line (krep.c:4996-5014 / 5211-5229).  Inside a window EVERY start
 * whose first and last byte match is memcmp-verified in ascending order, so
 * occurrences overlap freely whatever -o says.  What is observable beyond that:
 *   - the tail (< W bytes) is handed to a sub-search on the sub-buffer that
 *     starts at the cursor: boyer_moore_search for AVX2 (krep.c:5068), and
 *     simd_avx2_search -> boyer_moore_search (needle > 32) for AVX-512
 *     (krep.c:5268).  -w and -c are evaluated against the SUB-buffer there
 *     (no byte before its first byte; line starts clipped to it), -o takes
 *     BMH's pattern_len advance, and -m is re-based;
 *   - the tail's positions are re-based by index arithmetic on the result
 *     vector (krep.c:5072-5088 / 5271-5281), reproduced literally;
 *   - AVX-512 skips a window unverified when fewer than (m-1)+64 bytes remain
 *     (krep.c:5171, falls through to 5255) — matches there are lost.
 * A candidate whose last byte would lie past the buffer compares against the
 * zero padding of the safe buffer (krep.c:4944-4956): never a hit for needles
 * that do not end in NUL, which is what is restated here.
 * ========================================================================== */
static uint64_t window_search(const search_params_t *P, const char *text, size_t n, match_result_t *res, size_t W)
{
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    const size_t m = P->pattern_len, maxc = P->max_count;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, cur = 0, rem = n;
    while (rem >= W)
    {
        bool line_skipped = false;
        if (W == 64 && rem < (m - 1) + 64) { cur += 64; rem -= 64; continue; } /* krep.c:5171 */
        for (size_t i = 0; i < W; i++)
        {
            const size_t s = cur + i;
            if (s + m > n || !occurs(t + s, p, m, true)) continue;
            if (P->whole_word && !whole_word(text, n, s, s + m)) continue;
            bool bumped = false;
            if (P->count_lines_mode)
            {
                const size_t ls = line_start(text, n, s);
                if (ls != last_line)
                {
                    cnt++; last_line = ls; bumped = true;
                    if (cnt >= maxc) return cnt;
                    const size_t le = line_end(text, n, ls);
                    const size_t nx = le < n ? le + 1 : n;
                    if (nx > cur)
                    {
                        size_t adv = nx - cur;
                        if (adv > rem) adv = rem;
                        cur += adv; rem -= adv;
                        line_skipped = true;
                        break;
                    }
                }
            }
            else
            {
                cnt++; bumped = true;
                if (P->track_positions && res && cnt <= maxc) push(res, s, s + m);
            }
            if (bumped && cnt >= maxc) return cnt;
        }
        if (line_skipped) continue;
        cur += W; rem -= W;
    }
    if (rem >= m)
    {
        search_params_t tail = *P;
        if (maxc != SIZE_MAX) tail.max_count = cnt >= maxc ? 0 : maxc - cnt;
        const uint64_t tc = oracle_boyer_moore_search(&tail, text + cur, rem, res);
        if (res && P->track_positions && tc > 0)
        {
            if (W == 32)
            {
                uint64_t b0 = cnt > res->count ? res->count : cnt; /* krep.c:5077-5079 */
                for (uint64_t k = b0; k < res->count; k++)
                {
                    res->positions[k].start_offset += cur;
                    res->positions[k].end_offset += cur;
                }
            }
            else
            {
                uint64_t b0 = res->count >= tc ? res->count - tc : 0; /* krep.c:5275 */
                for (uint64_t k = 0; k < tc && b0 + k < res->count; k++)
                {
                    res->positions[b0 + k].start_offset += cur;
                    res->positions[b0 + k].end_offset += cur;
                }
            }
        }
        cnt += tc;
        if (W == 32 && maxc != SIZE_MAX && cnt > maxc) cnt = maxc; /* krep.c:5092 */
    }
    return cnt;
}

uint64_t oracle_avx2_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    const size_t m = P->pattern_len;
    if (m == 0 || m > 32 || !P->case_sensitive || n < m) return oracle_boyer_moore_search(P, text, n, res); /* krep.c:4883 */
    if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return 0;
    if (m <= 16) return oracle_sse42_search(P, text, n, res); /* krep.c:4892 */
    return window_search(P, text, n, res, 32);
}

/* As compiled into an AVX-512 build of the reference (Makefile:34-35). */
uint64_t oracle_avx512_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    const size_t m = P->pattern_len;
    if (m == 0 || m > 64 || !P->case_sensitive || n < m) return oracle_avx2_search(P, text, n, res); /* krep.c:5115 */
    if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return 0;
    if (m <= 32) return oracle_avx2_search(P, text, n, res); /* krep.c:5123 */
    return window_search(P, text, n, res, 64);
}

/* ==========================================================================
 * neon_search — krep.c:4506-4694 (the ARM build's kernel for every
 * case-sensitive needle; pinned against the reference's own source compiled
 * here with scalar stand-ins for its five NEON intrinsics, build_oracle.py).
 * 16-byte windows from a cursor; every start in the window whose first byte
 * matches is memcmp-verified in ascending order (overlaps kept, any needle
 * length).  Differs from the AVX2 walk in three observable ways: the -m limit
 * is tested BEFORE counting as well (krep.c:4571, 4601), the -c jump to the next
 * line only happens when the line has a newline (krep.c:4578), and the tail's
 * positions are re-based over the last tail_count entries (krep.c:4673-4680).
 * ========================================================================== */
uint64_t oracle_neon_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    const size_t m = P->pattern_len, maxc = P->max_count;
    if (m == 0 || !P->case_sensitive || n < m) return oracle_boyer_moore_search(P, text, n, res);
    if (maxc == 0 && (P->count_lines_mode || P->track_positions)) return 0;
    const unsigned char *t = (const unsigned char *)text, *p = (const unsigned char *)P->pattern;
    uint64_t cnt = 0;
    size_t last_line = SIZE_MAX, cur = 0, rem = n;
    while (rem >= 16)
    {
        bool jumped = false;
        for (size_t i = 0; i < 16; i++)
        {
            if (t[cur + i] != p[0] || rem - i < m || !occurs(t + cur + i, p, m, true)) continue;
            const size_t s = cur + i;
            if (P->whole_word && !whole_word(text, n, s, s + m)) continue;
            bool bumped = false;
            if (P->count_lines_mode)
            {
                const size_t ls = line_start(text, n, s);
                if (ls != last_line)
                {
                    if (cnt >= maxc) return cnt;
                    cnt++; last_line = ls; bumped = true;
                    const size_t le = line_end(text, n, ls);
                    if (le < n && le + 1 > cur)
                    {
                        size_t adv = le + 1 - cur;
                        if (adv > rem) adv = rem;
                        cur += adv; rem -= adv;
                        jumped = true;
                        break;
                    }
                }
            }
            else
            {
                if (cnt >= maxc) return cnt;
                cnt++; bumped = true;
                if (P->track_positions && res && cnt <= maxc) push(res, s, s + m);
            }
            if (bumped && cnt >= maxc) return cnt;
        }
        if (jumped) continue;
        cur += 16; rem -= 16;
    }
    if (rem >= m)
    {
        search_params_t tail = *P;
        if (maxc != SIZE_MAX) tail.max_count = cnt >= maxc ? 0 : maxc - cnt;
        const uint64_t tc = oracle_boyer_moore_search(&tail, text + cur, rem, res);
        if (res && P->track_positions && tc > 0 && res->count >= tc)
            for (uint64_t k = 0; k < tc; k++)
            {
                res->positions[res->count - tc + k].start_offset += cur;
                res->positions[res->count - tc + k].end_offset += cur;
            }
        cnt += tc;
    }
    return cnt;
}

/* ==========================================================================
 * ac_trie_build / aho_corasick_search — aho_corasick.c:111-271, 299-466.
 * Restated with array-indexed nodes.  Emission order: ascending end offset;
 * at one end offset the deepest node first, then along the failure chain;
 * inside a node pattern-list order (aho_corasick.c:353-431).  Outputs are not
 * merged along failure links at build time — the search walks the chain.
 * ========================================================================== */
typedef struct
{
    int32_t next[256];
    int32_t fail;
    int32_t *out;
    int32_t nout, cap;
} onode_t;

struct oracle_ac
{
    onode_t *nodes;
    int32_t n, cap;
    bool cs;
};

static int32_t onode_new(struct oracle_ac *A)
{
    if (A->n == A->cap)
    {
        A->cap = A->cap ? A->cap * 2 : 64;
        A->nodes = (onode_t *)realloc(A->nodes, (size_t)A->cap * sizeof(onode_t));
    }
    onode_t *nd = &A->nodes[A->n];
    memset(nd->next, 0xff, sizeof nd->next); /* -1 */
    nd->fail = 0; nd->out = NULL; nd->nout = nd->cap = 0;
    return A->n++;
}
static void onode_out(onode_t *nd, int32_t idx)
{
    if (nd->nout == nd->cap)
    {
        nd->cap = nd->cap ? nd->cap * 2 : 4;
        nd->out = (int32_t *)realloc(nd->out, (size_t)nd->cap * sizeof(int32_t));
    }
    nd->out[nd->nout++] = idx;
}

struct oracle_ac *oracle_ac_build(const search_params_t *P)
{
    if (!P || P->num_patterns == 0) return NULL;
    struct oracle_ac *A = (struct oracle_ac *)calloc(1, sizeof *A);
    A->cs = P->case_sensitive;
    onode_new(A); /* root = 0, fail = itself */
    for (size_t k = 0; k < P->num_patterns; k++)
    {
        const unsigned char *p = (const unsigned char *)P->patterns[k];
        size_t m = P->pattern_lens[k];
        int32_t cur = 0;
        for (size_t i = 0; i < m; i++)
        {
            unsigned char c = A->cs ? p[i] : lc(p[i]);
            if (A->nodes[cur].next[c] < 0)
            {
                int32_t nn = onode_new(A);
                A->nodes[cur].next[c] = nn;
            }
            cur = A->nodes[cur].next[c];
        }
        onode_out(&A->nodes[cur], (int32_t)k); /* empty pattern -> root (aho_corasick.c:145) */
    }
    /* BFS failure links */
    int32_t *q = (int32_t *)malloc((size_t)A->n * sizeof(int32_t));
    int32_t qh = 0, qt = 0;
    for (int c = 0; c < 256; c++)
        if (A->nodes[0].next[c] >= 0) { A->nodes[A->nodes[0].next[c]].fail = 0; q[qt++] = A->nodes[0].next[c]; }
    while (qh < qt)
    {
        int32_t u = q[qh++];
        for (int c = 0; c < 256; c++)
        {
            int32_t v = A->nodes[u].next[c];
            if (v < 0) continue;
            q[qt++] = v;
            int32_t f = A->nodes[u].fail;
            while (f != 0 && A->nodes[f].next[c] < 0) f = A->nodes[f].fail;
            A->nodes[v].fail = A->nodes[f].next[c] >= 0 ? A->nodes[f].next[c] : 0;
        }
    }
    free(q);
    return A;
}
void oracle_ac_free(struct oracle_ac *A)
{
    if (!A) return;
    for (int32_t i = 0; i < A->n; i++) free(A->nodes[i].out);
    free(A->nodes);
    free(A);
}
bool oracle_ac_root_has_outputs(const struct oracle_ac *A) { return A && A->n > 0 && A->nodes[0].nout > 0; }

/* params->ac_trie must hold a struct oracle_ac* built by oracle_ac_build. */
uint64_t oracle_aho_corasick_search(const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    if (!P || !P->ac_trie || !text) return 0;
    if (P->max_count == 0) return 0;
    const struct oracle_ac *A = (const struct oracle_ac *)P->ac_trie;
    const size_t maxc = P->max_count;
    uint64_t found = 0;
    size_t last_line = SIZE_MAX;
    int32_t cur = 0;
    for (size_t i = 0; i < n; i++)
    {
        unsigned char c = (unsigned char)text[i];
        if (!P->case_sensitive) c = lc(c);
        while (cur != 0 && A->nodes[cur].next[c] < 0) cur = A->nodes[cur].fail;
        if (A->nodes[cur].next[c] >= 0) cur = A->nodes[cur].next[c];
        for (int32_t o = cur; o != 0; o = A->nodes[o].fail)
        {
            const onode_t *nd = &A->nodes[o];
            for (int32_t j = 0; j < nd->nout; j++)
            {
                if (found >= maxc) return found;
                size_t k = (size_t)nd->out[j], m = P->pattern_lens[k];
                if (m == 0) continue;
                size_t s = i + 1 - m, e = i + 1;
                if (P->whole_word && !whole_word(text, n, s, e)) continue;
                if (P->count_lines_mode)
                {
                    size_t ls = line_start(text, n, s);
                    if (ls != last_line)
                    {
                        found++; last_line = ls;
                        if (found >= maxc) return found;
                    }
                }
                else
                {
                    found++;
                    if (P->track_positions && res) push(res, s, e);
                    if (found >= maxc) return found;
                }
            }
            if (found >= maxc) return found;
        }
    }
    if (n == 0 && A->nodes[0].nout > 0) /* aho_corasick.c:442-463 */
    {
        for (int32_t j = 0; j < A->nodes[0].nout; j++)
            if (P->pattern_lens[A->nodes[0].out[j]] == 0)
            {
                if (found < maxc)
                {
                    found++;
                    if (P->track_positions && res) push(res, 0, 0);
                }
                break;
            }
    }
    return found;
}
