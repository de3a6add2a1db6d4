// common.h — internal declarations shared by the engine's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/krep_b200.h"

namespace kb {

// ---------------------------------------------------------------------------------------------
// Occurrence keys.  Every occurrence the device reports is one 64-bit key; ascending key order is
// the order the emulated reference kernel would have produced.
//   literal plans : key = (global_start << 3) | (full << 2) | (ws_ok << 1) | we_ok
//                   full  = all pattern_len bytes match (always 1 unless the plan emits prefix hits,
//                           which only memchr_short_search's -o walk needs, krep.c:4495)
//                   ws_ok / we_ok = the two halves of is_whole_word_match(start, start+len) (krep.h:312):
//                           no word character before the start / after the end — both 1 when -w is off.
//                           Kept apart because the tail sub-search of simd_avx2_search / simd_avx512_search
//                           (krep.c:5068, 5268) cannot see the byte before its sub-buffer.
//   AC plans      : key = (global_end << 24) | ((1023 - (len-1)) << 14) | pattern_index
//                   (end ascending, then longest first, then pattern-list order: aho_corasick.c:353-431)
// ---------------------------------------------------------------------------------------------
static constexpr int LIT_TAG_BITS = 3;
static constexpr int AC_END_SHIFT = 24;
static constexpr int AC_LEN_SHIFT = 14;
static constexpr uint32_t AC_MAX_PATTERNS = 1u << 14;

enum FilterKind : int
{
    FILTER_ALIGNED4 = 0, // pattern_len >= 7: every occurrence contains one aligned 32-bit word; test each
                         // aligned text word against the 4 pattern words P[d..d+4), d = 0..3
    FILTER_WINDOW4 = 1,  // pattern_len < 7 (or prefix plans): test the 4-byte window at every byte offset
};

struct LitDevParams
{
    const uint8_t *text; // 16-byte aligned
    uint64_t avail_len;
    uint64_t own_begin, own_end;
    uint64_t global_offset;
    int32_t prev_byte, next_byte;
    uint64_t group_begin, group_end; // 16-byte groups [group_begin, group_end) are scanned by the vector loop
    uint64_t tail_start;             // starts >= tail_start are checked byte-wise by the tail warp
    uint32_t m;                      // full pattern length
    uint32_t emit_len;               // bytes that must match for a key to be emitted (== m, or 1 for prefix plans)
    uint32_t K[4];                   // filter constants (already folded)
    uint32_t fold;                   // AND-mask applied to text words before comparing (0xFFFFFFFF or 0xDFDFDFDF)
    uint32_t win_mask;               // WINDOW4: mask of the low min(4, emit_len) bytes
    uint32_t mulc[3];                // WINDOW4: 2^24, 2^16, 2^8 — window extraction on the FMA pipe (scan_literal.cu)
    const uint8_t *pat_val;          // device: pattern[k] & pat_mask[k]
    const uint8_t *pat_mask;         // device: 0xDF where case folds, else 0xFF
    uint64_t *out;                   // device key buffer (may be null when !want_positions)
    uint64_t cap;
    unsigned long long *counter;     // [0] = occurrences emitted (exact, also past cap)
    uint32_t whole_word;             // 0 none, 1 drop failures on device, 2 tag only
    uint32_t want_positions;
};

struct AcDevTables;  // scan_multi.cu: one device's copy of a pattern set's tables
struct AcHostTables; // scan_multi.cu: the tables as compiled on the host (uploaded to each device on first use)

static constexpr int MAX_DEV = 16; // CUDA devices one process can drive

// Device-resident half of a plan, one per CUDA device that has run it (uploaded lazily by plan_on_device()).
struct PlanDev
{
    bool ready = false;
    uint8_t *d_pat_val = nullptr, *d_pat_mask = nullptr; // literal
    AcDevTables *ac = nullptr;                           // pattern set
};

struct Plan
{
    int algo = 0;
    bool is_ac = false;
    // literal
    std::string pattern;
    bool case_sensitive = true;
    uint32_t m = 0, emit_len = 0;
    FilterKind filter = FILTER_ALIGNED4;
    uint32_t K[4] = {0, 0, 0, 0};
    uint32_t fold = 0xFFFFFFFFu, win_mask = 0xFFFFFFFFu;
    uint32_t whole_word = 0;
    std::vector<uint8_t> h_val, h_msk; // pattern[k] & mask[k], mask[k] (0xDF where case folds, else 0xFF)
    bool border_free = true; // no proper prefix is also a suffix: occurrences cannot overlap
    bool built_only_matching = false; // value of the -o global the plan was compiled for
    bool count_lines = false;         // -c: scan_shard also computes the line bounds of every occurrence on the device
    // AC
    std::vector<std::string> patterns;
    std::vector<uint32_t> pat_lens;
    uint32_t min_len = 0, max_len = 0;
    AcHostTables *ach = nullptr;
    std::string filter_name;
    PlanDev dev[MAX_DEV];
    uint64_t magic = 0x6b7265705f623230ull; // "krep_b20"
};

// engine.cu
struct DevCtx;
void set_error(int code, const char *fmt, ...);
void clear_error();

struct ScanOut
{
    uint64_t count = 0, stored = 0;
    const uint64_t *d_keys = nullptr;
    int overflow = 0;
    const uint64_t *d_bounds = nullptr; // -c plans: 2 words per stored key (line start, line end), see k_line_bounds
    const uint64_t *h_sorted = nullptr; // host copy of the sorted keys when the list was small enough to come back with the
                                        // count (k_finish, engine.cu); valid until the next scan on the same device slot
    int device = 0;
    uint64_t serial = 0;                // scan number on that device (stale-result detection in krep_b200_collect)
};

// Line bounds of an occurrence, global offsets: [0] = first byte of its line, [1] = position of the line's '\n' (or
// the text length).  The device writes these markers where the answer lies outside what one warp looked at:
static constexpr uint64_t LB_SAME_AS_PREV = ~0ull;     // no newline between the previous occurrence and this one
static constexpr uint64_t LB_SAME_AS_NEXT = ~0ull - 1; // no newline between this occurrence and the next one
static constexpr uint64_t LB_OUTSIDE_SHARD = ~0ull - 2; // the line continues into a neighbouring shard

// Launch one shard scan on `stream` of the device context; appends to that device's key list (no counter reset).
int launch_scan(DevCtx &C, const Plan *plan, const krep_b200_shard_t *sh, int want_positions, cudaStream_t stream, int slot = 0);
// literal kernels (scan_literal.cu)
void launch_literal(const Plan *plan, const LitDevParams &p, int sm_count, cudaStream_t s);
// multi kernels (scan_multi.cu)
int ac_build_tables(Plan *plan);                 // host side only: filter tables, exact table, pattern pool
void ac_free_tables(Plan *plan);                 // host tables (device copies are freed by ac_free_device)
AcDevTables *ac_upload_tables(const Plan *plan); // copies the host tables to the current device; nullptr on CUDA errors
void ac_free_device(AcDevTables *T);
struct AcLaunch
{
    const uint8_t *text;
    uint64_t avail_len, own_begin, own_end, global_offset;
    int32_t prev_byte, next_byte;
    uint64_t *out;
    uint64_t cap;
    unsigned long long *counter;
    uint32_t whole_word, want_positions;
};
void launch_ac(const Plan *plan, const AcDevTables *T, const AcLaunch &a, int sm_count, cudaStream_t s);
void count_launch(int n = 1);

// semantics.cpp — reference control flow replayed over the sorted occurrence list
struct Replay
{
    const uint64_t *keys;
    size_t n;
    const char *text; // host text (for -c line logic); may be null when `bounds` is given or -c is off
    size_t text_len;
    uint64_t base; // global offset subtracted from key offsets
    const uint64_t *bounds = nullptr; // resolved line bounds, 2 per key (device-side -c): used when text is null
};
uint64_t replay_literal(int algo, const search_params_t *P, bool only_matching, uint32_t m,
                        const Replay &r, match_result_t *res);
uint64_t replay_ac(const search_params_t *P, const Replay &r, match_result_t *res);
bool result_push(match_result_t *r, size_t s, size_t e);

// C-locale helpers shared by host code (krep.c:125-134, krep.h:298-301)
static inline unsigned char lower_c(unsigned char c) { return (c >= 'A' && c <= 'Z') ? (unsigned char)(c + 32) : c; }
static inline bool is_alpha_c(unsigned char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
static inline bool is_word_c(int c)
{
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

} // namespace kb
