/* krep_b200.h — C ABI of the B200-native scan engine that drops in behind krep's
 * search_func_t boundary.
 *
 * Every entry point below names the reference interface it replaces as
 * (file:line) into davidesantangelo/krep v2.2.0.  The data types are restated
 * byte-for-byte from krep.h:49-101 so that a krep host can pass its own
 * search_params_t / match_result_t straight through; if krep.h was included
 * first (KREP_H defined) the restatement is skipped and krep's own types are used.
 *
 * Nothing in this header mentions torch, CUDA runtime types or C++: plain
 * pointers and sizes only.  Device pointers are passed as const void* and
 * streams as void* (a cudaStream_t).
 */
#ifndef KREP_B200_H
#define KREP_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Types restated from krep.h (layout-identical; checked by tests/test_abi.py, */
/* which also drives the compiled reference with the very same structs)       */
/* ------------------------------------------------------------------------- */
#ifndef KREP_H

/* krep.h:49-53 */
typedef struct
{
   size_t start_offset; /* first byte of the match, relative to text_start   */
   size_t end_offset;   /* one past the last byte                             */
} match_position_t;

/* krep.h:55-60 — positions must be malloc-family memory (krep.c:244-251)     */
typedef struct match_result_t
{
   match_position_t *positions;
   uint64_t count;
   uint64_t capacity;
} match_result_t;

struct ac_trie;
typedef struct ac_trie ac_trie_t; /* krep.h:22-23, opaque */

/* krep.h:65-94 */
typedef struct search_params
{
   const char *pattern; /* single-literal kernels read these two (krep.c:1271) */
   size_t pattern_len;

   const char **patterns; /* multi-literal kernel reads these three            */
   size_t *pattern_lens;
   size_t num_patterns;

   bool case_sensitive;
   bool use_regex;
   bool count_lines_mode;   /* -c   */
   bool count_matches_mode; /* -co  */
   bool track_positions;    /* !(-c && !-o) */
   bool whole_word;         /* -w   */

   const void *compiled_regex; /* const regex_t* in krep.h; unused on this path */
   ac_trie_t *ac_trie;
   size_t max_count; /* SIZE_MAX = unlimited */
} search_params_t;

/* krep.h:98-101 */
typedef uint64_t (*search_func_t)(const search_params_t *params,
                                  const char *text_start,
                                  size_t text_len,
                                  match_result_t *result);
#endif /* KREP_H */

/* ------------------------------------------------------------------------- */
/* Lifetime                                                                  */
/* ------------------------------------------------------------------------- */

/* Make `device` the process's primary CUDA device and create its engine context
 * (streams, result buffers).  Optional: every entry point initialises lazily,
 * with the calling thread's current device as the primary one.  One process can
 * drive several devices: host-text searches spread over the devices chosen by
 * krep_b200_set_devices / KREP_B200_DEVICES (see below), and the resident-shard
 * API runs each shard on the device that owns its memory.  Returns 0, or a
 * negative value after printing "krep: ..." to stderr (the reference's error
 * convention, krep.c:1933).  There is no CPU fallback: without a usable sm_100
 * device every search entry point prints an error and aborts the call with
 * count 0 and krep_b200_last_error() != 0. */
int krep_b200_init(int device);
int krep_b200_device_count(void);          /* CUDA devices visible to the process */
/* Non-blocking: starts CUDA initialisation and the primary device's context on a background thread and returns.  A
 * host calls it as soon as it knows a literal search is coming (krep: after option parsing, before search_file opens
 * and maps the file, krep.c:3818), so the driver start-up overlaps the host's own file handling; the first entry point
 * that needs the GPU waits for it (and meanwhile pre-faults the caller's text with the staging threads). */
void krep_b200_warmup(void);
/* The devices a search_func_t call spreads the caller's text over — the analogue of krep's thread count
 * (krep.c:2851-2905 cuts the file into one chunk per pool thread; here into one contiguous range per GPU, each range
 * streamed over that GPU's own PCIe link and scanned there, per-device occurrence lists merged by key on the host).
 * devices == NULL or n == 0 restores the default: KREP_B200_DEVICES=<count> from the environment, else ONE device (from a
 * pageable file mapping more devices add nothing: the host's page faults and staging copies are the limit; they pay for
 * pinned text).  Start-up note: a process in which this library is the first user of CUDA hides the GPUs it will not use
 * from the driver before CUDA initialises (cuInit enumerates every visible GPU: 6.8 s on an 8-GPU box against 0.4 s for
 * one) — so choose the devices (this call, krep_b200_init, or KREP_B200_DEVICES / KREP_B200_KEEP_VISIBLE in the
 * environment) before the first search. */
void krep_b200_set_devices(const int *devices, int n);
void krep_b200_shutdown(void);
int krep_b200_last_error(void);           /* 0 = last call succeeded          */
const char *krep_b200_last_error_string(void);
const char *krep_b200_version(void);

/* ------------------------------------------------------------------------- */
/* The file-static globals of krep.c that the kernels read (krep.c:117-120). */
/* A host that links this library mirrors its own flags into these.          */
/* ------------------------------------------------------------------------- */
void krep_b200_set_only_matching(bool on); /* -o            krep.c:117 */
bool krep_b200_get_only_matching(void);
void krep_b200_set_force_no_simd(bool on); /* --no-simd     krep.c:118 */
void krep_b200_set_algo_override(const char *name); /* --algo=auto|bm|kmp krep.c:120; NULL = auto */

/* ------------------------------------------------------------------------- */
/* search_func_t replacements (host text in, match_result_t out).            */
/* Each reproduces the count, the offsets and their order of the named       */
/* reference function called once on the whole buffer (krep -t 1 semantics,  */
/* SURVEY §8 a12), including its overlap policy, -w/-c/-m behaviour.          */
/* ------------------------------------------------------------------------- */
uint64_t krep_b200_boyer_moore_search(const search_params_t *, const char *, size_t, match_result_t *);  /* krep.c:1260 */
uint64_t krep_b200_kmp_search(const search_params_t *, const char *, size_t, match_result_t *);          /* krep.c:1628 */
uint64_t krep_b200_memchr_search(const search_params_t *, const char *, size_t, match_result_t *);       /* krep.c:3891 */
uint64_t krep_b200_memchr_short_search(const search_params_t *, const char *, size_t, match_result_t *); /* krep.c:4371 */
uint64_t krep_b200_simd_sse42_search(const search_params_t *, const char *, size_t, match_result_t *);   /* krep.c:4702 */
uint64_t krep_b200_simd_avx2_search(const search_params_t *, const char *, size_t, match_result_t *);    /* krep.c:4877 */
uint64_t krep_b200_simd_avx512_search(const search_params_t *, const char *, size_t, match_result_t *);  /* krep.c:5108 */
uint64_t krep_b200_aho_corasick_search(const search_params_t *, const char *, size_t, match_result_t *); /* aho_corasick.c:299 */
/* the ARM build's kernel; never chosen by krep_b200_select_search_algorithm (which stands in for the x86 AVX2 build) */
uint64_t krep_b200_neon_search(const search_params_t *, const char *, size_t, match_result_t *);          /* krep.c:4506 */

/* Many texts, one launch — what search_directory_recursive (krep.c:3310) calling search_file once per small file
 * becomes when the per-call copy and launch latency matters.  `entry` is one of the ten functions above; text i gets
 * exactly the count (counts[i]) and positions (results[i], may be NULL, or results == NULL) that
 * entry(params, texts[i], lens[i], results[i]) would have produced.  Returns 0, or a negative error. */
int krep_b200_search_batch(search_func_t entry, const search_params_t *params, const char *const *texts,
                           const size_t *lens, size_t n_texts, uint64_t *counts, match_result_t *const *results);

/* krep.c:1771 — same decision order (regex excluded: returns NULL for
 * use_regex, the caller keeps its own regex_search), same globals. The
 * returned pointer is one of the eight functions above. */
search_func_t krep_b200_select_search_algorithm(const search_params_t *params);
/* krep.c:1964 */
const char *krep_b200_get_algorithm_name(search_func_t func);

/* aho_corasick.c:111 / 274 / 287.  The returned object is this library's own
 * device automaton (patterns folded, filter tables and verify tables resident
 * in HBM); krep_b200_aho_corasick_search also accepts a params->ac_trie that
 * was built by the reference's ac_trie_build (it only tests it for NULL, as
 * aho_corasick.c:306 does) and then compiles and caches its own automaton
 * from params->patterns. */
ac_trie_t *krep_b200_ac_trie_build(const search_params_t *params);
void krep_b200_ac_trie_free(ac_trie_t *trie);
bool krep_b200_ac_trie_root_has_outputs(const ac_trie_t *trie);

/* krep.c:139 / 175 / 244 / 256 — for hosts that do not link krep.c. */
match_result_t *krep_b200_match_result_init(uint64_t initial_capacity);
bool krep_b200_match_result_add(match_result_t *result, size_t start_offset, size_t end_offset);
void krep_b200_match_result_free(match_result_t *result);
bool krep_b200_match_result_merge(match_result_t *dest, const match_result_t *src, size_t chunk_offset);

/* ------------------------------------------------------------------------- */
/* HBM-resident shard API — what search_chunk_thread (krep.c:1919) becomes   */
/* when the chunk already lives on the GPU: one shard per device, owned      */
/* range + halo, matches owned by start offset.                 */
/* ------------------------------------------------------------------------- */

/* Emulated reference kernel: selects the overlap / -w / -m policy applied to
 * the raw occurrence list. */
enum
{
   KREP_B200_ALGO_BMH = 0,          /* boyer_moore_search  krep.c:1260 */
   KREP_B200_ALGO_KMP = 1,          /* kmp_search          krep.c:1628 */
   KREP_B200_ALGO_MEMCHR = 2,       /* memchr_search       krep.c:3891 */
   KREP_B200_ALGO_MEMCHR_SHORT = 3, /* memchr_short_search krep.c:4371 */
   KREP_B200_ALGO_SSE42 = 4,        /* simd_sse42_search   krep.c:4702 */
   KREP_B200_ALGO_AVX2 = 5,         /* simd_avx2_search    krep.c:4877 */
   KREP_B200_ALGO_AVX512 = 6,       /* simd_avx512_search  krep.c:5108 */
   KREP_B200_ALGO_AC = 7,           /* aho_corasick_search aho_corasick.c:299 */
   KREP_B200_ALGO_NEON = 8          /* neon_search         krep.c:4506 */
};

typedef struct krep_b200_plan krep_b200_plan_t; /* compiled pattern set, device-resident */

/* Compile params->pattern (algo != AC) or params->patterns[] (algo == AC)
 * into filter constants / tables on the current device. NULL on error. */
krep_b200_plan_t *krep_b200_plan_create(const search_params_t *params, int algo);
void krep_b200_plan_destroy(krep_b200_plan_t *plan);
/* Which device filter the plan uses (for bench/config reporting). */
const char *krep_b200_plan_filter_name(const krep_b200_plan_t *plan);

/* One shard of a corpus that is resident in device memory. */
typedef struct
{
   const void *d_text;     /* device pointer, 16-byte aligned                       */
   uint64_t avail_len;     /* bytes readable at d_text: owned range + halo            */
   uint64_t own_begin;     /* report matches whose start is in [own_begin, own_end)   */
   uint64_t own_end;       /*   (offsets relative to d_text)                          */
   uint64_t global_offset; /* added to every reported offset                          */
   int32_t prev_byte;      /* byte preceding d_text[0] in the whole text, -1 = none   */
   int32_t next_byte;      /* byte following d_text[avail_len-1], -1 = end of text    */
} krep_b200_shard_t;

/* Result of a device scan, left in device memory (sorted ascending). */
typedef struct
{
   uint64_t count;          /* occurrences found (exact even if capacity was exceeded) */
   uint64_t stored;         /* entries actually stored = min(count, capacity)           */
   const uint64_t *d_keys;  /* device: literal: start offset (global); AC: packed key   */
   int overflow;            /* 1 if count > capacity: call again with a larger capacity */
   uint64_t text_len;       /* global_offset + avail_len of the scanned shard: the length of the whole text when
                               the shard is the last (or only) one — what the window kernels' tail logic needs */
   const uint64_t *d_line_bounds; /* device, 2 words per stored key, only for plans created with count_lines_mode (-c):
                               global offset of the first byte of the occurrence's line, and of the line's newline (or
                               the text length) — find_line_start / find_line_end (krep.c:363, 401) computed on the GPU */
   int32_t device;          /* CUDA device that holds the lists                                                      */
   int32_t slot;            /* which of the device's result buffers the scan used                                    */
   uint64_t serial;         /* scan number on that device: the lists stay valid until the next scan of that device   */
} krep_b200_device_result_t;

/* Scan one shard on `stream` (cudaStream_t, NULL = engine stream) of the device that owns shard->d_text: launches
 * the filter+verify kernel and sorts the occurrence list on the device (lists of up to 16 384 occurrences by the
 * one-CTA finish kernel, which also hands count and list to the host in the scan's single synchronisation; longer
 * ones by a radix sort that is still running on `stream` when the call returns — krep_b200_collect /
 * krep_b200_export_* order themselves after it). For a
 * literal plan every key is (global start offset << 3 | tag bits) of one occurrence
 * that passed the plan's -w filter (tags: see csrc/common.h).  For an AC plan every key packs
 * (end_offset << 24 | (1023 - (len-1)) << 10 ... see krep_b200_ac_key_* below) so
 * that ascending key order is aho_corasick_search's emission order.
 * `want_positions` = 0 counts only (no list is written).
 * Blocks until the count is known. Returns 0 or a negative error. */
int krep_b200_scan_shard(const krep_b200_plan_t *plan, const krep_b200_shard_t *shard,
                         int want_positions, void *stream, krep_b200_device_result_t *out);
/* The same scan in two halves: _begin enqueues it and returns a ticket without waiting, _end waits for it.  A host
 * that has work of its own per scan (rank 0 of a multi-GPU job replaying the previous step's gathered list) does it
 * between the two.  At most two scans per device may be in flight, and _end only waits for its own scan: a host that
 * begins scan i+1 before it ends scan i keeps the GPU busy back to back (lists of up to 16 384 occurrences, which come
 * back through pinned memory; a longer list must be ended before the next scan begins). */
int krep_b200_scan_shard_begin(const krep_b200_plan_t *plan, const krep_b200_shard_t *shard,
                               int want_positions, void *stream, int *ticket);
int krep_b200_scan_shard_end(int ticket, krep_b200_device_result_t *out);

/* Copy the first min(out->stored, max_keys) sorted keys of a shard result into another device
 * buffer (device-to-device, on `stream`), e.g. a torch tensor that is then gathered with NCCL. */
int krep_b200_export_keys(const krep_b200_device_result_t *dev, void *d_dst, uint64_t max_keys, void *stream);
/* The row a multi-GPU host gathers: d_dst[0] = the shard's exact occurrence count, d_dst[1..] = its first
 * min(stored, max_keys) sorted keys — one device-to-device copy on `stream`. */
int krep_b200_export_packed(const krep_b200_device_result_t *dev, void *d_dst, uint64_t max_keys, void *stream);
/* The same row for a scan that is still in flight (its ticket): enqueued on the scan's stream behind the finish kernel,
 * always the whole fixed-size row (max_keys <= 16384); a longer list arrives as count > max_keys. */
int krep_b200_export_packed_async(int ticket, void *d_dst, uint64_t max_keys);
/* Merge step of a sharded search (krep.c:2928-3004 without its chunk-edge artefacts): n_lists ascending key lists
 * (one per shard, shards in text order) -> one ascending list in dst (room for the sum of counts; may alias the
 * first list).  Literal keys (ordered and owned by start offset) are already globally ordered after concatenation;
 * pattern-set keys are ordered by END offset (aho_corasick.c:353-431) but owned by START offset, so around every cut
 * a long match owned by the earlier shard can end after a short match owned by the later one — the merge puts them
 * back into emission order.  Returns the total. */
uint64_t krep_b200_merge_keys(const uint64_t *const *lists, const uint64_t *counts, uint32_t n_lists, uint64_t *dst);

/* Timing hook for bench.py: device time in milliseconds of the scan kernel(s)
 * of the most recent krep_b200_scan_shard / search call on this thread,
 * measured with CUDA events on the launching stream (kernel only, no sort). */
float krep_b200_last_kernel_ms(void);
/* Number of launches of this library's own (hand-written) kernels since the last reset; the CUB radix-sort
 * launches behind krep_b200_scan_shard are library code and are not included. */
uint64_t krep_b200_launch_count(void);
void krep_b200_reset_launch_count(void);

/* Fused -c (count_lines_mode) for single literals: the scan itself counts the lines that hold an occurrence
 * (krep.c:1331-1351 and the equivalent branches of the other literal kernels), so only this record leaves the GPU.
 * A shard that cuts lines still gives an exact total: records of shards in text order are folded with
 * krep_b200_combine_line_counts, which subtracts a line counted on both sides of a cut. */
typedef struct
{
   uint64_t lines;    /* lines of this shard that hold an occurrence it owns                                     */
   uint32_t flags;    /* KREP_B200_LINES_*                                                                        */
   uint32_t reserved;
} krep_b200_line_count_t;
enum
{
   KREP_B200_LINES_HAS_HIT = 1,      /* the shard owns at least one occurrence                                       */
   KREP_B200_LINES_FIRST_OPEN = 2,   /* its first occurrence lies before its first newline (the line began earlier)  */
   KREP_B200_LINES_LAST_PENDING = 4, /* no newline between its last occurrence and its end (the line goes on)         */
   KREP_B200_LINES_HAS_NL = 8        /* the shard holds a newline (computed for shards without an occurrence)         */
};
/* Plans created from params with count_lines_mode set; not for pattern sets, needles of 17..64 bytes taken by the
 * window kernels, patterns containing a newline or -w plans in tag mode (those need krep_b200_scan_shard +
 * krep_b200_collect): returns a negative error for them. */
int krep_b200_count_lines_shard(const krep_b200_plan_t *plan, const search_params_t *params, const krep_b200_shard_t *shard,
                                void *stream, krep_b200_line_count_t *out);
uint64_t krep_b200_combine_line_counts(const krep_b200_line_count_t *recs, size_t n, size_t max_count);

/* Apply the emulated reference kernel's policy (overlap rule, -w, -c, -m) to a
 * shard result and deliver it as krep's match_result_t (host, malloc memory).
 * Count-lines mode (-c) uses the line bounds the scan computed on the device
 * (plan created from params with count_lines_mode set, single shard: the
 * shard must not cut a line, i.e. prev_byte = next_byte = -1 or newline-aligned).
 * Returns the count the reference would return. */
uint64_t krep_b200_collect(const krep_b200_plan_t *plan, const search_params_t *params,
                           const krep_b200_device_result_t *dev, match_result_t *result);

/* Several resident shards (text order; on one GPU or spread over the GPUs of this process), one answer: search_file's
 * chunk loop and merge (krep.c:2851-3004) for text that already lives in HBM.  Shards on distinct devices are scanned
 * concurrently, the per-shard lists merged by key, and the policy replayed ONCE over the whole list, so overlap rules,
 * -m and the emission order are those of the reference's single-chunk run.  -c is answered by the fused line count
 * (single literals; line cuts between shards are resolved). */
uint64_t krep_b200_search_shards(const krep_b200_plan_t *plan, const search_params_t *params,
                                 const krep_b200_shard_t *shards, uint32_t n_shards, match_result_t *result);

/* Policy replay over a caller-supplied, ascending occurrence-key list in HOST memory (what
 * krep_b200_collect does after reading the device list back).  A multi-GPU host gathers the
 * per-shard lists, merges them with krep_b200_merge_keys and calls this once.  `text` may be NULL unless params->count_lines_mode is set; `text_len` (the length of the
 * whole text) may be 0 = unknown, except for the AVX2 / AVX-512 window kernels with needles > 16 bytes, whose tail
 * handling (krep.c:5059, 5260) depends on it.
 * `algo` is a KREP_B200_ALGO_* value; `only_matching` is the -o global to emulate. */
uint64_t krep_b200_replay(int algo, const search_params_t *params, bool only_matching,
                          const uint64_t *keys, uint64_t nkeys,
                          const char *text, size_t text_len, match_result_t *result);

/* The same replay without any host text: `bounds` holds two words per key — the global offset of the first byte of
 * the occurrence's line and of that line's newline (or the text length) — as krep_b200_scan_shard computes them on
 * the device for -c plans (krep_b200_device_result_t.d_line_bounds, markers resolved).  A host that gathers keys from
 * several newline-aligned shards gathers the bounds with them. */
uint64_t krep_b200_replay_lines(int algo, const search_params_t *params, bool only_matching,
                                const uint64_t *keys, uint64_t nkeys, const uint64_t *bounds,
                                size_t text_len, match_result_t *result);

/* AC key layout helpers */
uint64_t krep_b200_ac_key_end(uint64_t key);
uint64_t krep_b200_ac_key_start(uint64_t key);
uint32_t krep_b200_ac_key_pattern(uint64_t key);

/* ------------------------------------------------------------------------- */
/* Synthetic corpus: byte[i] is a pure function of (seed, i), so */
/* any shard can be materialised in place on any GPU without transfers.      */
/* ------------------------------------------------------------------------- */
typedef struct
{
   uint64_t seed;            /* text seed                                          */
   uint64_t plant_seed;      /* needle placement seed                              */
   uint64_t plant_period;    /* one planted needle per this many bytes (0 = none)  */
   const char *needle;       /* host pointer, needle_len bytes (copied)            */
   uint32_t needle_len;      /* <= 64                                              */
   uint32_t flags;           /* KREP_B200_CORPUS_*                                 */
} krep_b200_corpus_spec_t;

enum
{
   KREP_B200_CORPUS_RANDOM_CASE = 1, /* planted needles get per-letter pseudo-random case   */
   KREP_B200_CORPUS_EMBED_HALF = 2   /* odd-numbered plants are glued inside a longer word  */
};

/* Fill d_dst[0..len) with corpus bytes [global_offset, global_offset+len). */
int krep_b200_corpus_generate(const krep_b200_corpus_spec_t *spec, void *d_dst,
                              uint64_t global_offset, uint64_t len, void *stream);
/* Host twin of the generator (same bytes), for tests and the CPU baseline. */
int krep_b200_corpus_generate_host(const krep_b200_corpus_spec_t *spec, void *dst,
                                   uint64_t global_offset, uint64_t len);

#ifdef __cplusplus
}
#endif
#endif /* KREP_B200_H */
