// host_api.cu — the search_func_t-typed entry points (host text in, match_result_t out) and the rest of
// the host-facing C ABI: option globals, dispatch (select_search_algorithm), AC trie handles,
// match_result helpers.
//
// Data path of one call (north_star): the caller's buffer (krep's mmap, krep.c:2680) is staged into
// HBM in chunks — straight from the caller's memory when it is already page-locked, otherwise through
// a ring of pinned staging buffers filled by host threads — with cudaMemcpyAsync on a copy stream,
// while the scan stream runs the filter kernel on every chunk whose bytes have landed.  All chunk
// kernels append to one device occurrence list, which is sorted on the device, read back, and replayed
// under the emulated kernel's policy (semantics.cpp).  There is no CPU scan anywhere on this path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#include <sys/mman.h>
#include <thread>
#include "common.h"
#include "engine.h"

namespace kb {

// ---- krep.c:117-120 mirrored option globals ---------------------------------------------------
static bool g_only_matching = false;
static bool g_force_no_simd = false;
static std::string g_algo_override; // "", "auto", "bm", "kmp"

// The reference binary this library stands in for is the AVX2 build (Makefile:31-35, krep.c:47-59):
// KREP_USE_AVX2 = KREP_USE_SSE42 = 1, SIMD_MAX_PATTERN_LEN = 32 (krep.c:104-106).
static constexpr size_t SIMD_MAX_PATTERN_LEN = 32;

// Precondition fallbacks of the simd_* entry points (krep.c:4708-4712, 4883-4895, 5115-5126).
int resolve_algo(const search_params_t *P, int algo)
{
    const size_t m = P->pattern_len;
    if (algo == KREP_B200_ALGO_AVX512)
    {
        if (m == 0 || m > 64 || !P->case_sensitive) return KREP_B200_ALGO_BMH;
        if (m <= 32) algo = KREP_B200_ALGO_AVX2;
    }
    if (algo == KREP_B200_ALGO_AVX2)
    {
        if (m == 0 || m > 32 || !P->case_sensitive) return KREP_B200_ALGO_BMH;
        if (m <= 16) algo = KREP_B200_ALGO_SSE42;
    }
    if (algo == KREP_B200_ALGO_SSE42)
    {
        if (m == 0 || m > 16 || !P->case_sensitive) return KREP_B200_ALGO_BMH;
    }
    if (algo == KREP_B200_ALGO_NEON && (m == 0 || !P->case_sensitive)) return KREP_B200_ALGO_BMH; // krep.c:4511
    return algo;
}

// ---- plan cache ---------------------------------------------------------------------------------
static std::vector<Plan *> g_plan_cache; // most recent last; plans are device-independent (uploaded per device on use)

static bool plan_matches(const Plan *pl, const search_params_t *P, int algo, bool only_matching)
{
    if (pl->algo != algo || pl->case_sensitive != P->case_sensitive || pl->count_lines != P->count_lines_mode) return false;
    if (pl->is_ac)
    {
        if ((pl->whole_word != 0) != P->whole_word) return false;
        if (pl->patterns.size() != P->num_patterns) return false;
        for (size_t k = 0; k < P->num_patterns; k++)
        {
            if (pl->pat_lens[k] != P->pattern_lens[k]) return false;
            if (P->pattern_lens[k] && memcmp(pl->patterns[k].data(), P->patterns[k], P->pattern_lens[k]) != 0) return false;
        }
        return true;
    }
    size_t m = algo == KREP_B200_ALGO_MEMCHR ? (P->pattern_len ? 1 : 0) : P->pattern_len;
    if (pl->m != m || memcmp(pl->pattern.data(), P->pattern, m) != 0) return false;
    if ((pl->whole_word != 0) != P->whole_word) return false;
    return pl->built_only_matching == only_matching;
}

static Plan *cached_plan(const search_params_t *P, int algo, bool only_matching)
{
    for (size_t i = 0; i < g_plan_cache.size(); i++)
        if (plan_matches(g_plan_cache[i], P, algo, only_matching))
        {
            Plan *pl = g_plan_cache[i];
            g_plan_cache.erase(g_plan_cache.begin() + i);
            g_plan_cache.push_back(pl); // most recent last
            return pl;
        }
    Plan *pl = plan_build(P, algo, only_matching);
    if (!pl) return nullptr;
    if (g_plan_cache.size() >= 16)
    {
        plan_free(g_plan_cache.front());
        g_plan_cache.erase(g_plan_cache.begin());
    }
    g_plan_cache.push_back(pl);
    return pl;
}

void plan_cache_clear()
{
    for (Plan *p : g_plan_cache) plan_free(p);
    g_plan_cache.clear();
}

// AC-trie handles given to the host (krep_b200_ac_trie_build) own their plan: they are not part of the cache and live
// until krep_b200_ac_trie_free.  The magic word sits first so that a foreign pointer (the reference's own ac_trie_t,
// whose first word is a node pointer) can be told apart by reading 8 bytes.
struct TrieHandle
{
    uint64_t magic;
    Plan *plan;
};
static constexpr uint64_t TRIE_MAGIC = 0x6b7265705f747269ull; // "krep_tri"

// ---- staging: host text -> HBM, overlapped with the scan ----------------------------------------
#define CKH(call)                                                                                  \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
        {                                                                                          \
            set_error(-2, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                      cudaGetErrorString(e_));                                                     \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

static size_t env_mb(const char *name, size_t dflt_mb)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt_mb << 20;
    long x = atol(v);
    return x > 0 ? (size_t)x << 20 : dflt_mb << 20;
}

// The caller's text streams through a small ring of device buffers (3 slots of one chunk + halo each) instead of
// being made resident as a whole: HBM use is bounded whatever the file size, nothing proportional to the text is
// allocated, and a slot is refilled as soon as the scan of its previous occupant has finished.
static int ensure_ring(DevCtx &E, size_t slot_bytes, int slots)
{
    slot_bytes = (slot_bytes + 255) & ~(size_t)255;
    if (E.ring_slot_bytes >= slot_bytes && E.ring_slots >= slots) return 0;
    CKH(cudaDeviceSynchronize());
    cudaFree(E.d_ring);
    E.d_ring = nullptr;
    E.ring_slot_bytes = 0;
    E.ring_slots = 0;
    cudaError_t e = cudaMalloc(&E.d_ring, slot_bytes * slots);
    if (e != cudaSuccess)
    {
        set_error(-2, "cannot allocate %zu bytes of HBM for the staging ring (%s)", slot_bytes * slots, cudaGetErrorString(e));
        return -2;
    }
    E.ring_slot_bytes = slot_bytes;
    E.ring_slots = slots;
    while ((int)E.ring_landed.size() < slots)
    {
        cudaEvent_t a, b;
        CKH(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        E.ring_landed.push_back(a);
        E.ring_scanned.push_back(b);
    }
    return 0;
}

static int ensure_stage(DevCtx &E, size_t bytes, int slots)
{
    if (E.stage_bytes >= bytes && (int)E.stage.size() >= slots) return 0;
    for (auto &s : E.stage)
    {
        cudaFreeHost(s.buf);
        if (s.ev) cudaEventDestroy(s.ev);
    }
    E.stage.clear();
    E.stage.resize(slots);
    for (auto &s : E.stage)
    {
        CKH(cudaMallocHost(&s.buf, bytes));
        CKH(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        s.in_flight = false;
    }
    E.stage_bytes = bytes;
    return 0;
}

static cudaEvent_t pool_event(DevCtx &E, size_t idx)
{
    while (E.ev_pool.size() <= idx)
    {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        E.ev_pool.push_back(ev);
    }
    return E.ev_pool[idx];
}

static int copy_threads()
{
    static int nt = 0;
    if (!nt)
    {
        const char *v = getenv("KREP_B200_COPY_THREADS");
        nt = v ? atoi(v) : 8; // host threads per device (8 already saturate one PCIe link: 53 GB/s measured; 32+ oversubscribe and halve it) that move the caller's (pageable) text into the pinned staging ring
        nt = std::max(1, std::min(nt, std::max(1, omp_get_num_procs())));
    }
    return nt;
}

// Pageable source (krep's file mapping, not pre-populated): populate the page tables of a piece in one batched call
// before copying it, instead of taking one minor fault per 4 KiB page inside memcpy.  MADV_POPULATE_READ (Linux 5.14);
// a kernel without it returns EINVAL and the copy simply faults its way through.
static inline void populate_piece(const uint8_t *src, size_t n)
{
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif
    static const bool on = !getenv("KREP_B200_NO_POPULATE_READ");
    if (!on) return;
    const uintptr_t a = (uintptr_t)src & ~(uintptr_t)4095, e = ((uintptr_t)src + n + 4095) & ~(uintptr_t)4095;
    (void)madvise((void *)a, e - a, MADV_POPULATE_READ);
}

static void parallel_copy(uint8_t *dst, const uint8_t *src, size_t n)
{
    const int nt = copy_threads();
    if (n < (8u << 20) || nt == 1)
    {
        populate_piece(src, n);
        memcpy(dst, src, n);
        return;
    }
    const size_t piece = (n / nt + 4095) & ~(size_t)4095;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int i = 0; i < nt; i++)
    {
        const size_t off = (size_t)i * piece;
        if (off < n)
        {
            populate_piece(src + off, std::min(piece, n - off));
            memcpy(dst + off, src + off, std::min(piece, n - off));
        }
    }
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Called by the warm-up thread once the context exists: the buffers a search on pageable host text (krep's mmap) will
// ask for — device ring, pinned staging ring, occurrence list — are allocated while the host is still opening its file.
void prewarm_host_path(DevCtx &E)
{
    const size_t slot = env_mb("KREP_B200_STAGE_MB", 32) + 4096; // chunk + the longest possible halo (1025) + slack
    if (ensure_ring(E, slot, 3) != 0 || ensure_stage(E, slot, 3) != 0 || ensure_keys(E, 1) != 0) clear_error();
}

// One device's share of a search call: global bytes [begin, end) of the caller's text (end - begin a multiple of the
// chunk size except for the last device), streamed chunk by chunk — host (pinned directly, pageable through the pinned
// staging ring filled by host threads) -> copy stream -> ring slot -> scan stream — every chunk owning the starts in
// its own bytes and reading `halo` bytes of the next chunk (copied with it: the source is host memory, so overlapping
// reads cost nothing) and taking its -w context bytes straight from the host text.  All chunk scans of the device
// append to one occurrence list; k_finish / the radix sort order it.  If the list overflows, it is grown and the
// range is staged again.
struct RangeJob
{
    DevCtx *C = nullptr;
    const Plan *plan = nullptr;
    const char *text = nullptr;
    size_t n = 0, begin = 0, end = 0, chunk = 0;
    int want_positions = 0;
    bool pinned = false;
    bool count_lines = false; // fused -c: every chunk leaves one line record instead of occurrence keys
    std::vector<uint64_t> line_recs; // (lines, flags) per chunk, text order
    // results
    int rc = 0;
    ScanOut so;
    const uint64_t *h_keys = nullptr;
    std::vector<uint64_t> own_keys; // copy of the keys when the device's buffers are reused before the merge
    float kernel_ms = 0.f;
    ErrState err;
};

static int stream_range(RangeJob &J)
{
    DevCtx &E = *J.C;
    CKH(cudaSetDevice(E.device));
    const Plan *plan = J.plan;
    if (!plan_on_device(plan, E)) return -2;
    const uint32_t halo = (plan->is_ac ? plan->max_len : plan->m) + 1; // occurrence + the byte after it (-w)
    const size_t chunk = J.chunk, n = J.n;
    const size_t span = J.end - J.begin;
    const size_t nchunks = (span + chunk - 1) / chunk;
    const int nslots = nchunks >= 3 ? 3 : (int)std::max<size_t>(nchunks, 1);
    if (ensure_ring(E, std::min(chunk, span) + halo + 64, nslots) != 0) return -2;
    if (!J.pinned && ensure_stage(E, std::min(chunk, span) + halo + 64, nslots) != 0) return -2;
    if (J.want_positions && ensure_keys(E, 1) != 0) return -2;
    if (J.count_lines && ensure_line_out(E, nchunks) != 0) return -2;
    reset_kernel_ms();
    const int slot = 0;
    for (int attempt = 0; attempt < 3; attempt++)
    {
        CKH(cudaStreamWaitEvent(E.scan_stream, E.ev_done[slot], 0));
        if (reset_counter(E, slot, E.scan_stream) != 0) return -2;
        for (size_t c = 0; c < nchunks; c++)
        {
            const size_t off = J.begin + c * chunk, len = std::min(chunk, J.end - off);
            const size_t src_len = std::min(len + halo, n - off);
            const int rs = (int)(c % E.ring_slots);
            uint8_t *d_slot = E.d_ring + (size_t)rs * E.ring_slot_bytes;
            if (c >= (size_t)E.ring_slots) CKH(cudaStreamWaitEvent(E.copy_stream, E.ring_scanned[rs], 0));
            if (J.pinned)
                CKH(cudaMemcpyAsync(d_slot, J.text + off, src_len, cudaMemcpyHostToDevice, E.copy_stream));
            else
            {
                StageSlot &s = E.stage[c % E.stage.size()];
                if (s.in_flight) CKH(cudaEventSynchronize(s.ev));
                parallel_copy(s.buf, (const uint8_t *)J.text + off, src_len);
                CKH(cudaMemcpyAsync(d_slot, s.buf, src_len, cudaMemcpyHostToDevice, E.copy_stream));
                CKH(cudaEventRecord(s.ev, E.copy_stream));
                s.in_flight = true;
            }
            CKH(cudaEventRecord(E.ring_landed[rs], E.copy_stream));
            CKH(cudaStreamWaitEvent(E.scan_stream, E.ring_landed[rs], 0));
            krep_b200_shard_t part;
            part.d_text = d_slot;
            part.avail_len = src_len;
            part.own_begin = 0;
            part.own_end = len;
            part.global_offset = off;
            part.prev_byte = off > 0 ? (int32_t)(uint8_t)J.text[off - 1] : -1;
            part.next_byte = off + src_len < n ? (int32_t)(uint8_t)J.text[off + src_len] : -1;
            cudaEvent_t a = pool_event(E, 2 * c), b = pool_event(E, 2 * c + 1);
            CKH(cudaEventRecord(a, E.scan_stream));
            int rc = J.count_lines ? launch_count_lines(E, plan, &part, E.scan_stream, c)
                                   : launch_scan(E, plan, &part, J.want_positions, E.scan_stream, slot);
            if (rc != 0) return rc;
            CKH(cudaEventRecord(b, E.scan_stream));
            CKH(cudaEventRecord(E.ring_scanned[rs], E.scan_stream));
        }
        CKH(cudaGetLastError());
        if (!J.count_lines && finish_scan(E, slot, J.want_positions, E.scan_stream) != 0) return -2;
        CKH(cudaStreamSynchronize(E.scan_stream));
        if (!J.count_lines) CKH(cudaEventSynchronize(E.ev_done[slot])); // k_finish runs on the finish stream
        for (auto &s : E.stage) s.in_flight = false;
        for (size_t c = 0; c < nchunks; c++)
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, E.ev_pool[2 * c], E.ev_pool[2 * c + 1]) == cudaSuccess) add_kernel_ms(ms);
            else cudaGetLastError();
        }
        if (J.count_lines)
        {
            J.line_recs.assign(E.h_line_out, E.h_line_out + 2 * nchunks);
            J.so = ScanOut();
            return 0;
        }
        const uint64_t cnt = E.h_pack[slot][0];
        J.so = ScanOut();
        J.so.count = cnt;
        J.so.device = E.device;
        J.so.serial = ++E.serial;
        E.result_stream = E.scan_stream;
        if (!J.want_positions) return 0;
        if (cnt <= E.key_cap)
        {
            J.so.stored = cnt;
            if (cnt <= PACK_KEYS)
            {
                J.so.d_keys = E.d_pack[slot] + 1;
                J.so.h_sorted = E.h_pack[slot] + 1;
                return 0;
            }
            return sort_keys(E, slot, cnt, key_end_bit(plan, n), E.scan_stream, &J.so.d_keys);
        }
        // list overflowed: grow it and stage the range again (the ring holds only the last chunks)
        J.so.overflow = 1;
        if (ensure_keys(E, cnt + cnt / 8 + 1024) != 0) return -2;
        reset_kernel_ms();
    }
    set_error(-4, "occurrence list kept overflowing");
    return -4;
}

// ---- which devices a host-text call uses --------------------------------------------------------
static std::vector<int> g_devices; // krep_b200_set_devices; empty = automatic

static std::vector<int> host_devices(size_t n)
{
    std::vector<int> out;
    const int vis = visible_devices();
    if (vis == 0) return out;
    if (!g_devices.empty())
    {
        for (int d : g_devices)
            if (d >= 0 && d < vis && std::find(out.begin(), out.end(), d) == out.end()) out.push_back(d);
        if (!out.empty()) return out;
    }
    const int prim = primary_device();
    if (prim < 0) return out;
    // default: ONE device.  Measured (profiles/r2m2_cli_timing.txt, r2m8_cli_timing.txt): from a pageable file mapping a
    // second GPU adds nothing — the host side (page faults + staging copies of one process) is the limit, 30-45 GB/s — and
    // every further context costs 0.15-1.3 s of start-up.  KREP_B200_DEVICES=<k> (or krep_b200_set_devices) spreads the
    // call over k devices, which pays for pinned host text (110 GB/s on 2, 185 GB/s on 8 devices from one NUMA node).
    int want = 1;
    const char *v = getenv("KREP_B200_DEVICES");
    if (v && *v && atoi(v) > 0) want = atoi(v);
    want = std::max(1, std::min(want, vis));
    out.push_back(prim);
    for (int d = 0; d < vis && (int)out.size() < want; d++)
        if (d != prim) out.push_back(d);
    return out;
}

// Result of one host-text search over all devices: total count, and (if wanted) the merged sorted key list on the host.
struct HostScan
{
    uint64_t count = 0;
    const uint64_t *keys = nullptr;
    uint64_t nkeys = 0;
    std::vector<uint64_t> merged; // backing store when several devices contributed
};

static int stage_and_scan(const Plan *plan, const char *text, size_t n, int want_positions, HostScan *hs, bool count_lines = false)
{
    const bool pinned = is_pinned(text);
    const size_t chunk = pinned ? env_mb("KREP_B200_CHUNK_MB", 256) : env_mb("KREP_B200_STAGE_MB", 32);
    std::vector<int> devs = host_devices(n);
    if (devs.empty())
    {
        set_error(-1, "no CUDA device available; this engine has no CPU fallback");
        return -1;
    }
    const size_t nchunks = std::max<size_t>((n + chunk - 1) / chunk, 1);
    if (devs.size() > nchunks) devs.resize(nchunks);
    const size_t D = devs.size();
    // one contiguous range of chunks per device; KREP_B200_RANGES=<k> cuts the text into more ranges than devices
    // (a device then takes its ranges one after the other) — used by the tests to drive the cross-range merge on one GPU
    size_t R = D;
    if (const char *v = getenv("KREP_B200_RANGES"))
        if (atoi(v) > 0) R = std::min<size_t>(std::max<size_t>((size_t)atoi(v), D), nchunks);
    std::vector<RangeJob> jobs(R);
    const size_t per = (nchunks + R - 1) / R; // chunks per range
    for (size_t i = 0; i < R; i++)
    {
        RangeJob &J = jobs[i];
        J.plan = plan;
        J.text = text;
        J.n = n;
        J.chunk = chunk;
        J.begin = std::min(i * per * chunk, n);
        J.end = std::min((i + 1) * per * chunk, n);
        J.want_positions = want_positions;
        J.pinned = pinned;
        J.count_lines = count_lines;
    }
    trace("search: %zu bytes (%s host memory), %zu device(s), %zu range(s), chunk %zu MiB", n, pinned ? "pinned" : "pageable", D, R,
          chunk >> 20);
    // the ranges run on one host thread per device (on the calling thread when there is only one device); every thread
    // brings up its own device's context if it does not exist yet, so several contexts are created side by side
    auto run_device = [&jobs, &devs, D, R](size_t d) {
        for (size_t i = d; i < R; i += D)
        {
            RangeJob &J = jobs[i];
            J.C = ctx_get(devs[d]);
            if (!J.C)
            {
                J.rc = -1;
                break;
            }
            J.rc = J.begin < J.end ? stream_range(J) : 0;
            J.kernel_ms = get_kernel_ms();
            if (J.rc != 0) break;
            if (R > D && J.so.stored) // the device's buffers are reused by its next range: keep this range's keys
            {
                const uint64_t *k = nullptr;
                if ((J.rc = fetch_keys(*J.C, J.so, &k)) != 0) break;
                J.own_keys.assign(k, k + J.so.stored);
                J.h_keys = J.own_keys.data();
            }
        }
    };
    if (D == 1)
    {
        run_device(0);
        float ksum = 0.f;
        for (auto &J : jobs)
        {
            if (J.rc != 0) return J.rc;
            ksum += J.kernel_ms;
        }
        set_kernel_ms(ksum);
    }
    else
    {
        std::vector<std::thread> th;
        for (size_t d = 0; d < D; d++)
            th.emplace_back([&jobs, &run_device, d, D, R] {
                clear_error();
                run_device(d);
                for (size_t i = d; i < R; i += D) get_error(&jobs[i].err);
            });
        for (auto &t : th) t.join();
        std::vector<float> kdev(D, 0.f);
        for (size_t i = 0; i < R; i++)
        {
            if (jobs[i].rc != 0)
            {
                adopt_error(jobs[i].err);
                return jobs[i].rc;
            }
            kdev[i % D] += jobs[i].kernel_ms;
        }
        set_kernel_ms(*std::max_element(kdev.begin(), kdev.end())); // devices scan concurrently: the slowest one's time
    }
    trace("search: all ranges scanned (slowest device: %.2f ms of scan kernels)", get_kernel_ms());
    hs->count = 0;
    hs->keys = nullptr;
    hs->nkeys = 0;
    if (count_lines)
    {
        // chunk records of all ranges, in text order -> matching lines (a line cut by a chunk / range / device edge is
        // counted on both sides and subtracted once)
        std::vector<uint64_t> all;
        for (auto &J : jobs) all.insert(all.end(), J.line_recs.begin(), J.line_recs.end());
        hs->count = combine_line_records(all.data(), all.size() / 2);
        return 0;
    }
    for (auto &J : jobs) hs->count += J.so.count;
    if (!want_positions) return 0;
    for (auto &J : jobs)
    {
        if (J.so.stored == 0 || J.h_keys || !J.C) continue;
        cudaSetDevice(J.C->device);
        if (fetch_keys(*J.C, J.so, &J.h_keys) != 0) return -2;
    }
    if (R == 1)
    {
        hs->keys = jobs[0].h_keys;
        hs->nkeys = jobs[0].so.stored;
        return 0;
    }
    // ranges are in text order: literal keys concatenate in order, pattern-set keys (ordered by end, owned by start) need
    // the merge around each cut
    std::vector<const uint64_t *> lists;
    std::vector<uint64_t> counts;
    uint64_t total = 0;
    for (auto &J : jobs)
    {
        lists.push_back(J.h_keys);
        counts.push_back(J.so.stored);
        total += J.so.stored;
    }
    hs->merged.resize(total);
    hs->nkeys = merge_key_lists(lists.data(), counts.data(), (uint32_t)lists.size(), hs->merged.data());
    hs->keys = hs->merged.data();
    trace("search: merged %llu keys from %zu ranges", (unsigned long long)hs->nkeys, R);
    return 0;
}

// Does the emulated kernel keep every (whole-word-valid) occurrence?  If so a bare count is enough.
static bool keeps_all(int algo, bool only_matching, const search_params_t *P, const Plan *pl)
{
    if (pl->is_ac) return true;
    if (pl->emit_len != pl->m) return false;
    // window kernels: tail sub-search, AVX-512's skipped windows and the -m re-basing all need the list
    if (algo == KREP_B200_ALGO_AVX2 || algo == KREP_B200_ALGO_AVX512 || algo == KREP_B200_ALGO_NEON) return false;
    if (pl->border_free) return true; // occurrences cannot overlap: every overlap policy keeps all
    switch (algo)
    {
    case KREP_B200_ALGO_BMH: return !(only_matching && !P->count_lines_mode);
    case KREP_B200_ALGO_MEMCHR: return true;
    case KREP_B200_ALGO_MEMCHR_SHORT: return !only_matching;
    case KREP_B200_ALGO_SSE42: return only_matching;
    default: return false;
    }
}

// Return value of the emulated kernel when all `total` occurrences are kept and only the -m limit acts.
static uint64_t limited_count(int algo, const search_params_t *P, uint64_t total)
{
    const uint64_t maxc = P->max_count;
    switch (algo)
    {
    case KREP_B200_ALGO_BMH:
    case KREP_B200_ALGO_MEMCHR_SHORT:
        // count first, test afterwards (krep.c:1355-1367): with -m 0 in a pure count mode one match still counts
        if (total == 0) return 0;
        return maxc == 0 ? 1 : std::min<uint64_t>(total, maxc);
    default:
        return std::min<uint64_t>(total, maxc);
    }
}

// The early returns every reference kernel takes before it looks at the text.  Returns true when the call is already
// answered (*ret); otherwise *algo is the kernel whose policy applies (precondition fallbacks resolved).
static bool early_answer(int entry_algo, const search_params_t *P, const char *text, size_t n, match_result_t *res, int *algo_out,
                         uint64_t *ret)
{
    *ret = 0;
    int algo = entry_algo;
    size_t m = 0;
    if (algo == KREP_B200_ALGO_AC)
    {
        if (!P->ac_trie || !text) return true;  // aho_corasick.c:306
        if (P->max_count == 0) return true;     // aho_corasick.c:316
        if (n == 0)                             // aho_corasick.c:442-463
        {
            for (size_t k = 0; k < P->num_patterns; k++)
                if (P->pattern_lens[k] == 0)
                {
                    if (P->track_positions && res) result_push(res, 0, 0);
                    *ret = 1;
                    return true;
                }
            return true;
        }
    }
    else
    {
        algo = resolve_algo(P, algo);
        m = P->pattern_len;
        switch (algo)
        {
        case KREP_B200_ALGO_KMP:
            if (P->max_count == 0) return true;
            if (m == 0 || n < m) return true;
            break;
        case KREP_B200_ALGO_MEMCHR:
            if (P->max_count == 0 || n == 0) return true;
            m = 1;
            break;
        case KREP_B200_ALGO_MEMCHR_SHORT:
            if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return true;
            if (m < 2 || m > 3 || n < m) return true;
            break;
        default: // BMH, SSE42, the long simd entries, NEON
            if (P->max_count == 0 && (P->count_lines_mode || P->track_positions)) return true;
            if (m == 0 || n < m) return true;
            break;
        }
        if (!P->pattern) return true;
        if (m > 1024)
        {
            set_error(-3, "pattern longer than 1024 bytes (MAX_PATTERN_LENGTH, krep.c:77)");
            return true;
        }
    }
    *algo_out = algo;
    return false;
}

// The plan a call runs: the host's own trie handle when it was built by krep_b200_ac_trie_build for these very
// patterns, else the cache.
static Plan *plan_for(const search_params_t *P, int algo, bool only_matching)
{
    if (algo == KREP_B200_ALGO_AC && P->ac_trie)
    {
        const TrieHandle *h = reinterpret_cast<const TrieHandle *>(P->ac_trie);
        if (h->magic == TRIE_MAGIC && h->plan && plan_matches(h->plan, P, algo, only_matching)) return h->plan;
    }
    return cached_plan(P, algo, only_matching);
}

static uint64_t run_search(int entry_algo, const search_params_t *P, const char *text, size_t n, match_result_t *res)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!P) return 0;
    const bool only_matching = g_only_matching;
    int algo = entry_algo;
    uint64_t early = 0;
    if (early_answer(entry_algo, P, text, n, res, &algo, &early)) return early;
    if (warm_running())
    {
        // the context is still being created on the warm-up thread: meanwhile fault the caller's pages in (a file the
        // host mapped without MAP_POPULATE) with the staging threads, so that the copy loop later runs at link speed
        // (opt-in: measured on the bench box, faulting pages in the same process while cuInit / context creation run
        // more than doubles their time — both sides fight over the address-space lock; profiles/r2b_cli_timing.txt)
        if (n >= (64u << 20) && getenv("KREP_B200_PREFAULT"))
        {
            trace("search: pre-faulting %zu bytes while the context comes up", n);
            const long pages = (long)((n + 4095) / 4096);
            unsigned long sink = 0;
#pragma omp parallel for num_threads(copy_threads()) schedule(static) reduction(+ : sink)
            for (long pg = 0; pg < pages; pg++) sink += (unsigned char)text[(size_t)pg * 4096];
            if (sink == 0x5EED5EED5EEDull) trace("(unlikely checksum)");
        }
        warm_join();
        trace("search: context ready");
    }
    if (visible_devices() == 0)
    {
        set_error(-1, "no CUDA device available; this engine has no CPU fallback");
        return 0;
    }
    DeviceGuard guard;
    Plan *plan = plan_for(P, algo, only_matching);
    if (!plan) return 0;
    const bool want_result = P->track_positions && res;
    const bool need_list = P->count_lines_mode || want_result || !keeps_all(algo, only_matching, P, plan) ||
                           plan->whole_word == 2;
    HostScan hs;
    if (count_lines_eligible(plan, P, algo) && !getenv("KREP_B200_NO_FUSED_COUNT"))
    {
        // -c: the scan counts matching lines itself; the -m limit caps the count (every kernel stops at max_count lines,
        // max_count == 0 was answered above)
        if (stage_and_scan(plan, text, n, 0, &hs, true) != 0) return 0;
        return std::min<uint64_t>(hs.count, P->max_count);
    }
    if (stage_and_scan(plan, text, n, need_list ? 1 : 0, &hs) != 0) return 0;
    if (!need_list) return limited_count(algo, P, hs.count);
    Replay r{hs.keys, (size_t)hs.nkeys, text, n, 0};
    const uint64_t ret = plan->is_ac ? replay_ac(P, r, res) : replay_literal(algo, P, only_matching, plan->m, r, res);
    trace("search: replay done (%llu)", (unsigned long long)ret);
    return ret;
}

// Many texts, one launch (small files lose to launch and copy latency one by one).  The texts are packed
// into one pinned buffer at 16-byte aligned offsets, separated by zero gaps longer than the longest pattern, copied and
// scanned as ONE shard; the sorted occurrence list is then cut per text — an occurrence belongs to a text only if it
// lies wholly inside it — and each cut is replayed exactly as a separate call would have been (same early returns, own
// -m limit, own line context).  A gap byte is 0, i.e. not a word character: -w sees a text boundary there, as it should.
static int run_batch(int entry_algo, const search_params_t *P, const char *const *texts, const size_t *lens, size_t nt,
                     uint64_t *counts, match_result_t *const *results)
{
    warm_join();
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!P || !texts || !lens || !counts)
    {
        set_error(-3, "krep_b200_search_batch: null argument");
        return -3;
    }
    const bool only_matching = g_only_matching;
    std::vector<int> algo_of(nt, -1);
    int algo = -1;
    for (size_t f = 0; f < nt; f++)
    {
        int a = entry_algo;
        uint64_t early = 0;
        counts[f] = 0;
        if (early_answer(entry_algo, P, texts[f], lens[f], results ? results[f] : nullptr, &a, &early)) counts[f] = early;
        else algo_of[f] = algo = a; // the resolved kernel depends on params only, except for the n < m early return above
    }
    if (krep_b200_last_error() != 0) return -3;
    if (algo < 0) return 0; // every text was answered by an early return
    DeviceGuard guard;
    DevCtx *Cp = ctx_primary();
    if (!Cp) return -1;
    Plan *plan = plan_for(P, algo, only_matching);
    if (!plan) return -2;
    const size_t gap = (size_t)(plan->is_ac ? plan->max_len : plan->m) + 16;
    std::vector<uint64_t> off(nt, 0);
    uint64_t total = 0;
    for (size_t f = 0; f < nt; f++)
        if (algo_of[f] >= 0)
        {
            off[f] = total;
            total = (total + lens[f] + gap + 15) & ~15ull;
        }
    DevCtx &E = *Cp;
    if (total > E.h_batch_cap)
    {
        cudaFreeHost(E.h_batch);
        E.h_batch = nullptr;
        E.h_batch_cap = 0;
        CKH(cudaMallocHost(&E.h_batch, total + total / 4 + 4096));
        E.h_batch_cap = total + total / 4 + 4096;
    }
    {
        // pack with the staging threads; only the gaps are zeroed
        std::vector<size_t> live;
        for (size_t f = 0; f < nt; f++)
            if (algo_of[f] >= 0) live.push_back(f);
        const long nl = (long)live.size();
        uint8_t *const hb = E.h_batch;
#pragma omp parallel for num_threads(copy_threads()) schedule(dynamic, 16)
        for (long i = 0; i < nl; i++)
        {
            const size_t f = live[(size_t)i];
            const uint64_t end = off[f] + lens[f], next = i + 1 < nl ? off[live[(size_t)i + 1]] : total;
            memcpy(hb + off[f], texts[f], lens[f]);
            memset(hb + end, 0, next - end);
        }
    }
    HostScan hs;
    if (stage_and_scan(plan, (const char *)E.h_batch, total, 1, &hs) != 0) return -2;
    const uint64_t *keys = hs.keys;
    struct { uint64_t stored; } so{hs.nkeys};
    // cut the list per text (texts are in ascending offset order; keys ascend by start, or by end for pattern sets)
    std::vector<uint64_t> mine;
    size_t j = 0;
    for (size_t f = 0; f < nt; f++)
    {
        if (algo_of[f] < 0) continue;
        const uint64_t lo = off[f], hi = off[f] + lens[f];
        mine.clear();
        auto span = [&](uint64_t key, uint64_t *s, uint64_t *e) {
            if (plan->is_ac)
            {
                *e = key >> AC_END_SHIFT;
                *s = *e - (1024 - ((key >> AC_LEN_SHIFT) & 1023));
            }
            else
            {
                *s = key >> LIT_TAG_BITS;
                *e = *s + ((key >> 2) & 1 ? plan->m : plan->emit_len);
            }
        };
        while (j < so.stored)
        {
            uint64_t s, e;
            span(keys[j], &s, &e);
            const uint64_t ord = plan->is_ac ? e : s; // the coordinate the list is sorted by
            if (ord >= hi + (plan->is_ac ? gap : 0)) break; // belongs to a later text
            if (s >= lo && e <= hi) mine.push_back(keys[j]);
            j++;
        }
        Replay r{mine.data(), mine.size(), texts[f], lens[f], lo};
        match_result_t *res = results ? results[f] : nullptr;
        counts[f] = plan->is_ac ? replay_ac(P, r, res) : replay_literal(algo, P, only_matching, plan->m, r, res);
    }
    return 0;
}

} // namespace kb

using namespace kb;

extern "C" {

void krep_b200_set_only_matching(bool on) { g_only_matching = on; }
bool krep_b200_get_only_matching(void) { return g_only_matching; }
void krep_b200_set_force_no_simd(bool on) { g_force_no_simd = on; }
void krep_b200_set_algo_override(const char *name) { g_algo_override = name ? name : ""; }

uint64_t krep_b200_boyer_moore_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_BMH, p, t, n, r);
}
uint64_t krep_b200_kmp_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_KMP, p, t, n, r);
}
uint64_t krep_b200_memchr_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_MEMCHR, p, t, n, r);
}
uint64_t krep_b200_memchr_short_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_MEMCHR_SHORT, p, t, n, r);
}
uint64_t krep_b200_simd_sse42_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_SSE42, p, t, n, r);
}
uint64_t krep_b200_simd_avx2_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_AVX2, p, t, n, r);
}
uint64_t krep_b200_simd_avx512_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_AVX512, p, t, n, r);
}
uint64_t krep_b200_aho_corasick_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_AC, p, t, n, r);
}
uint64_t krep_b200_neon_search(const search_params_t *p, const char *t, size_t n, match_result_t *r)
{
    return run_search(KREP_B200_ALGO_NEON, p, t, n, r);
}

int krep_b200_search_batch(search_func_t entry, const search_params_t *params, const char *const *texts, const size_t *lens,
                           size_t n_texts, uint64_t *counts, match_result_t *const *results)
{
    int algo = -1;
    if (entry == krep_b200_boyer_moore_search) algo = KREP_B200_ALGO_BMH;
    else if (entry == krep_b200_kmp_search) algo = KREP_B200_ALGO_KMP;
    else if (entry == krep_b200_memchr_search) algo = KREP_B200_ALGO_MEMCHR;
    else if (entry == krep_b200_memchr_short_search) algo = KREP_B200_ALGO_MEMCHR_SHORT;
    else if (entry == krep_b200_simd_sse42_search) algo = KREP_B200_ALGO_SSE42;
    else if (entry == krep_b200_simd_avx2_search) algo = KREP_B200_ALGO_AVX2;
    else if (entry == krep_b200_simd_avx512_search) algo = KREP_B200_ALGO_AVX512;
    else if (entry == krep_b200_aho_corasick_search) algo = KREP_B200_ALGO_AC;
    else if (entry == krep_b200_neon_search) algo = KREP_B200_ALGO_NEON;
    if (algo < 0)
    {
        set_error(-3, "krep_b200_search_batch: entry must be one of this library's search_func_t entry points");
        return -3;
    }
    return run_batch(algo, params, texts, lens, n_texts, counts, results);
}

// krep.c:1873-1914
static bool is_repetitive_pattern(const char *pattern, size_t len)
{
    if (len < 3) return false;
    size_t run = 0;
    char prev = pattern[0];
    for (size_t i = 1; i < len; i++)
    {
        if (pattern[i] == prev)
        {
            if (++run >= len / 2) return true;
        }
        else
        {
            run = 0;
            prev = pattern[i];
        }
    }
    for (size_t period = 2; period <= len / 2; period++)
    {
        bool periodic = true;
        for (size_t i = period; i < len && periodic; i++) periodic = pattern[i] == pattern[i % period];
        if (periodic) return true;
    }
    return false;
}

// krep.c:1771-1870, for the AVX2 build of the reference (SIMD_MAX_PATTERN_LEN 32).
search_func_t krep_b200_select_search_algorithm(const search_params_t *P)
{
    if (!P || P->use_regex) return NULL; // regex stays with the host's regex_search
    if (P->num_patterns > 1) return krep_b200_aho_corasick_search;
    if (!g_algo_override.empty() && g_algo_override != "auto")
    {
        if (g_algo_override == "bm") return krep_b200_boyer_moore_search;
        if (g_algo_override == "kmp") return krep_b200_kmp_search;
    }
    const size_t m = P->pattern_len;
    const bool can_simd = !g_force_no_simd && m <= SIMD_MAX_PATTERN_LEN;
    if (m == 1) return krep_b200_memchr_search;
    if (m < 4) return (can_simd && P->case_sensitive) ? krep_b200_simd_avx2_search : krep_b200_memchr_short_search;
    if (can_simd && m <= 32) return krep_b200_simd_avx2_search;
    if (m < 8 && is_repetitive_pattern(P->pattern, m)) return krep_b200_kmp_search;
    return krep_b200_boyer_moore_search;
}

const char *krep_b200_get_algorithm_name(search_func_t f)
{
    if (f == krep_b200_boyer_moore_search) return "Boyer-Moore-Horspool";
    if (f == krep_b200_kmp_search) return "Knuth-Morris-Pratt";
    if (f == krep_b200_aho_corasick_search) return "Aho-Corasick";
    if (f == krep_b200_memchr_search) return "memchr";
    if (f == krep_b200_memchr_short_search) return "memchr-short";
    if (f == krep_b200_simd_sse42_search) return "SSE4.2";
    if (f == krep_b200_simd_avx2_search) return "AVX2";
    if (f == krep_b200_simd_avx512_search) return "AVX-512";
    if (f == krep_b200_neon_search) return "NEON";
    return "Unknown";
}

// ---- AC trie handles (aho_corasick.c:111 / 274 / 287) ----
ac_trie_t *krep_b200_ac_trie_build(const search_params_t *params)
{
    warm_join();
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!params || params->num_patterns == 0) return NULL; // aho_corasick.c:113
    if (visible_devices() == 0)
    {
        set_error(-1, "no CUDA device available; this engine has no CPU fallback");
        return NULL;
    }
    Plan *pl = plan_build(params, KREP_B200_ALGO_AC, false); // owned by the handle, not by the plan cache
    if (!pl) return NULL;
    TrieHandle *h = new TrieHandle{TRIE_MAGIC, pl};
    return reinterpret_cast<ac_trie_t *>(h);
}
void krep_b200_ac_trie_free(ac_trie_t *trie)
{
    TrieHandle *h = reinterpret_cast<TrieHandle *>(trie);
    if (!h || h->magic != TRIE_MAGIC) return;
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    plan_free(h->plan);
    h->magic = 0;
    delete h;
}
bool krep_b200_ac_trie_root_has_outputs(const ac_trie_t *trie)
{
    const TrieHandle *h = reinterpret_cast<const TrieHandle *>(trie);
    if (!h || h->magic != TRIE_MAGIC || !h->plan) return false;
    for (uint32_t len : h->plan->pat_lens)
        if (len == 0) return true; // an empty pattern's index sits on the root (aho_corasick.c:145)
    return false;
}

void krep_b200_set_devices(const int *devices, int n)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    if (devices && n > 0) keep_devices_visible();
    g_devices.clear();
    for (int i = 0; devices && i < n; i++) g_devices.push_back(devices[i]);
}

// ---- match_result helpers (krep.c:139 / 175 / 244 / 256) ----
match_result_t *krep_b200_match_result_init(uint64_t initial_capacity)
{
    match_result_t *r = (match_result_t *)malloc(sizeof *r);
    if (!r) return NULL;
    if (initial_capacity == 0) initial_capacity = 16;
    if (initial_capacity > SIZE_MAX / sizeof(match_position_t))
    {
        free(r);
        return NULL;
    }
    r->positions = (match_position_t *)malloc(initial_capacity * sizeof(match_position_t));
    if (!r->positions)
    {
        free(r);
        return NULL;
    }
    r->count = 0;
    r->capacity = initial_capacity;
    return r;
}
bool krep_b200_match_result_add(match_result_t *r, size_t s, size_t e) { return result_push(r, s, e); }
void krep_b200_match_result_free(match_result_t *r)
{
    if (!r) return;
    free(r->positions);
    free(r);
}
bool krep_b200_match_result_merge(match_result_t *dest, const match_result_t *src, size_t chunk_offset)
{
    if (!dest || !src || src->count == 0) return true;
    for (uint64_t i = 0; i < src->count; i++)
        if (!result_push(dest, src->positions[i].start_offset + chunk_offset, src->positions[i].end_offset + chunk_offset))
            return false;
    return true;
}

uint64_t krep_b200_replay(int algo, const search_params_t *P, bool only_matching, const uint64_t *keys, uint64_t nkeys,
                          const char *text, size_t text_len, match_result_t *result)
{
    if (!P) return 0;
    if (P->count_lines_mode && !text && nkeys)
    {
        set_error(-3, "krep_b200_replay: -c line counting needs the host text");
        return 0;
    }
    Replay r{keys, (size_t)nkeys, text, text_len ? text_len : (SIZE_MAX >> 1), 0};
    if (algo == KREP_B200_ALGO_AC) return replay_ac(P, r, result);
    algo = resolve_algo(P, algo);
    const uint32_t m = algo == KREP_B200_ALGO_MEMCHR ? 1u : (uint32_t)P->pattern_len;
    return replay_literal(algo, P, only_matching, m, r, result);
}

uint64_t krep_b200_replay_lines(int algo, const search_params_t *P, bool only_matching, const uint64_t *keys, uint64_t nkeys,
                                const uint64_t *bounds, size_t text_len, match_result_t *result)
{
    if (!P) return 0;
    if (P->count_lines_mode && !bounds && nkeys)
    {
        set_error(-3, "krep_b200_replay_lines: -c needs the line bounds");
        return 0;
    }
    Replay r{keys, (size_t)nkeys, nullptr, text_len ? text_len : (SIZE_MAX >> 1), 0, bounds};
    if (algo == KREP_B200_ALGO_AC) return replay_ac(P, r, result);
    algo = resolve_algo(P, algo);
    const uint32_t m = algo == KREP_B200_ALGO_MEMCHR ? 1u : (uint32_t)P->pattern_len;
    return replay_literal(algo, P, only_matching, m, r, result);
}

// ---- fused -c on resident shards ----
int krep_b200_count_lines_shard(const krep_b200_plan_t *plan_, const search_params_t *P, const krep_b200_shard_t *shard, void *stream,
                                krep_b200_line_count_t *out)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    const Plan *plan = reinterpret_cast<const Plan *>(plan_);
    if (!plan || !P || !shard || !out)
    {
        set_error(-3, "krep_b200_count_lines_shard: null argument");
        return -3;
    }
    if (!count_lines_eligible(plan, P, plan->algo))
    {
        set_error(-3, "krep_b200_count_lines_shard: this plan's -c result needs the occurrence list (pattern set, window kernel, "
                      "newline in the pattern or tag-mode -w): use krep_b200_scan_shard + krep_b200_collect");
        return -3;
    }
    DeviceGuard guard;
    cudaPointerAttributes a;
    DevCtx *C = (cudaPointerGetAttributes(&a, shard->d_text) == cudaSuccess && a.type == cudaMemoryTypeDevice) ? ctx_get(a.device) : ctx_primary();
    if (!C) return -1;
    cudaStream_t st = stream ? (cudaStream_t)stream : C->scan_stream;
    reset_kernel_ms();
    if (cudaEventRecord(C->ev_ca, st) != cudaSuccess) return -2;
    int rc = launch_count_lines(*C, plan, shard, st, 0);
    if (rc != 0) return rc;
    if (cudaEventRecord(C->ev_cb, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    {
        set_error(-2, "CUDA error in the fused line count (%s)", cudaGetErrorString(cudaGetLastError()));
        return -2;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, C->ev_ca, C->ev_cb);
    add_kernel_ms(ms);
    out->lines = C->h_line_out[0];
    out->flags = (uint32_t)C->h_line_out[1];
    out->reserved = 0;
    return 0;
}

uint64_t krep_b200_combine_line_counts(const krep_b200_line_count_t *recs, size_t n, size_t max_count)
{
    if (!recs) return 0;
    std::vector<uint64_t> flat(2 * n);
    for (size_t i = 0; i < n; i++)
    {
        flat[2 * i] = recs[i].lines;
        flat[2 * i + 1] = recs[i].flags;
    }
    return std::min<uint64_t>(combine_line_records(flat.data(), n), max_count);
}

// ---- several resident shards, one answer: search_file's chunk loop + merge (krep.c:2851-3004) for text that already
// lives in HBM, possibly on several GPUs of this process.  Scans run concurrently on distinct devices; per-shard lists are
// merged by key; the emulated kernel's policy is replayed once over the whole list (so -m, overlap rules and the
// emission order are global, not per shard).
uint64_t krep_b200_search_shards(const krep_b200_plan_t *plan_, const search_params_t *P, const krep_b200_shard_t *shards,
                                 uint32_t n_shards, match_result_t *result)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    const Plan *plan = reinterpret_cast<const Plan *>(plan_);
    if (!plan || !P || (!shards && n_shards))
    {
        set_error(-3, "krep_b200_search_shards: null argument");
        return 0;
    }
    DeviceGuard guard;
    std::vector<DevCtx *> ctx(n_shards, nullptr);
    uint64_t text_len = 0;
    for (uint32_t i = 0; i < n_shards; i++)
    {
        cudaPointerAttributes a;
        ctx[i] = (cudaPointerGetAttributes(&a, shards[i].d_text) == cudaSuccess && a.type == cudaMemoryTypeDevice) ? ctx_get(a.device)
                                                                                                                    : ctx_primary();
        if (!ctx[i]) return 0;
        text_len = std::max<uint64_t>(text_len, shards[i].global_offset + shards[i].avail_len);
    }
    reset_kernel_ms();
    if (P->count_lines_mode)
    {
        if (!count_lines_eligible(plan, P, plan->algo))
        {
            set_error(-3, "krep_b200_search_shards: -c over several shards is only available where the scan counts lines itself "
                          "(single literals; see krep_b200_count_lines_shard)");
            return 0;
        }
        if (P->max_count == 0) return 0;
        std::vector<uint64_t> recs(2 * (size_t)n_shards);
        for (uint32_t i = 0; i < n_shards; i++) // size every device's record array first: growing it later would move it
        {
            cudaSetDevice(ctx[i]->device);
            if (ensure_line_out(*ctx[i], n_shards) != 0) return 0;
        }
        for (uint32_t i = 0; i < n_shards; i++)
        {
            cudaSetDevice(ctx[i]->device);
            if (launch_count_lines(*ctx[i], plan, &shards[i], ctx[i]->scan_stream, i) != 0) return 0;
        }
        for (uint32_t i = 0; i < n_shards; i++)
        {
            cudaSetDevice(ctx[i]->device);
            if (cudaStreamSynchronize(ctx[i]->scan_stream) != cudaSuccess)
            {
                set_error(-2, "CUDA error in the fused line count (%s)", cudaGetErrorString(cudaGetLastError()));
                return 0;
            }
            recs[2 * i] = ctx[i]->h_line_out[2 * i];
            recs[2 * i + 1] = ctx[i]->h_line_out[2 * i + 1];
        }
        return std::min<uint64_t>(combine_line_records(recs.data(), n_shards), P->max_count);
    }
    const bool need_list = (P->track_positions && result) || !keeps_all(plan->algo, plan->built_only_matching, P, plan) || plan->whole_word == 2;
    std::vector<std::vector<uint64_t>> keys(n_shards);
    std::vector<int> pending(MAX_DEV, -1); // shard whose scan is in flight on each device
    std::vector<int> slot_of(n_shards, 0);
    uint64_t total_count = 0;
    float kmax = 0.f;
    auto finish = [&](int i) -> int {
        DevCtx &C = *ctx[i];
        cudaSetDevice(C.device);
        ScanOut so;
        int rc = scan_end(C, slot_of[i], &so);
        kmax = std::max(kmax, get_kernel_ms());
        if (rc != 0) return rc;
        total_count += so.count;
        if (need_list && so.stored)
        {
            const uint64_t *k = nullptr;
            if ((rc = fetch_keys(C, so, &k)) != 0) return rc;
            keys[i].assign(k, k + so.stored);
        }
        return 0;
    };
    for (uint32_t i = 0; i < n_shards; i++)
    {
        DevCtx &C = *ctx[i];
        if (pending[C.device] >= 0)
        {
            if (finish(pending[C.device]) != 0) return 0;
            pending[C.device] = -1;
        }
        cudaSetDevice(C.device);
        if (scan_begin(C, plan, &shards[i], need_list ? 1 : 0, nullptr, &slot_of[i]) != 0) return 0;
        pending[C.device] = (int)i;
    }
    for (int d = 0; d < MAX_DEV; d++)
        if (pending[d] >= 0 && finish(pending[d]) != 0) return 0;
    set_kernel_ms(kmax);
    if (!need_list) return limited_count(plan->algo, P, total_count);
    std::vector<const uint64_t *> lists;
    std::vector<uint64_t> counts;
    uint64_t total = 0;
    for (auto &k : keys)
    {
        lists.push_back(k.data());
        counts.push_back(k.size());
        total += k.size();
    }
    std::vector<uint64_t> merged(total ? total : 1);
    const uint64_t nk = merge_key_lists(lists.data(), counts.data(), (uint32_t)lists.size(), merged.data());
    Replay r{merged.data(), (size_t)nk, nullptr, text_len ? (size_t)text_len : (SIZE_MAX >> 1), 0};
    if (plan->is_ac) return replay_ac(P, r, result);
    return replay_literal(plan->algo, P, plan->built_only_matching, plan->m, r, result);
}

// ---- shard result -> match_result_t under the emulated kernel's policy ----
uint64_t krep_b200_collect(const krep_b200_plan_t *plan_, const search_params_t *P, const krep_b200_device_result_t *dev,
                           match_result_t *result)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    const Plan *plan = reinterpret_cast<const Plan *>(plan_);
    if (!plan || !P || !dev) return 0;
    if (P->count_lines_mode && dev->stored && !dev->d_line_bounds)
    {
        set_error(-3, "krep_b200_collect: -c needs a plan created with count_lines_mode (line bounds are computed by the scan)");
        return 0;
    }
    if (!dev->stored) return limited_count(plan->algo, P, plan->is_ac || keeps_all(plan->algo, g_only_matching, P, plan) ? dev->count : 0);
    DeviceGuard guard;
    DevCtx *Cp = ctx_get(dev->device);
    if (!Cp) return 0;
    DevCtx &E = *Cp;
    ScanOut so;
    so.count = dev->count;
    so.stored = dev->stored;
    so.d_keys = dev->d_keys;
    // the keys came back with the count if the list was short and no later scan has reused the slot
    if (dev->serial == E.serial && dev->stored <= PACK_KEYS && dev->slot >= 0 && dev->slot < SCAN_SLOTS)
        so.h_sorted = E.h_pack[dev->slot] + 1;
    const uint64_t *keys = nullptr;
    if (fetch_keys(E, so, &keys) != 0) return 0;
    Replay r{keys, (size_t)so.stored, nullptr, dev->text_len ? (size_t)dev->text_len : (SIZE_MAX >> 1), 0};
    if (P->count_lines_mode)
    {
        // read the device-computed line bounds back and resolve the "same line as my neighbour" markers
        const uint64_t nb = 2 * so.stored;
        if (nb > E.h_bounds_cap)
        {
            cudaFreeHost(E.h_bounds);
            E.h_bounds = nullptr;
            E.h_bounds_cap = 0;
            if (cudaMallocHost(&E.h_bounds, (nb + nb / 4 + 1024) * sizeof(uint64_t)) != cudaSuccess)
            {
                set_error(-2, "cannot allocate pinned memory for line bounds");
                return 0;
            }
            E.h_bounds_cap = nb + nb / 4 + 1024;
        }
        cudaStream_t st = E.result_stream ? E.result_stream : E.scan_stream; // ordered after the sort and k_line_bounds
        if (cudaMemcpyAsync(E.h_bounds, dev->d_line_bounds, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
        {
            set_error(-2, "reading line bounds back failed");
            return 0;
        }
        uint64_t *b = E.h_bounds;
        for (uint64_t i = 0; i < so.stored; i++)
            if (b[2 * i] == LB_SAME_AS_PREV) b[2 * i] = i ? b[2 * (i - 1)] : LB_OUTSIDE_SHARD;
        for (uint64_t i = so.stored; i-- > 0;)
            if (b[2 * i + 1] == LB_SAME_AS_NEXT) b[2 * i + 1] = i + 1 < so.stored ? b[2 * (i + 1) + 1] : LB_OUTSIDE_SHARD;
        for (uint64_t i = 0; i < nb; i++)
            if (b[i] == LB_OUTSIDE_SHARD)
            {
                set_error(-3, "krep_b200_collect: a matching line continues into a neighbouring shard; -c needs newline-aligned shards");
                return 0;
            }
        r.bounds = b;
    }
    if (plan->is_ac) return replay_ac(P, r, result);
    return replay_literal(plan->algo, P, plan->built_only_matching, plan->m, r, result);
}

} // extern "C"
