// engine.cu — process-wide engine context, plan compilation, shard scan (launch + device sort),
// synthetic corpus generator, and the device-level half of the C ABI (include/krep_b200.h).
#include <cub/device/device_radix_sort.cuh>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.h"
#include "corpus.h"
#include "engine.h"

namespace kb {

// ---------------------------------------------------------------------------------------------
// errors (reference convention: "krep: ..." on stderr, no in-band channel — krep.c:1933)
// ---------------------------------------------------------------------------------------------
static thread_local int t_err = 0;
static thread_local char t_errmsg[512] = "";

void set_error(int code, const char *fmt, ...)
{
    t_err = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_errmsg, sizeof t_errmsg, fmt, ap);
    va_end(ap);
    fprintf(stderr, "krep: %s\n", t_errmsg);
}
void clear_error()
{
    t_err = 0;
    t_errmsg[0] = 0;
}

#define CK(call)                                                                                   \
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

static Engine g_engine;
static std::recursive_mutex g_mu;
static uint64_t g_launches = 0;
static thread_local float t_kernel_ms = 0.f;

Engine &engine() { return g_engine; }
std::recursive_mutex &engine_mutex() { return g_mu; }
void count_launch(int n) { g_launches += (uint64_t)n; }
void add_kernel_ms(float ms) { t_kernel_ms += ms; }
void reset_kernel_ms() { t_kernel_ms = 0.f; }

int engine_init(int device)
{
    Engine &E = g_engine;
    if (E.ready) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        set_error(-1, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
        return -1;
    }
    if (device < 0)
    {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        set_error(-1, "device %d (%s, sm_%d%d) is not an sm_100 part; kernels are built for sm_100a only", device,
                  prop.name, prop.major, prop.minor);
        return -1;
    }
    E.device = device;
    E.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&E.scan_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&E.copy_stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&E.d_counter, 64));
    CK(cudaMallocHost(&E.h_counter, 64));
    CK(cudaEventCreate(&E.ev_a));
    CK(cudaEventCreate(&E.ev_b));
    E.ready = true;
    return 0;
}

bool engine_ok()
{
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (g_engine.ready) return true;
    return engine_init(-1) == 0;
}

void engine_shutdown()
{
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    Engine &E = g_engine;
    if (!E.ready) return;
    cudaSetDevice(E.device);
    cudaDeviceSynchronize();
    for (auto *p : E.plan_cache) plan_free(p);
    E.plan_cache.clear();
    cudaFree(E.d_keys[0]);
    cudaFree(E.d_keys[1]);
    cudaFree(E.d_sort_tmp);
    cudaFree(E.d_bounds);
    cudaFreeHost(E.h_bounds);
    cudaFreeHost(E.h_batch);
    cudaFree(E.d_counter);
    cudaFree(E.d_text);
    cudaFreeHost(E.h_counter);
    cudaFreeHost(E.h_keys);
    for (auto &s : E.stage) cudaFreeHost(s.buf);
    for (auto &s : E.stage) if (s.ev) cudaEventDestroy(s.ev);
    for (auto ev : E.ev_pool) cudaEventDestroy(ev);
    cudaEventDestroy(E.ev_a);
    cudaEventDestroy(E.ev_b);
    cudaStreamDestroy(E.scan_stream);
    cudaStreamDestroy(E.copy_stream);
    E = Engine();
}

int ensure_keys(uint64_t cap)
{
    Engine &E = g_engine;
    if (cap <= E.key_cap) return 0;
    uint64_t ncap = E.key_cap ? E.key_cap : (1ull << 20);
    while (ncap < cap) ncap *= 2;
    CK(cudaStreamSynchronize(E.scan_stream));
    cudaFree(E.d_keys[0]);
    cudaFree(E.d_keys[1]);
    E.d_keys[0] = E.d_keys[1] = nullptr;
    E.key_cap = 0;
    CK(cudaMalloc(&E.d_keys[0], ncap * sizeof(uint64_t)));
    CK(cudaMalloc(&E.d_keys[1], ncap * sizeof(uint64_t)));
    E.key_cap = ncap;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// plan compilation
// ---------------------------------------------------------------------------------------------
static uint32_t le32(const uint8_t *b, uint32_t n)
{
    uint32_t v = 0;
    for (uint32_t k = 0; k < n && k < 4; k++) v |= (uint32_t)b[k] << (8 * k);
    return v;
}

void plan_free(Plan *p)
{
    if (!p) return;
    cudaFree(p->d_pat_val);
    cudaFree(p->d_pat_mask);
    if (p->ac) ac_free_tables(p);
    p->magic = 0;
    delete p;
}

static bool border_free(const std::string &s, bool cs)
{
    const size_t m = s.size();
    if (m < 2) return true;
    std::vector<int> pi(m, 0);
    auto ch = [&](size_t i) { return cs ? (unsigned char)s[i] : lower_c((unsigned char)s[i]); };
    for (size_t i = 1; i < m; i++)
    {
        int k = pi[i - 1];
        while (k > 0 && ch(i) != ch((size_t)k)) k = pi[(size_t)k - 1];
        if (ch(i) == ch((size_t)k)) k++;
        pi[i] = k;
    }
    return pi[m - 1] == 0;
}

// Maps (reference function, params) to what the device has to enumerate.
Plan *plan_build(const search_params_t *P, int algo, bool only_matching)
{
    if (!P) return nullptr;
    Plan *pl = new Plan();
    pl->algo = algo;
    pl->case_sensitive = P->case_sensitive;
    pl->count_lines = P->count_lines_mode;
    if (algo == KREP_B200_ALGO_AC)
    {
        pl->is_ac = true;
        if (P->num_patterns > AC_MAX_PATTERNS)
        {
            set_error(-3, "too many patterns (%zu > %u)", (size_t)P->num_patterns, AC_MAX_PATTERNS);
            delete pl;
            return nullptr;
        }
        for (size_t k = 0; k < P->num_patterns; k++)
        {
            const size_t len = P->pattern_lens[k];
            if (len > 1024)
            {
                set_error(-3, "pattern %zu longer than 1024 bytes (krep.c:77)", k);
                delete pl;
                return nullptr;
            }
            pl->patterns.emplace_back(P->patterns[k] ? P->patterns[k] : "", len);
            pl->pat_lens.push_back((uint32_t)len);
        }
        pl->whole_word = P->whole_word ? 1 : 0;
        if (ac_build_tables(pl) != 0)
        {
            delete pl;
            return nullptr;
        }
        return pl;
    }
    // ---- single literal ----
    size_t m = P->pattern_len;
    if (algo == KREP_B200_ALGO_MEMCHR) m = m ? 1 : 0; // memchr_search reads pattern[0] only (krep.c:3902)
    if (m == 0 || m > 1024 || !P->pattern)
    {
        set_error(-3, "literal plan needs 1..1024 pattern bytes (got %zu)", m);
        delete pl;
        return nullptr;
    }
    pl->pattern.assign(P->pattern, m);
    pl->m = (uint32_t)m;
    // memchr_short_search -o walks first-byte hits, not occurrences (krep.c:4495)
    pl->emit_len = (algo == KREP_B200_ALGO_MEMCHR_SHORT && only_matching) ? 1u : (uint32_t)m;
    pl->border_free = border_free(pl->pattern, pl->case_sensitive);
    pl->built_only_matching = only_matching;
    if (P->whole_word)
    {
        // Kernels whose cursor also moves past a -w reject (kmp krep.c:1686, sse4.2 krep.c:4839-4848) need the
        // rejected occurrences in the list — but only if occurrences can overlap at all.  Prefix plans always tag.
        bool tag = pl->emit_len != pl->m;
        // the window kernels' tail sub-search re-evaluates -w against its sub-buffer (krep.c:5068): needs both halves
        if (algo == KREP_B200_ALGO_AVX2 || algo == KREP_B200_ALGO_AVX512 || algo == KREP_B200_ALGO_NEON) tag = true;
        if (!tag && !pl->border_free)
            tag = algo == KREP_B200_ALGO_KMP || (algo == KREP_B200_ALGO_SSE42 && !only_matching);
        pl->whole_word = tag ? 2 : 1;
    }
    std::vector<uint8_t> val(m), msk(m);
    const uint8_t *pb = (const uint8_t *)pl->pattern.data();
    for (size_t k = 0; k < m; k++)
    {
        msk[k] = (!pl->case_sensitive && is_alpha_c(pb[k])) ? 0xDF : 0xFF;
        val[k] = pb[k] & msk[k];
    }
    pl->fold = pl->case_sensitive ? 0xFFFFFFFFu : 0xDFDFDFDFu;
    if (pl->emit_len >= 7)
    {
        pl->filter = FILTER_ALIGNED4;
        for (int d = 0; d < 4; d++) pl->K[d] = le32(pb + d, 4) & pl->fold;
        pl->win_mask = 0xFFFFFFFFu;
        pl->filter_name = pl->case_sensitive ? "aligned4" : "aligned4-fold";
    }
    else
    {
        pl->filter = FILTER_WINDOW4;
        const uint32_t wl = pl->emit_len < 4 ? pl->emit_len : 4;
        pl->win_mask = wl == 4 ? 0xFFFFFFFFu : ((1u << (8 * wl)) - 1);
        pl->K[0] = le32(pb, wl) & pl->fold & pl->win_mask;
        pl->filter_name = pl->case_sensitive ? "window4" : "window4-fold";
    }
    if (cudaMalloc(&pl->d_pat_val, m) != cudaSuccess || cudaMalloc(&pl->d_pat_mask, m) != cudaSuccess ||
        cudaMemcpy(pl->d_pat_val, val.data(), m, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(pl->d_pat_mask, msk.data(), m, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        set_error(-2, "CUDA allocation failed while compiling the pattern");
        plan_free(pl);
        return nullptr;
    }
    return pl;
}

// ---------------------------------------------------------------------------------------------
// shard scan
// ---------------------------------------------------------------------------------------------
int launch_scan(const Plan *plan, const krep_b200_shard_t *sh, int want_positions, cudaStream_t stream)
{
    Engine &E = g_engine;
    if (((uintptr_t)sh->d_text & 15) != 0)
    {
        set_error(-3, "shard text pointer must be 16-byte aligned");
        return -3;
    }
    uint64_t own_end = sh->own_end < sh->avail_len ? sh->own_end : sh->avail_len;
    if (plan->is_ac)
    {
        AcLaunch a;
        a.text = (const uint8_t *)sh->d_text;
        a.avail_len = sh->avail_len;
        a.own_begin = sh->own_begin;
        a.own_end = own_end;
        a.global_offset = sh->global_offset;
        a.prev_byte = sh->prev_byte;
        a.next_byte = sh->next_byte;
        a.out = E.d_keys[0];
        a.cap = want_positions ? E.key_cap : 0;
        a.counter = E.d_counter;
        a.whole_word = plan->whole_word;
        a.want_positions = (uint32_t)want_positions;
        launch_ac(plan, a, stream);
        return 0;
    }
    LitDevParams p;
    memset(&p, 0, sizeof p);
    p.text = (const uint8_t *)sh->d_text;
    p.avail_len = sh->avail_len;
    p.own_begin = sh->own_begin;
    p.own_end = own_end;
    p.global_offset = sh->global_offset;
    p.prev_byte = sh->prev_byte;
    p.next_byte = sh->next_byte;
    uint64_t total_groups;
    if (plan->filter == FILTER_ALIGNED4)
    {
        total_groups = sh->avail_len / 16;
        p.tail_start = total_groups ? total_groups * 16 - 3 : 0;
    }
    else
    {
        total_groups = sh->avail_len >= 20 ? (sh->avail_len - 20) / 16 + 1 : 0; // vector + next word readable
        p.tail_start = total_groups * 16;
    }
    p.group_begin = sh->own_begin / 16;
    p.group_end = (own_end + 2) / 16 + 1;
    if (p.group_end > total_groups) p.group_end = total_groups;
    if (p.group_begin > p.group_end) p.group_begin = p.group_end;
    p.m = plan->m;
    p.emit_len = plan->emit_len;
    for (int d = 0; d < 4; d++) p.K[d] = plan->K[d];
    p.fold = plan->fold;
    p.win_mask = plan->win_mask;
    p.mulc[0] = 1u << 24;
    p.mulc[1] = 1u << 16;
    p.mulc[2] = 1u << 8;
    p.pat_val = plan->d_pat_val;
    p.pat_mask = plan->d_pat_mask;
    p.out = E.d_keys[0];
    p.cap = want_positions ? E.key_cap : 0;
    p.counter = E.d_counter;
    p.whole_word = plan->whole_word;
    p.want_positions = (uint32_t)want_positions;
    launch_literal(plan, p, stream);
    return 0;
}

int read_counter(cudaStream_t stream, uint64_t *count)
{
    Engine &E = g_engine;
    CK(cudaMemcpyAsync(E.h_counter, E.d_counter, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    *count = *E.h_counter;
    return 0;
}

int reset_counter(cudaStream_t stream)
{
    CK(cudaMemsetAsync(g_engine.d_counter, 0, 64, stream));
    return 0;
}

// Sorts the first n keys of d_keys[0]; the sorted list ends up in *sorted (either buffer).
int sort_keys(uint64_t n, int end_bit, cudaStream_t stream, const uint64_t **sorted)
{
    Engine &E = g_engine;
    *sorted = E.d_keys[0];
    if (n < 2) return 0;
    if (n > (uint64_t)INT32_MAX * 2)
    {
        set_error(-3, "occurrence list too long to sort (%llu)", (unsigned long long)n);
        return -3;
    }
    cub::DoubleBuffer<uint64_t> db(E.d_keys[0], E.d_keys[1]);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, need, db, (int64_t)n, 0, end_bit, stream));
    if (need > E.sort_tmp_bytes)
    {
        CK(cudaStreamSynchronize(stream));
        cudaFree(E.d_sort_tmp);
        E.d_sort_tmp = nullptr;
        E.sort_tmp_bytes = 0;
        CK(cudaMalloc(&E.d_sort_tmp, need));
        E.sort_tmp_bytes = need;
    }
    CK(cub::DeviceRadixSort::SortKeys(E.d_sort_tmp, need, db, (int64_t)n, 0, end_bit, stream));
    // (CUB's own launches — histogram, scan, one onesweep pass per 8 key bits — are library kernels and are not counted
    // by krep_b200_launch_count, which reports this library's hand-written kernels only)
    *sorted = db.Current();
    return 0;
}

static int bits_for(uint64_t v)
{
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

int key_end_bit(const Plan *plan, uint64_t max_offset)
{
    const int shift = plan->is_ac ? AC_END_SHIFT : LIT_TAG_BITS;
    int b = bits_for(max_offset) + shift;
    return b > 64 ? 64 : b;
}

// ---------------------------------------------------------------------------------------------
// -c on the device: line bounds of every occurrence (find_line_start / find_line_end, krep.c:363-408), so that the
// line-counting replay needs no host copy of the text.  One warp per (sorted) occurrence.  For literal plans the keys
// are sorted by start, so the backward scan stops at the previous occurrence and the forward scan at the next one
// (markers LB_SAME_AS_* are resolved by one pass on the host): the total work is O(text), however long the lines are.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_line_bounds(const uint8_t *__restrict__ text, uint64_t avail, uint64_t go,
                                                     const uint64_t *__restrict__ keys, uint64_t n, int is_ac, int has_prev,
                                                     int has_next, uint64_t *__restrict__ out)
{
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    auto start_of = [&](uint64_t key) -> uint64_t {
        if (!is_ac) return (key >> LIT_TAG_BITS) - go;
        return (key >> AC_END_SHIFT) - (1024 - ((key >> AC_LEN_SHIFT) & 1023)) - go;
    };
    const uint64_t s = start_of(keys[i]);
    const uint64_t lb = (!is_ac && i > 0) ? start_of(keys[i - 1]) : 0;       // backward scan covers [lb, s)
    const uint64_t ub = (!is_ac && i + 1 < n) ? start_of(keys[i + 1]) : avail; // forward scan covers [s, ub)
    // backward: last '\n' in [lb, s), 128 bytes per step (4 per lane)
    uint64_t ls = (!is_ac && i > 0) ? LB_SAME_AS_PREV : (has_prev ? LB_OUTSIDE_SHARD : go);
    for (uint64_t hi = s; hi > lb;)
    {
        const uint64_t w0 = hi >= lb + 128 ? hi - 128 : lb; // window [w0, hi)
        int best = -1;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const uint64_t p = w0 + (uint64_t)lane * 4 + k;
            if (p < hi && text[p] == '\n') best = lane * 4 + k;
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if (best >= 0)
        {
            ls = go + w0 + (uint64_t)best + 1;
            break;
        }
        hi = w0;
    }
    // forward: first '\n' in [s, ub)
    uint64_t le = (!is_ac && i + 1 < n) ? LB_SAME_AS_NEXT : (has_next ? LB_OUTSIDE_SHARD : go + avail);
    for (uint64_t lo = s; lo < ub; lo += 128)
    {
        int best = 1 << 20;
#pragma unroll
        for (int k = 3; k >= 0; k--)
        {
            const uint64_t p = lo + (uint64_t)lane * 4 + k;
            if (p < ub && text[p] == '\n') best = lane * 4 + k;
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best < (1 << 20))
        {
            le = go + lo + (uint64_t)best;
            break;
        }
    }
    if (lane == 0)
    {
        out[2 * i] = ls;
        out[2 * i + 1] = le;
    }
}

static int line_bounds(const Plan *plan, const krep_b200_shard_t *sh, const uint64_t *d_sorted, uint64_t n, cudaStream_t stream,
                       const uint64_t **d_bounds)
{
    Engine &E = g_engine;
    *d_bounds = nullptr;
    if (n == 0) return 0;
    if (2 * n > E.bounds_cap)
    {
        CK(cudaStreamSynchronize(stream));
        cudaFree(E.d_bounds);
        E.d_bounds = nullptr;
        E.bounds_cap = 0;
        const uint64_t cap = 2 * n + n / 4 + 1024;
        CK(cudaMalloc(&E.d_bounds, cap * sizeof(uint64_t)));
        E.bounds_cap = cap;
    }
    const uint64_t threads = n * 32;
    k_line_bounds<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>((const uint8_t *)sh->d_text, sh->avail_len, sh->global_offset,
                                                                        d_sorted, n, plan->is_ac ? 1 : 0, sh->prev_byte >= 0,
                                                                        sh->next_byte >= 0, E.d_bounds);
    CK(cudaGetLastError());
    count_launch();
    *d_bounds = E.d_bounds;
    return 0;
}

// Full single-shard scan: reset, launch (rerun with a larger list if it overflowed), sort.
int scan_shard(const Plan *plan, const krep_b200_shard_t *sh, int want_positions, cudaStream_t stream, ScanOut *out)
{
    Engine &E = g_engine;
    if (!stream) stream = E.scan_stream;
    if (want_positions && ensure_keys(1) != 0) return -2;
    reset_kernel_ms();
    for (int attempt = 0; attempt < 3; attempt++)
    {
        if (reset_counter(stream) != 0) return -2;
        CK(cudaEventRecord(E.ev_a, stream));
        int rc = launch_scan(plan, sh, want_positions, stream);
        if (rc != 0) return rc;
        CK(cudaEventRecord(E.ev_b, stream));
        CK(cudaGetLastError());
        uint64_t cnt = 0;
        if (read_counter(stream, &cnt) != 0) return -2;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, E.ev_a, E.ev_b);
        add_kernel_ms(ms);
        out->count = cnt;
        out->overflow = 0;
        if (!want_positions)
        {
            out->stored = 0;
            out->d_keys = nullptr;
            return 0;
        }
        if (cnt <= E.key_cap)
        {
            out->stored = cnt;
            rc = sort_keys(cnt, key_end_bit(plan, sh->global_offset + sh->avail_len), stream, &out->d_keys);
            if (rc == 0 && plan->count_lines) rc = line_bounds(plan, sh, out->d_keys, cnt, stream, &out->d_bounds);
            return rc;
        }
        out->overflow = 1;
        if (ensure_keys(cnt + cnt / 8 + 1024) != 0) return -2;
    }
    set_error(-4, "occurrence list kept overflowing");
    return -4;
}

// ---------------------------------------------------------------------------------------------
// synthetic corpus
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_corpus(const __grid_constant__ CorpusParams c, uint8_t *dst, uint64_t global_offset,
                                                uint64_t len)
{
    // dst[k] = byte(global_offset + k); global_offset is a multiple of 16 (checked on the host)
    const uint64_t groups = (len + 15) / 16;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x)
    {
        alignas(16) uint8_t b[16];
        corpus_fill16(c, global_offset + g * 16, b);
        const uint64_t off = g * 16;
        if (off + 16 <= len && (((uintptr_t)(dst + off)) & 15) == 0)
            *reinterpret_cast<uint4 *>(dst + off) = *reinterpret_cast<const uint4 *>(b);
        else
            for (int k = 0; k < 16 && off + k < len; k++) dst[off + k] = b[k];
    }
}

static int corpus_params(const krep_b200_corpus_spec_t *spec, CorpusParams *c)
{
    memset(c, 0, sizeof *c);
    c->seed = spec->seed;
    c->plant_seed = spec->plant_seed;
    c->plant_period = spec->plant_period;
    c->needle_len = spec->needle_len;
    c->flags = spec->flags;
    if (spec->needle_len > 64)
    {
        set_error(-3, "corpus needle longer than 64 bytes");
        return -3;
    }
    if (spec->plant_period && (spec->plant_period % 16 != 0 || spec->plant_period < 4ull * spec->needle_len + 32))
    {
        set_error(-3, "corpus plant_period must be a multiple of 16 and >= 4*needle_len+32");
        return -3;
    }
    if (spec->needle_len) memcpy(c->needle, spec->needle, spec->needle_len);
    return 0;
}

} // namespace kb

using namespace kb;

// ---------------------------------------------------------------------------------------------
// C ABI: lifetime, plans, shard scan, corpus
// ---------------------------------------------------------------------------------------------
extern "C" {

int krep_b200_init(int device)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    return engine_init(device);
}
void krep_b200_shutdown(void) { engine_shutdown(); }
int krep_b200_last_error(void) { return t_err; }
const char *krep_b200_last_error_string(void) { return t_errmsg; }
const char *krep_b200_version(void) { return "krep_b200 0.1.0 (sm_100a)"; }

float krep_b200_last_kernel_ms(void) { return t_kernel_ms; }
uint64_t krep_b200_launch_count(void) { return g_launches; }
void krep_b200_reset_launch_count(void) { g_launches = 0; }

krep_b200_plan_t *krep_b200_plan_create(const search_params_t *params, int algo)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!engine_ok()) return nullptr;
    return reinterpret_cast<krep_b200_plan_t *>(plan_build(params, resolve_algo(params, algo), krep_b200_get_only_matching()));
}
void krep_b200_plan_destroy(krep_b200_plan_t *plan)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    plan_free(reinterpret_cast<Plan *>(plan));
}
const char *krep_b200_plan_filter_name(const krep_b200_plan_t *plan)
{
    return plan ? reinterpret_cast<const Plan *>(plan)->filter_name.c_str() : "";
}

int krep_b200_scan_shard(const krep_b200_plan_t *plan, const krep_b200_shard_t *shard, int want_positions,
                         void *stream, krep_b200_device_result_t *out)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!engine_ok()) return -1;
    if (!plan || !shard || !out)
    {
        set_error(-3, "krep_b200_scan_shard: null argument");
        return -3;
    }
    ScanOut so;
    int rc = scan_shard(reinterpret_cast<const Plan *>(plan), shard, want_positions, (cudaStream_t)stream, &so);
    out->count = so.count;
    out->stored = so.stored;
    out->d_keys = so.d_keys;
    out->overflow = so.overflow;
    out->text_len = shard->global_offset + shard->avail_len;
    out->d_line_bounds = so.d_bounds;
    return rc;
}

int krep_b200_export_keys(const krep_b200_device_result_t *dev, void *d_dst, uint64_t max_keys, void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!dev || !d_dst) return -3;
    const uint64_t n = dev->stored < max_keys ? dev->stored : max_keys;
    if (n == 0) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : engine().scan_stream;
    CK(cudaMemcpyAsync(d_dst, dev->d_keys, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    if (!stream) CK(cudaStreamSynchronize(s));
    return 0;
}

uint64_t krep_b200_ac_key_end(uint64_t key) { return key >> AC_END_SHIFT; }
uint64_t krep_b200_ac_key_start(uint64_t key)
{
    const uint64_t len = 1024 - ((key >> AC_LEN_SHIFT) & 1023);
    return (key >> AC_END_SHIFT) - len;
}
uint32_t krep_b200_ac_key_pattern(uint64_t key) { return (uint32_t)(key & (AC_MAX_PATTERNS - 1)); }

int krep_b200_corpus_generate(const krep_b200_corpus_spec_t *spec, void *d_dst, uint64_t global_offset, uint64_t len,
                              void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(engine_mutex());
    clear_error();
    if (!engine_ok()) return -1;
    CorpusParams c;
    if (corpus_params(spec, &c) != 0) return -3;
    if (global_offset % 16 != 0)
    {
        set_error(-3, "corpus global_offset must be a multiple of 16");
        return -3;
    }
    if (len == 0) return 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : engine().scan_stream;
    const uint64_t groups = (len + 15) / 16;
    uint64_t blocks = (groups + 255) / 256;
    const uint64_t maxb = (uint64_t)engine().sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    k_corpus<<<(unsigned)blocks, 256, 0, s>>>(c, (uint8_t *)d_dst, global_offset, len);
    CK(cudaGetLastError());
    if (!stream) CK(cudaStreamSynchronize(s));
    return 0;
}

int krep_b200_corpus_generate_host(const krep_b200_corpus_spec_t *spec, void *dst, uint64_t global_offset, uint64_t len)
{
    CorpusParams c;
    if (corpus_params(spec, &c) != 0) return -3;
    uint8_t *o = (uint8_t *)dst;
    uint64_t i = global_offset, end = global_offset + len;
    while (i < end)
    {
        const uint64_t g0 = i & ~15ull;
        uint8_t b[16];
        corpus_fill16(c, g0, b);
        for (uint64_t k = i - g0; k < 16 && g0 + k < end; k++) o[g0 + k - global_offset] = b[k];
        i = g0 + 16;
    }
    return 0;
}

} // extern "C"
