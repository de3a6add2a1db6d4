// scan_literal.cu — single-literal scan kernels for sm_100a.
//
// Replaces the inner loops of boyer_moore_search (krep.c:1294-1382), kmp_search (krep.c:1663-1763),
// memchr_search / memchr_short_search (krep.c:3918-4023, 4396-4500) and the simd_* searches
// (krep.c:4737-4866, 4914-5056, 5145-5255).  All of those enumerate the occurrences of one literal;
// they differ only in which occurrences they keep afterwards (overlap policy, -w, -c, -m), which
// semantics.cpp replays over the sorted occurrence list.  So the device work is one thing:
// emit every position where the literal occurs, exactly once, at HBM speed.
//
// Design (HBM-bound byte scan, no tensor cores — nothing here is a contraction):
//   * every thread streams 16-byte vectors (LDG.128, coalesced: a warp reads 512 contiguous bytes),
//     UNROLL vectors in flight per thread; the grid is sized to the SM count x resident CTAs and walks
//     the shard with a grid stride, so neighbouring CTAs read neighbouring DRAM pages;
//   * the hot loop only FILTERS, with ~1 integer op per corpus byte:
//       ALIGNED4 (pattern_len >= 7): an occurrence at p fully contains the aligned word j = ceil(p/4),
//         which then equals P[d..d+4) with d = 4j-p in 0..3.  Each aligned text word is compared with
//         those four constants — no halo, no shuffles, no shared memory; case-insensitive search ANDs
//         the word with 0xDFDFDFDF first (a superset filter: folds letters exactly, aliases a few
//         punctuation bytes, verified exactly afterwards);
//       WINDOW4 (pattern_len < 7): the 4-byte window at every byte offset, built with funnel shifts from
//         the thread's own words plus the first word of the next vector, is compared with P[0..4) under a
//         length mask;
//   * a thread whose vector contains a candidate (rare) calls the out-of-line verifier, which compares
//     all pattern bytes under the exact per-byte case mask, evaluates the whole-word boundary against the
//     global text (shard context bytes at the edges, so results equal the reference's single-chunk run,
//     SURVEY §8 a12), checks shard ownership by start offset and appends one 64-bit key.
//
// Algorithmic traffic: 1 byte read per corpus byte (+ 8 B written per occurrence).
#include <mutex>
#include "common.h"

namespace kb {

__device__ __forceinline__ bool dev_is_word(int c)
{
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}

// Exact check of one candidate start + emission. Out of line on purpose: keeps the streaming loop's
// register footprint small; executed for a vanishing fraction of positions on low-hit-rate corpora.
// Returns 1 when the occurrence was only counted (count-only launch), else 0.
__device__ __noinline__ unsigned verify_emit(const LitDevParams &p, long long cand)
{
    if (cand < (long long)p.own_begin || cand >= (long long)p.own_end) return 0;
    const uint64_t c = (uint64_t)cand;
    if (c + p.emit_len > p.avail_len) return 0;
    const uint8_t *t = p.text + c;
    const uint8_t *val = p.pat_val, *msk = p.pat_mask;
    for (uint32_t k = 0; k < p.emit_len; k++)
        if ((t[k] & msk[k]) != val[k]) return 0;
    unsigned full = 1;
    if (p.m > p.emit_len)
    {
        if (c + p.m > p.avail_len) full = 0;
        else
            for (uint32_t k = p.emit_len; k < p.m; k++)
                if ((t[k] & msk[k]) != val[k]) { full = 0; break; }
    }
    unsigned ww_tag = 3; // ws_ok << 1 | we_ok
    if (p.whole_word)
    {
        const uint64_t e = c + p.m;
        const int pb = c > 0 ? (int)t[-1] : p.prev_byte;
        const int nb = e < p.avail_len ? (int)p.text[e] : p.next_byte;
        ww_tag = (dev_is_word(pb) ? 0u : 2u) | (dev_is_word(nb) ? 0u : 1u);
        if (p.whole_word == 1 && ww_tag != 3) return 0;
    }
    if (p.want_positions)
    {
        const unsigned long long slot = atomicAdd(p.counter, 1ULL);
        if (slot < p.cap) p.out[slot] = ((p.global_offset + c) << LIT_TAG_BITS) | (full << 2) | ww_tag;
        return 0;
    }
    return 1;
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *ptr)
{
    return __ldcs(ptr); // ld.global.cs: streamed once, evict-first
}

// ------------------------------------------------------------------------------------ ALIGNED4
template <bool FOLD>
__device__ __forceinline__ bool hit_word(uint32_t w, uint32_t fold, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    if (FOLD) w &= fold;
    return (w == k0) | (w == k1) | (w == k2) | (w == k3);
}
template <bool FOLD>
__device__ __forceinline__ bool hit_vec(const uint4 &v, uint32_t fold, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    return hit_word<FOLD>(v.x, fold, k0, k1, k2, k3) | hit_word<FOLD>(v.y, fold, k0, k1, k2, k3) |
           hit_word<FOLD>(v.z, fold, k0, k1, k2, k3) | hit_word<FOLD>(v.w, fold, k0, k1, k2, k3);
}

__device__ __noinline__ unsigned slow_aligned4(const LitDevParams &p, uint64_t group, uint4 v)
{
    unsigned n = 0;
    const uint32_t w[4] = {v.x & p.fold, v.y & p.fold, v.z & p.fold, v.w & p.fold};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (w[k] == p.K[d]) n += verify_emit(p, (long long)(group * 16 + 4 * k) - d);
    return n;
}

__device__ __forceinline__ void tail_and_count(const LitDevParams &p, unsigned long long local_cnt)
{
    if (blockIdx.x == 0 && threadIdx.x < 32 && p.avail_len >= p.emit_len)
    {
        const uint64_t last = p.avail_len - p.emit_len;
        for (uint64_t s = p.tail_start + threadIdx.x; s <= last; s += 32) local_cnt += verify_emit(p, (long long)s);
    }
    if (!p.want_positions)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_cnt += __shfl_xor_sync(0xffffffffu, local_cnt, o);
        if ((threadIdx.x & 31) == 0 && local_cnt) atomicAdd(p.counter, local_cnt);
    }
}

template <bool FOLD, int UNROLL>
__global__ void __launch_bounds__(256, 4) k_lit_aligned4(const __grid_constant__ LitDevParams p)
{
    const uint4 *__restrict__ t4 = reinterpret_cast<const uint4 *>(p.text);
    const uint32_t k0 = p.K[0], k1 = p.K[1], k2 = p.K[2], k3 = p.K[3], fold = p.fold;
    unsigned long long local_cnt = 0;
    const uint64_t tile = (uint64_t)blockDim.x * UNROLL;
    const uint64_t stride = (uint64_t)gridDim.x * tile;
    uint64_t g0 = p.group_begin + (uint64_t)blockIdx.x * tile;
    for (; g0 + tile <= p.group_end; g0 += stride)
    {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = ld_stream(t4 + g0 + (uint64_t)u * blockDim.x + threadIdx.x);
        bool hit = false;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) hit |= hit_vec<FOLD>(v[u], fold, k0, k1, k2, k3);
        if (hit)
        {
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (hit_vec<FOLD>(v[u], fold, k0, k1, k2, k3))
                    local_cnt += slow_aligned4(p, g0 + (uint64_t)u * blockDim.x + threadIdx.x, v[u]);
        }
    }
    if (g0 < p.group_end) // the one ragged tile
    {
        for (int u = 0; u < UNROLL; u++)
        {
            const uint64_t g = g0 + (uint64_t)u * blockDim.x + threadIdx.x;
            if (g < p.group_end)
            {
                const uint4 v = ld_stream(t4 + g);
                if (hit_vec<FOLD>(v, fold, k0, k1, k2, k3)) local_cnt += slow_aligned4(p, g, v);
            }
        }
    }
    tail_and_count(p, local_cnt);
}

// ------------------------------------------------------------------------------------ WINDOW4
// The 4-byte window at byte offset 4k+r is (lo >> 8r) | (hi << (32-8r)).  A funnel shift would put it on the ALU
// pipe next to the compares, which is what bounds this kernel (SHF/LOP3/ISETP all issue there at half rate).  The
// same value is umulhi(lo, 2^(32-8r)) + hi * 2^(32-8r) — an IMAD.HI and an IMAD on the otherwise idle FMA pipe —
// so per text word the ALU pipe only sees the case fold (one LOP3, -i only) and the four compares.  The
// multipliers come from kernel parameters so that the compiler cannot strength-reduce them back into shifts.
template <bool MASKED>
__device__ __forceinline__ bool hit_pair(uint32_t lo, uint32_t hi, uint32_t mask, uint32_t k0, uint32_t c1, uint32_t c2,
                                         uint32_t c3)
{
    uint32_t x0 = lo, x1 = hi * c1 + __umulhi(lo, c1), x2 = hi * c2 + __umulhi(lo, c2), x3 = hi * c3 + __umulhi(lo, c3);
    if (MASKED)
    {
        x0 &= mask; x1 &= mask; x2 &= mask; x3 &= mask;
    }
    return (x0 == k0) | (x1 == k0) | (x2 == k0) | (x3 == k0);
}
template <bool FOLD, bool MASKED>
__device__ __forceinline__ bool hit_vec_w(const uint4 &v, uint32_t nx, uint32_t fold, uint32_t mask, uint32_t k0,
                                          uint32_t c1, uint32_t c2, uint32_t c3)
{
    uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w, w4 = nx;
    if (FOLD)
    {
        w0 &= fold; w1 &= fold; w2 &= fold; w3 &= fold; w4 &= fold;
    }
    return hit_pair<MASKED>(w0, w1, mask, k0, c1, c2, c3) | hit_pair<MASKED>(w1, w2, mask, k0, c1, c2, c3) |
           hit_pair<MASKED>(w2, w3, mask, k0, c1, c2, c3) | hit_pair<MASKED>(w3, w4, mask, k0, c1, c2, c3);
}

__device__ __noinline__ unsigned slow_window4(const LitDevParams &p, uint64_t group, uint4 v, uint32_t nx)
{
    unsigned n = 0;
    const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
    const uint32_t mask = p.fold & p.win_mask, k0 = p.K[0];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const uint32_t win = r == 0 ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * r);
            if ((win & mask) == k0) n += verify_emit(p, (long long)(group * 16 + 4 * k + r));
        }
    return n;
}

// Main loads of the window kernel: the sector that holds a warp's "next word" is touched twice (once as lane 31's
// next-word load, once as the following warp's vector), so these loads keep the default L2 policy instead of
// evict-first; KREP_B200_W4_CS=1 at build time restores the streaming hint for comparison.
#ifdef KREP_B200_W4_CS
#define WLOAD(q) ld_stream(q)
#else
#define WLOAD(q) __ldg(q)
#endif
template <bool FOLD, bool MASKED, int UNROLL>
__global__ void __launch_bounds__(256, 4) k_lit_window4(const __grid_constant__ LitDevParams p)
{
    const uint4 *__restrict__ t4 = reinterpret_cast<const uint4 *>(p.text);
    const uint32_t fold = p.fold, mask = p.win_mask, k0 = p.K[0];
    const uint32_t c1 = p.mulc[0], c2 = p.mulc[1], c3 = p.mulc[2]; // 2^24, 2^16, 2^8
    const bool lane31 = (threadIdx.x & 31) == 31;
    unsigned long long local_cnt = 0;
    const uint64_t tile = (uint64_t)blockDim.x * UNROLL;
    const uint64_t stride = (uint64_t)gridDim.x * tile;
    uint64_t g0 = p.group_begin + (uint64_t)blockIdx.x * tile;
    for (; g0 + tile <= p.group_end; g0 += stride)
    {
        uint4 v[UNROLL];
        uint32_t nx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const uint4 *q = t4 + g0 + (uint64_t)u * blockDim.x + threadIdx.x;
            v[u] = WLOAD(q);
            // the word after the vector is the next lane's v.x: only lane 31 has to load it (one sector per 512 bytes
            // instead of one 4-byte request per lane)
            nx[u] = lane31 ? __ldg(reinterpret_cast<const uint32_t *>(q + 1)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const uint32_t nb = __shfl_down_sync(0xffffffffu, v[u].x, 1);
            if (!lane31) nx[u] = nb;
        }
        uint32_t hm = 0; // bit u: vector u holds a candidate
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            hm |= hit_vec_w<FOLD, MASKED>(v[u], nx[u], fold, mask, k0, c1, c2, c3) ? (1u << u) : 0u;
        if (hm)
        {
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if ((hm >> u) & 1u) local_cnt += slow_window4(p, g0 + (uint64_t)u * blockDim.x + threadIdx.x, v[u], nx[u]);
        }
    }
    if (g0 < p.group_end)
    {
        for (int u = 0; u < UNROLL; u++)
        {
            const uint64_t g = g0 + (uint64_t)u * blockDim.x + threadIdx.x;
            if (g < p.group_end)
            {
                const uint4 v = ld_stream(t4 + g);
                const uint32_t nx = __ldg(reinterpret_cast<const uint32_t *>(t4 + g + 1));
                if (hit_vec_w<FOLD, MASKED>(v, nx, fold, mask, k0, c1, c2, c3)) local_cnt += slow_window4(p, g, v, nx);
            }
        }
    }
    tail_and_count(p, local_cnt);
}

// ------------------------------------------------------------------------------------ launch
static int g_occ[6] = {0, 0, 0, 0, 0, 0}; // identical on every device of the box (all sm_100)

template <typename K>
static int occupancy(K kernel)
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, 0);
    return nb > 0 ? nb : 1;
}

void launch_literal(const Plan *plan, const LitDevParams &p, int sm_count, cudaStream_t s)
{
    constexpr int UNROLL = 4;
    static std::once_flag once;
    std::call_once(once, [] {
        g_occ[0] = occupancy(k_lit_aligned4<false, UNROLL>);
        g_occ[1] = occupancy(k_lit_aligned4<true, UNROLL>);
        g_occ[2] = occupancy(k_lit_window4<false, false, UNROLL>);
        g_occ[3] = occupancy(k_lit_window4<true, false, UNROLL>);
        g_occ[4] = occupancy(k_lit_window4<false, true, UNROLL>);
        g_occ[5] = occupancy(k_lit_window4<true, true, UNROLL>);
    });
    const uint64_t groups = p.group_end > p.group_begin ? p.group_end - p.group_begin : 0;
    const uint64_t tile = 256ull * UNROLL;
    uint64_t tiles = (groups + tile - 1) / tile;
    if (tiles == 0) tiles = 1; // still need the tail warp
    const bool folded = plan->fold != 0xFFFFFFFFu, masked = plan->win_mask != 0xFFFFFFFFu;
    const int which = plan->filter == FILTER_WINDOW4 ? 2 + (folded ? 1 : 0) + (masked ? 2 : 0) : (folded ? 1 : 0);
    uint64_t resident = (uint64_t)sm_count * g_occ[which];
    const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
    if (which == 2)
        k_lit_window4<false, false, UNROLL><<<grid, 256, 0, s>>>(p);
    else if (which == 3)
        k_lit_window4<true, false, UNROLL><<<grid, 256, 0, s>>>(p);
    else if (which == 4)
        k_lit_window4<false, true, UNROLL><<<grid, 256, 0, s>>>(p);
    else if (which == 5)
        k_lit_window4<true, true, UNROLL><<<grid, 256, 0, s>>>(p);
    else if (which == 1)
        k_lit_aligned4<true, UNROLL><<<grid, 256, 0, s>>>(p);
    else
        k_lit_aligned4<false, UNROLL><<<grid, 256, 0, s>>>(p);
    count_launch();
}

} // namespace kb
