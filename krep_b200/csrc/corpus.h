// corpus.h — position-addressable synthetic corpus.  byte(i) is a pure function of
// (spec, i): every GPU can materialise its own shard (plus halo) in place, and the host twin produces
// the same bytes for tests and for the CPU baseline.  One definition, compiled for host and device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KB_HD __host__ __device__ __forceinline__
#else
#define KB_HD inline
#endif

namespace kb {

struct CorpusParams
{
    uint64_t seed, plant_seed, plant_period;
    uint32_t needle_len, flags;
    uint8_t needle[64];
};

KB_HD uint64_t mix64(uint64_t z) // splitmix64 finaliser
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// 64-entry alphabet: lowercase-heavy English-like letter mix, spaces, a few capitals, digits, a comma.
KB_HD uint8_t corpus_alpha(uint32_t idx)
{
    const char *A = "etaoinshrdlucmfwypvbgkjqxz" "eeettaaooiinnsshhrr" "        " "ETAOIN" "0123" ",";
    return (uint8_t)A[idx & 63];
}

static constexpr uint32_t CORPUS_LINE_SEG = 96; // exactly one '\n' per 96-byte segment -> lines of 1..191 bytes

// start offset of the needle planted for period-block b
KB_HD uint64_t plant_start(const CorpusParams &c, uint64_t b)
{
    return b * c.plant_period + mix64(c.plant_seed ^ (b * 0xD1B54A32D192ED03ull)) % c.plant_period;
}

// if position i is covered by (or glued to) the plant of block b that starts at s, writes the byte
KB_HD bool plant_byte(const CorpusParams &c, uint64_t b, uint64_t s, uint64_t i, uint8_t *out)
{
    const uint32_t L = c.needle_len;
    if (i >= s && i < s + L)
    {
        uint8_t ch = c.needle[i - s];
        if (c.flags & 1u) // RANDOM_CASE: per-plant, per-letter case flip
        {
            const uint64_t bits = mix64(c.plant_seed ^ b ^ 0xCA5ECA5Eull);
            const bool letter = (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z');
            if (letter && ((bits >> ((i - s) & 63)) & 1)) ch ^= 0x20;
        }
        *out = ch;
        return true;
    }
    if (c.flags & 2u) // EMBED_HALF: odd plants are glued inside a longer word, even plants are delimited
    {
        if ((s > 0 && i == s - 1) || i == s + L)
        {
            *out = (b & 1) ? (uint8_t)'x' : (uint8_t)' ';
            return true;
        }
    }
    return false;
}

// Fills out[0..16) with corpus bytes [i0, i0+16); i0 must be a multiple of 16.  plant_period is a
// multiple of 16 (checked by the caller), so one group never straddles two period blocks, one
// newline segment (96 = 6*16) or more than two 8-byte hash blocks: all hashes are hoisted.
KB_HD void corpus_fill16(const CorpusParams &c, uint64_t i0, uint8_t *out)
{
    const bool plants = c.plant_period && c.needle_len;
    uint64_t b = 0, s_prev = 0, s_cur = 0, s_next = 0;
    if (plants)
    {
        b = i0 / c.plant_period;
        s_prev = b > 0 ? plant_start(c, b - 1) : 0;
        s_cur = plant_start(c, b);
        s_next = plant_start(c, b + 1);
    }
    const uint64_t seg = i0 / CORPUS_LINE_SEG;
    const uint64_t nl = seg * CORPUS_LINE_SEG + mix64(c.seed ^ 0x4E4C4E4Cull ^ (seg * 0xA24BAED4963EE407ull)) % CORPUS_LINE_SEG;
    const uint64_t h0 = mix64(c.seed ^ ((i0 >> 3) * 0x9FB21C651E98DF25ull));
    const uint64_t h1 = mix64(c.seed ^ (((i0 >> 3) + 1) * 0x9FB21C651E98DF25ull));
    for (int k = 0; k < 16; k++)
    {
        const uint64_t i = i0 + k;
        uint8_t ch;
        if (plants)
        {
            // the previous block's plant may spill into this block and wins; then this block's plant;
            // then the glue byte in front of the next block's plant
            if (b > 0 && plant_byte(c, b - 1, s_prev, i, &ch)) { out[k] = ch; continue; }
            if (plant_byte(c, b, s_cur, i, &ch)) { out[k] = ch; continue; }
            if ((c.flags & 2u) && s_next == i + 1) { out[k] = ((b + 1) & 1) ? (uint8_t)'x' : (uint8_t)' '; continue; }
        }
        if (i == nl) { out[k] = (uint8_t)'\n'; continue; }
        const uint64_t h = k < 8 ? h0 : h1;
        out[k] = corpus_alpha((uint32_t)(h >> (6 * (k & 7))));
    }
}

} // namespace kb
