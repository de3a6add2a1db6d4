/* arm_neon.h — scalar stand-ins for the FIVE NEON intrinsics krep.c's neon_search uses (krep.c:4527-4546), so that
 * the reference's ARM-only kernel can be compiled, unmodified, on this x86 box and used to pin oracle_neon_search.
 * TEST INFRASTRUCTURE ONLY (see oracle/krep_oracle.c); semantics per the Arm intrinsics reference:
 *   vdupq_n_u8(x)    all 16 lanes = x                 vld1q_u8(p) / vst1q_u8(p, v)   16-byte load / store
 *   vceqq_u8(a, b)   lane = 0xFF where equal else 0    vmaxvq_u8(v)                   maximum over the 16 lanes */
#ifndef KREP_B200_NEON_SHIM_H
#define KREP_B200_NEON_SHIM_H
#include <stdint.h>
#include <string.h>
typedef struct { uint8_t v[16]; } uint8x16_t;
static inline uint8x16_t vdupq_n_u8(uint8_t x) { uint8x16_t r; memset(r.v, x, 16); return r; }
static inline uint8x16_t vld1q_u8(const uint8_t *p) { uint8x16_t r; memcpy(r.v, p, 16); return r; }
static inline void vst1q_u8(uint8_t *p, uint8x16_t a) { memcpy(p, a.v, 16); }
static inline uint8x16_t vceqq_u8(uint8x16_t a, uint8x16_t b)
{
    uint8x16_t r;
    for (int i = 0; i < 16; i++) r.v[i] = a.v[i] == b.v[i] ? 0xFF : 0;
    return r;
}
static inline uint8_t vmaxvq_u8(uint8x16_t a)
{
    uint8_t m = 0;
    for (int i = 0; i < 16; i++) if (a.v[i] > m) m = a.v[i];
    return m;
}
#endif
