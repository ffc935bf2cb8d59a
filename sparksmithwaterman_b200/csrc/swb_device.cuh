// swb_device.cuh -- device helpers: DPX wrappers (inline PTX so that ptxas emits exactly
// VIADDMNMX / VIMNMX3 / VIMNMX and no operand-repacking PRMTs), profile loads, checkpoints.
#pragma once
#include <cstdint>
#include "swb_internal.h"

namespace swb {

__host__ __device__ __forceinline__ uint32_t pack2(int lo, int hi)
{
    return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
}

// per-halfword max(a + b, c)            -> VIADDMNMX.S16x2
__device__ __forceinline__ uint32_t viaddmax(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// per-halfword max(a + b, c, 0)         -> VIADDMNMX.S16x2.RELU
__device__ __forceinline__ uint32_t viaddmax_relu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2.relu %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// per-halfword a + b                     -> VIADD.16x2
__device__ __forceinline__ uint32_t viadd2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// per-halfword max(a, b)                -> VIMNMX.S16x2
__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// per-halfword max(a, b, c)             -> VIMNMX3.S16x2
__device__ __forceinline__ uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("{.reg .b32 t; max.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// running maximum over a lane's K cells of one column: ceil(K/2) VIMNMX3
template <int K>
__device__ __forceinline__ uint32_t colmax(uint32_t m, const uint32_t (&H)[K])
{
#pragma unroll
    for (int r = 0; r + 1 < K; r += 2) m = vmax3(m, H[r], H[r + 1]);
    if (K & 1) m = vmax2(m, H[K - 1]);
    return m;
}

// SUBSAMPLED running maximum: rows 0, 2, 4, ... and the lane's last row K-1 only.  A cell in an odd row r has
// the tracked cell (r+1, same column) right below it in the same lane, and H(r+1) >= H(r) + gap.
template <int K>
__device__ __forceinline__ uint32_t colmax_even(uint32_t m, const uint32_t (&H)[K])
{
    constexpr int NE = (K + 1) / 2;                 // even rows 0 .. 2*(NE-1)
#pragma unroll
    for (int e = 0; e + 1 < NE; e += 2) m = vmax3(m, H[2 * e], H[2 * e + 2]);
    if (NE & 1) {
        if (!(K & 1)) m = vmax3(m, H[2 * (NE - 1)], H[K - 1]);       // K even: last even row + the odd last row
        else m = vmax2(m, H[2 * (NE - 1)]);                           // K odd: the last even row IS row K-1
    } else if (!(K & 1)) m = vmax2(m, H[K - 1]);
    return m;
}

// KP profile words of one reference code for this lane: KP/4 conflict-free LDS.128
template <int KP>
__device__ __forceinline__ void load_profile(const uint32_t *p, uint32_t (&sv)[KP])
{
    const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) {
        const uint4 v = p4[q];
        sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
    }
}

// Block record of one lane (Geo<K>::RW words, contiguous): words 0..K-1 = H at the block start, word K = diag,
// words KW..KW+CB-1 = the seam.  The fill stages a group's 8 records in shared memory (lane t at stage + t*RW;
// 8 lanes x 16 B at a 144 B stride: conflict-free) and copies them out with coalesced STG.128.
template <int K>
__device__ __forceinline__ void stage_checkpoint(uint32_t *my_stage, const uint32_t (&H)[K], uint32_t diag)
{
    constexpr int KW = Geo<K>::KW;
#pragma unroll
    for (int q = 0; q < KW / 4; ++q) {
        uint4 v;
        v.x = (4 * q + 0 < K) ? H[(4 * q + 0 < K) ? 4 * q + 0 : 0] : ((4 * q + 0 == K) ? diag : 0u);
        v.y = (4 * q + 1 < K) ? H[(4 * q + 1 < K) ? 4 * q + 1 : 0] : ((4 * q + 1 == K) ? diag : 0u);
        v.z = (4 * q + 2 < K) ? H[(4 * q + 2 < K) ? 4 * q + 2 : 0] : ((4 * q + 2 == K) ? diag : 0u);
        v.w = (4 * q + 3 < K) ? H[(4 * q + 3 < K) ? 4 * q + 3 : 0] : ((4 * q + 3 == K) ? diag : 0u);
        *reinterpret_cast<uint4 *>(my_stage + 4 * q) = v;
    }
}

// the group's staged block (GL * RW words) -> global, 128 contiguous bytes per round of the 8 lanes
template <int K>
__device__ __forceinline__ void copy_out_block(uint32_t *dst, const uint32_t *stage, int t)
{
    constexpr int RW = Geo<K>::RW;
#pragma unroll
    for (int q = 0; q < RW / 4; ++q) {
        const uint4 v = *reinterpret_cast<const uint4 *>(stage + q * (GL * 4) + t * 4);
        *reinterpret_cast<uint4 *>(dst + q * (GL * 4) + t * 4) = v;
    }
}

__device__ __forceinline__ int half_of(uint32_t w, int half)
{
    return half ? (int)(int16_t)(w >> 16) : (int)(int16_t)(w & 0xffffu);
}

}  // namespace swb
