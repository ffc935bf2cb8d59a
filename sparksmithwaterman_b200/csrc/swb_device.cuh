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

// Block records in HBM.  The record of (block, lane t) is RP PIECES of 32 bytes: pieces 0 .. CKP-1 hold the checkpoint
// (words 0..K-1 = H at the block start, word K = diag), the last CB/8 the seam.  Piece p of a block's eight lanes is
// contiguous (256 bytes at ((blk * RP + p) * GL + t) * 32), so the fill writes every piece straight from registers with
// one 256-bit store per lane (STG.E.ENL2.256, sm_100) -- no shared-memory staging (staging + copy-out cost 108 LSU
// wavefronts per warp and block, the direct stores 40) -- and a traceback reads a lane's record as RP whole sectors.
// (16-byte chunks interleaved over the lanes gave the fill the same gain but made a tile nine half-used sectors:
// tile traceback 4.1 -> 5.4 ms per step.)
constexpr int REC_P = GL * 8;                                        // words between consecutive pieces of one lane
template <int K>
__device__ __forceinline__ uint32_t *rec_lane(uint32_t *rec, int64_t blk, int t)
{
    return rec + blk * (int64_t)(GL * Geo<K>::RW) + t * 8;
}
template <int K>
__device__ __forceinline__ const uint32_t *rec_lane(const uint32_t *rec, int64_t blk, int t)
{
    return rec + blk * (int64_t)(GL * Geo<K>::RW) + t * 8;
}
__device__ __forceinline__ void stg256(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                       uint32_t g, uint32_t h)
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
                 "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void ldg256(const uint32_t *p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// the lane's checkpoint (K cells + diag) -> pieces 0 .. CKP-1 of its record
template <int K>
__device__ __forceinline__ void store_checkpoint(uint32_t *lane_rec, const uint32_t (&H)[K], uint32_t diag)
{
#pragma unroll
    for (int p = 0; p < Geo<K>::CKP; ++p) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int w = 8 * p + e;
            v[e] = (w < K) ? H[w < K ? w : 0] : (w == K ? diag : 0u);
        }
        stg256(lane_rec + p * REC_P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    }
}

// pieces 0 .. CKP-1 of a lane's record -> the checkpointed words (w < K: cell w, w == K: diag) through fn(w, word)
template <int K, class Fn>
__device__ __forceinline__ void load_checkpoint(const uint32_t *lane_rec, bool ld, Fn &&fn)
{
#pragma unroll
    for (int p = 0; p < Geo<K>::CKP; ++p) {
        uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ld) ldg256(lane_rec + p * REC_P, v);
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (8 * p + e <= K) fn(8 * p + e, v[e]);
    }
}

__device__ __forceinline__ int half_of(uint32_t w, int half)
{
    return half ? (int)(int16_t)(w >> 16) : (int)(int16_t)(w & 0xffffu);
}

}  // namespace swb
