// swb_wide.cu -- the general ("wide") path: int32 scores, any read length, any scores that do
// not overflow int32, 8-bit symbol codes (any alphabet).  Used for long pairs (BASELINE config 3:
// 100 kbp x 100 kbp) and for every input outside the s16x2 short-read domain.
//
// Intra-pair parallelism: the matrix is cut in BANDS of BH = 32 lanes x KL rows.  One warp
// sweeps a band along the reference as a 32-lane anti-diagonal wavefront (lane t computes
// column s - t + 1 at step s; the boundary row moves down the lanes with __shfl_up_sync),
// consecutive bands of a pair run concurrently on different warps, coupled through the
// band's bottom row in HBM (brow) and a per-band progress counter: band b may compute a
// 32-column chunk once band b-1 has published those columns.  Work is handed out by a global
// ticket in (band, pair) order, so a warp only ever waits for a lower ticket, which is held by
// a running warp: no deadlock, no cooperative launch.
//
// As in the short path the fill is score-only (SmithWaterman.java:157-187, :217-252) and leaves
// register checkpoints every WCB steps + per-lane tile maxima; locate/trace recompute single
// (band, block) tiles -- exact, because a tile's left edge is a checkpoint and its top edge is
// the previous band's bottom row.  Traceback = GetAlignment.call (SmithWaterman.java:354-436).
#include "swb_internal.h"

#include <algorithm>
#include <cstdlib>

namespace swb {

using namespace wide;

namespace {

__device__ __forceinline__ int ld_cg(const int32_t *p) { return __ldcg(p); }
// progress counters: release store by the lane that wrote the band's bottom row, acquire load by the polling lane
__device__ __forceinline__ void st_release(int32_t *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int32_t *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct WCtx {
    const uint8_t *ref;     // codes of this pair's reference
    const uint8_t *read;    // codes of this pair's read
    int n, m;
    int match, mismatch, gap;
    int n_blocks;           // blocks per band = ceil((n + 31) / WCB)
    int32_t *brow;          // [bands][n + 1] bottom rows of this pair
    int32_t *ck;            // [bands][n_blocks][KL + 1][32]
    int32_t *tmx;           // [bands][n_blocks][32]
};

__device__ __forceinline__ WCtx make_ctx(const WideParams &P, int pair)
{
    WCtx C;
    const int ro = P.pair_ref[pair], rd = P.pair_read[pair];
    C.ref = P.ref_codes + P.ref_off[ro];
    C.n = (int)(P.ref_off[ro + 1] - P.ref_off[ro]);
    C.read = P.read_codes + P.read_off[rd];
    C.m = (int)(P.read_off[rd + 1] - P.read_off[rd]);
    C.match = P.match; C.mismatch = P.mismatch; C.gap = P.gap;
    C.n_blocks = (C.n + WL - 1 + WCB - 1) / WCB;
    if (C.n_blocks < 1) C.n_blocks = 1;
    C.brow = P.brow + P.brow_off[pair];
    C.ck = P.ck + P.blk_off[pair] * (int64_t)((KL + 1) * WL);
    C.tmx = P.tmx + P.blk_off[pair] * (int64_t)WL;
    return C;
}

// read codes of this lane's rows in `band` (0x100 + r: matches nothing) and row validity
__device__ __forceinline__ void load_rows(const WCtx &C, int band, int lane, int (&rc)[KL], bool &all_valid)
{
    all_valid = true;
#pragma unroll
    for (int r = 0; r < KL; ++r) {
        const int row = band * BH + lane * KL + r;          // 0-based
        if (row < C.m) rc[r] = C.read[row];
        else { rc[r] = 0x100 + r; all_valid = false; }
    }
}

// One 32-step chunk of a band.  `tbuf` holds, in lane L, the top-boundary value of column
// s0 + 1 + L (bottom row of the band above; 0 for band 0).  sink(u, top, H, valid, j).
// Reference codes reach the lanes through registers, not per-step loads: lane L holds the code of
// column s0 - 31 + L (cprev) and of column s0 + 1 + L (ccur); at step u lane t needs column
// s0 + u - t + 1, i.e. ccur[u - t] if u >= t, else cprev[32 + u - t] -- one SEL + one SHFL.
__device__ __forceinline__ int code_prefetch(const WCtx &C, int s0, int lane)
{
    const int j = s0 + 1 + lane;
    return (j >= 1 && j <= C.n) ? (int)C.ref[j - 1] : 0x200;
}

template <class Sink>
__device__ __forceinline__ void wide_chunk(const WCtx &C, int s0, int lane, int tbuf, int cprev, int ccur,
                                           const int (&rc)[KL], int (&H)[KL], int &diag, Sink &&sink)
{
    if (s0 >= WL - 1 && s0 + 32 <= C.n) {
        // interior chunk: every lane's 32 columns are inside the matrix -> no per-step branch, steps overlap.
        // (Rows beyond the read's end compute garbage below the real rows, as in the general loop.)
#pragma unroll 4
        for (int u = 0; u < 32; ++u) {
            int top = __shfl_up_sync(0xffffffffu, H[KL - 1], 1);
            const int t0 = __shfl_sync(0xffffffffu, tbuf, u);
            if (lane == 0) top = t0;
            const int c = __shfl_sync(0xffffffffu, lane <= u ? ccur : cprev, (u - lane) & 31);
            int nw = diag, nn = top;
#pragma unroll
            for (int r = 0; r < KL; ++r) {
                const int sc = (rc[r] == c) ? C.match : C.mismatch;
                const int pre = __viaddmax_s32_relu(H[r], C.gap, nw + sc);
                nw = H[r];
                H[r] = __viaddmax_s32(nn, C.gap, pre);
                nn = H[r];
            }
            sink(u, top, H, true, s0 + u - lane + 1);
            diag = top;
        }
        return;
    }
#pragma unroll 1
    for (int u = 0; u < 32; ++u) {
        const int s = s0 + u;
        int top = __shfl_up_sync(0xffffffffu, H[KL - 1], 1);
        const int t0 = __shfl_sync(0xffffffffu, tbuf, u);
        if (lane == 0) top = t0;
        const int c = __shfl_sync(0xffffffffu, lane <= u ? ccur : cprev, (u - lane) & 31);
        const int j = s - lane + 1;
        const bool valid = (j >= 1) && (j <= C.n);
        if (valid) {
            int nw = diag, nn = top;
#pragma unroll
            for (int r = 0; r < KL; ++r) {
                const int sc = (rc[r] == c) ? C.match : C.mismatch;
                const int pre = __viaddmax_s32_relu(H[r], C.gap, nw + sc);     // max(W+gap, NW+s, 0)
                nw = H[r];
                H[r] = __viaddmax_s32(nn, C.gap, pre);                          // max(N+gap, pre)
                nn = H[r];
            }
        }
        sink(u, top, H, valid, j);
        diag = top;
    }
}

__device__ __forceinline__ int top_prefetch(const WCtx &C, int band, int s0, int lane)
{
    if (band == 0) return 0;
    const int j = s0 + 1 + lane;
    return (j <= C.n) ? ld_cg(C.brow + (int64_t)(band - 1) * (C.n + 1) + j) : 0;
}

__device__ __forceinline__ void load_wide_state(const WCtx &C, int band, int blk, int lane, int (&H)[KL], int &diag)
{
    if (blk == 0) {
#pragma unroll
        for (int r = 0; r < KL; ++r) H[r] = 0;
        diag = 0;
        return;
    }
    const int32_t *p = C.ck + ((int64_t)band * C.n_blocks + blk) * ((KL + 1) * WL) + lane;
#pragma unroll
    for (int r = 0; r < KL; ++r) H[r] = p[r * WL];
    diag = p[KL * WL];
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Fill.  PROF = true (alphabets of up to 8 symbols): the substitution score comes from a per-warp
// shared-memory profile of the band ([code][KL/4][lane][4], conflict-free LDS.128) and the NW + s add is an
// IMAD on the FMA pipe, leaving two DPX ops per cell on the integer pipe.  PROF = false: compare + select.
template <bool PROF, int NC>
__global__ void __launch_bounds__(128) wide_fill_kernel(const WideParams P, const int2 *items, int n_items,
                                                         uint32_t *ticket, int one, int lag_chunks)
{
    extern __shared__ __align__(16) int32_t wprof_all[];          // PROF: per warp [NC codes][KL/4][32 lanes][4], NC = 4 or 8
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *wprof = wprof_all + warp * (NC * KL * WL);
    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(ticket, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= (uint32_t)n_items) break;
        const int pair = items[it].x, band = items[it].y;
        const WCtx C = make_ctx(P, pair);
        int rc[KL]; bool all_valid;
        load_rows(C, band, lane, rc, all_valid);
        const bool warp_all_valid = __all_sync(0xffffffffu, all_valid);
        if (PROF) {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int r = 0; r < KL; ++r)
                    wprof[((c * (KL / 4) + (r >> 2)) * WL + lane) * 4 + (r & 3)] = (rc[r] == c) ? C.match : C.mismatch;
            __syncwarp();
        }
        int H[KL], diag = 0;
#pragma unroll
        for (int r = 0; r < KL; ++r) H[r] = 0;
        int tmax = 0, bmax = 0;
        const int nsteps = C.n + WL - 1;
        int32_t *my_brow = C.brow + (int64_t)band * (C.n + 1);
        const int32_t *prog_up = band > 0 ? P.prog + P.band_off[pair] + band - 1 : nullptr;
        int32_t *prog_me = P.prog + P.band_off[pair] + band;
        int seen = 0;                                             // columns of the band above known complete
        int cprev = 0x200;
        for (int s0 = 0; s0 < nsteps; s0 += 32) {
            const int ccur = code_prefetch(C, s0, lane);
            if (band > 0) {
                const int need = min(C.n, s0 == 0 ? 32 * lag_chunks : s0 + 32);   // start with `lag_chunks` of slack: a late chunk above no longer stalls the whole cascade below
                if (seen < need) {
                    // lane 0 acquires; the other lanes' boundary loads (ld.cg, L2) are issued after the loop's
                    // branch has resolved on the acquired value, i.e. after the publisher's release
                    if (lane == 0) {
                        int v = ld_acquire(prog_up);
                        while (v < need) { __nanosleep(32); v = ld_acquire(prog_up); }
                        seen = v;
                    }
                    seen = __shfl_sync(0xffffffffu, seen, 0);
                }
            }
            const int tbuf = top_prefetch(C, band, s0, lane);
            if (PROF && warp_all_valid && s0 >= WL - 1 && s0 + 32 <= C.n) {
                // interior chunk: every lane's 32 columns are inside the matrix and every row is a read row -> no
                // per-step branches, so the unrolled steps overlap (profile loads of step u+1 under the chain of step u)
#pragma unroll 8
                for (int u = 0; u < 32; ++u) {
                    int top = __shfl_up_sync(0xffffffffu, H[KL - 1], 1);
                    const int t0 = __shfl_sync(0xffffffffu, tbuf, u);
                    if (lane == 0) top = t0;
                    const int c = __shfl_sync(0xffffffffu, lane <= u ? ccur : cprev, (u - lane) & 31);
                    int sv[KL];
                    const int4 *pp = reinterpret_cast<const int4 *>(wprof) + (c & (NC - 1)) * (KL / 4) * WL + lane;
#pragma unroll
                    for (int q = 0; q < KL / 4; ++q) {
                        const int4 v = pp[q * WL];
                        sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
                    }
                    int nw = diag, nn = top;
#pragma unroll
                    for (int r = 0; r < KL; ++r) {
                        const int tt = nw * one + sv[r];
                        const int pre = __viaddmax_s32_relu(H[r], C.gap, tt);
                        nw = H[r];
                        H[r] = __viaddmax_s32(nn, C.gap, pre);
                        nn = H[r];
                    }
                    if (lane == WL - 1) my_brow[s0 + u - lane + 1] = H[KL - 1];
#pragma unroll
                    for (int r = 0; r < KL; r += 2) tmax = __vimax3_s32(tmax, H[r], H[r + 1]);
                    diag = top;
                }
            } else {
#pragma unroll 4
            for (int u = 0; u < 32; ++u) {
                const int s = s0 + u;
                int top = __shfl_up_sync(0xffffffffu, H[KL - 1], 1);
                const int t0 = __shfl_sync(0xffffffffu, tbuf, u);
                if (lane == 0) top = t0;
                const int c = __shfl_sync(0xffffffffu, lane <= u ? ccur : cprev, (u - lane) & 31);
                const int j = s - lane + 1;
                if ((j >= 1) && (j <= C.n)) {
                    int sv[KL];
                    if (PROF) {
                        const int4 *pp = reinterpret_cast<const int4 *>(wprof) + (c & (NC - 1)) * (KL / 4) * WL + lane;
#pragma unroll
                        for (int q = 0; q < KL / 4; ++q) {
                            const int4 v = pp[q * WL];
                            sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < KL; ++r) sv[r] = (rc[r] == c) ? C.match : C.mismatch;
                    }
                    int nw = diag, nn = top;
#pragma unroll
                    for (int r = 0; r < KL; ++r) {
                        const int tt = PROF ? nw * one + sv[r] : nw + sv[r];           // IMAD (FMA pipe) when PROF
                        const int pre = __viaddmax_s32_relu(H[r], C.gap, tt);           // max(W+gap, NW+s, 0)
                        nw = H[r];
                        H[r] = __viaddmax_s32(nn, C.gap, pre);                          // max(N+gap, pre)
                        nn = H[r];
                    }
                    if (lane == WL - 1) my_brow[j] = H[KL - 1];
                    if (all_valid) {
#pragma unroll
                        for (int r = 0; r < KL; r += 2) tmax = __vimax3_s32(tmax, H[r], H[r + 1]);
                    } else {
#pragma unroll
                        for (int r = 0; r < KL; ++r) if (rc[r] < 0x100) tmax = max(tmax, H[r]);
                    }
                }
                diag = top;
            }
            }
            cprev = ccur;
            // publish: lane 31 (the only writer of the bottom row) has finished every column <= s0 + 1
            if (lane == WL - 1) st_release(prog_me, min(C.n, max(0, s0 + 1)));
            const int s_next = s0 + 32;
            if ((s_next % WCB) == 0) {
                const int b = s_next / WCB;
                C.tmx[((int64_t)band * C.n_blocks + b - 1) * WL + lane] = tmax;
                bmax = max(bmax, tmax);
                tmax = 0;
                if (s_next < nsteps) {
                    int32_t *p = C.ck + ((int64_t)band * C.n_blocks + b) * ((KL + 1) * WL) + lane;
#pragma unroll
                    for (int r = 0; r < KL; ++r) p[r * WL] = H[r];
                    p[KL * WL] = diag;
                }
            }
        }
        {
            const int s_end = ((nsteps + 31) >> 5) << 5;
            if ((s_end % WCB) != 0) {
                C.tmx[((int64_t)band * C.n_blocks + s_end / WCB) * WL + lane] = tmax;
                bmax = max(bmax, tmax);
            }
        }
        if (lane == WL - 1) st_release(prog_me, C.n);
#pragma unroll
        for (int o = 16; o; o >>= 1) bmax = max(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        if (lane == 0 && bmax > 0)
            atomicMax(P.scores + (int64_t)P.pair_ref[pair] * P.n_reads + P.pair_read[pair], bmax);
    }
}

// one thread per (pair, band, block): lanes whose tile maximum equals the pair's score
__global__ void wide_flag_kernel(const WideParams P, int64_t total_blocks, WideTask *tasks, uint32_t cap, uint32_t *count)
{
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total_blocks;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = P.n_pairs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.blk_off[mid] <= idx) lo = mid; else hi = mid; }
        const int pair = lo;
        const int64_t rel = idx - P.blk_off[pair];
        const int ro = P.pair_ref[pair];
        const int n = (int)(P.ref_off[ro + 1] - P.ref_off[ro]);
        int nb = (n + WL - 1 + WCB - 1) / WCB; if (nb < 1) nb = 1;
        const int band = (int)(rel / nb), blk = (int)(rel - (int64_t)band * nb);
        const int S = P.scores[(int64_t)ro * P.n_reads + P.pair_read[pair]];
        if (S <= 0) continue;
        const int32_t *tm = P.tmx + idx * WL;
        uint32_t mask = 0;
        for (int l = 0; l < WL; ++l) if (tm[l] == S) mask |= 1u << l;
        if (mask) {
            const uint32_t k = atomicAdd(count, 1u);
            if (k < cap) tasks[k] = WideTask{pair, band, blk, mask};
        }
    }
}

__global__ void __launch_bounds__(128) wide_locate_kernel(const WideParams P, const WideTask *tasks,
                                                           const uint32_t *n_tasks_ptr, uint32_t cap_tasks,
                                                           uint64_t *keys, uint32_t cap, uint32_t *count)
{
    const uint32_t n_tasks = min(*n_tasks_ptr, cap_tasks);
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < n_tasks; task += n_warps) {
        const WideTask T = tasks[task];
        const WCtx C = make_ctx(P, T.pair);
        int rc[KL]; bool all_valid;
        load_rows(C, T.band, lane, rc, all_valid);
        int H[KL], diag;
        load_wide_state(C, T.band, T.block, lane, H, diag);
        const int S = P.scores[(int64_t)P.pair_ref[T.pair] * P.n_reads + P.pair_read[T.pair]];
        const bool mine = (T.lane_mask >> lane) & 1u;
        int cprev = code_prefetch(C, T.block * WCB - 32, lane);
        for (int u0 = 0; u0 < WCB; u0 += 32) {
            const int s0 = T.block * WCB + u0;
            const int tbuf = top_prefetch(C, T.band, s0, lane);
            const int ccur = code_prefetch(C, s0, lane);
            wide_chunk(C, s0, lane, tbuf, cprev, ccur, rc, H, diag,
                       [&](int, int, const int (&Hc)[KL], bool valid, int j) {
                           if (!(valid && mine)) return;
#pragma unroll
                           for (int r = 0; r < KL; ++r) {
                               const int i = T.band * BH + lane * KL + r + 1;
                               if (Hc[r] == S && i <= C.m) {
                                   const uint32_t k = atomicAdd(count, 1u);
                                   if (k < cap) keys[k] = wide_key((uint64_t)T.pair, (uint32_t)i, (uint32_t)j);
                               }
                           }
                       });
            cprev = ccur;
        }
    }
}

// Traceback: one warp per max cell.  Tile of the current (band, block):
//   tile[c][lane][w], c = 0..WCB (c = 0: checkpointed column), w = 0: boundary row above the
//   lane (row band*BH + lane*KL of the matrix), w = 1..KL: the lane's rows.
// BYTE: the tile keeps only the low 8 bits of every score (12 bytes per lane and column instead of 36), which
// quadruples the warps per SM.  That is enough because the walker carries the exact score of its cell (the pair
// maximum minus the moves so far) and a candidate never lies 250 or more below H (tile_trace_ok): equality of the
// low bytes is equality.  Other score sets use the int32 tile.
template <bool BYTE>
__global__ void __launch_bounds__(32) wide_trace_kernel(const WideParams P, const uint64_t *keys, uint32_t n_cells,
                                                         int32_t *beginnings, int32_t *op_lens, uint32_t *ops,
                                                         int64_t ops_stride)
{
    extern __shared__ int32_t wtile[];                       // [WCB + 1][WL][KL + 1] words, or [WCB + 1][WL][3] words of bytes
    const int lane = threadIdx.x;
    constexpr int LW = BYTE ? (KL + 1 + 3) / 4 : KL + 1;      // words per lane and column
    constexpr int CW = WL * LW;
    auto store_col = [&](int c, int top, const int (&Hc)[KL]) {
        int32_t *col = wtile + c * CW + lane * LW;
        if (BYTE) {
#pragma unroll
            for (int w = 0; w < LW; ++w) {
                uint32_t v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * w + e;                       // byte k: 0 = boundary row, 1..KL = the lane's rows
                    v[e] = k == 0 ? (uint32_t)top : (k <= KL ? (uint32_t)Hc[k - 1 < KL && k >= 1 ? k - 1 : 0] : 0u);
                }
                col[w] = (int32_t)__byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
            }
        } else {
            col[0] = top;
#pragma unroll
            for (int r = 0; r < KL; ++r) col[r + 1] = Hc[r];
        }
    };
    for (uint32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const uint64_t key = keys[cell];
        const int pair = (int)wide_key_pair(key);
        int ci = (int)wide_key_i(key), cj = (int)wide_key_j(key);
        const WCtx C = make_ctx(P, pair);
        int hcur = P.scores[(int64_t)P.pair_ref[pair] * P.n_reads + P.pair_read[pair]];
        int beginning = 0;
        int64_t oplen = 0;
        uint32_t opword = 0;
        uint32_t *myops = ops + (int64_t)cell * ops_stride;
        while (hcur > 0) {                                     // warp-uniform: state is broadcast below
            const int band = (ci - 1) / BH;
            const int tl = ((ci - 1) % BH) / KL;
            const int blk = (cj - 1 + tl) / WCB;
            int rc[KL]; bool all_valid;
            load_rows(C, band, lane, rc, all_valid);
            int H[KL], diag;
            load_wide_state(C, band, blk, lane, H, diag);
            store_col(0, diag, H);
            int cprev = code_prefetch(C, blk * WCB - 32, lane);
            for (int u0 = 0; u0 < WCB; u0 += 32) {
                const int s0 = blk * WCB + u0;
                const int tbuf = top_prefetch(C, band, s0, lane);
                const int ccur = code_prefetch(C, s0, lane);
                wide_chunk(C, s0, lane, tbuf, cprev, ccur, rc, H, diag,
                           [&](int u, int top, const int (&Hc)[KL], bool, int) { store_col(u0 + u + 1, top, Hc); });
                cprev = ccur;
            }
            __syncwarp();
            if (lane == 0) {
                while (hcur > 0) {
                    if ((ci - 1) / BH != band) break;                     // left the band upwards
                    const int tc = ((ci - 1) % BH) / KL;
                    const int r = (ci - 1) % KL + 1;                        // 1..KL
                    const int c = cj - (blk * WCB - tc);
                    if (c < 1 || c > WCB) break;
                    int hw, hn, hnw;
                    if (BYTE) {
                        const uint8_t *lt = reinterpret_cast<const uint8_t *>(wtile + c * CW + tc * LW) + r;
                        hw = lt[-CW * 4]; hn = lt[-1]; hnw = lt[-CW * 4 - 1];
                    } else {
                        const int32_t *lt = wtile + c * CW + tc * LW + r;
                        hw = lt[-CW]; hn = lt[-1]; hnw = lt[-CW - 1];
                    }
                    const int sc = (C.read[ci - 1] == C.ref[cj - 1]) ? C.match : C.mismatch;
                    const int mask = BYTE ? 0xff : -1;
                    const bool eq_a = ((hnw + sc - hcur) & mask) == 0, eq_i = ((hn + C.gap - hcur) & mask) == 0,
                               eq_d = ((hw + C.gap - hcur) & mask) == 0;
                    const uint32_t op = P.tie_gt ? (eq_d ? 3u : (eq_i ? 2u : 1u)) : (eq_a ? 1u : (eq_i ? 2u : 3u));
                    beginning = cj;
                    hcur -= (op == 1u) ? sc : C.gap;                      // exact score of the next cell
                    ci -= (op != 3u);
                    cj -= (op != 2u);
                    opword |= op << (2 * (int)(oplen & 15));
                    ++oplen;
                    if ((oplen & 15) == 0) { myops[(oplen >> 4) - 1] = opword; opword = 0; }
                }
            }
            hcur = __shfl_sync(0xffffffffu, hcur, 0);
            ci = __shfl_sync(0xffffffffu, ci, 0);
            cj = __shfl_sync(0xffffffffu, cj, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (oplen & 15) myops[oplen >> 4] = opword;
            beginnings[cell] = beginning;
            op_lens[cell] = (int32_t)oplen;
        }
    }
}

// ---------------------------------------------------------------------------------------
cudaError_t launch_wide_fill(const WideParams &P, const int2 *items, int n_items, uint32_t *ticket, int sm_count,
                             cudaStream_t st)
{
    if (n_items == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    static const int lag = getenv("SWB_WIDE_LAG") ? std::max(1, atoi(getenv("SWB_WIDE_LAG"))) : 1;
    static const int env_ctas = getenv("SWB_WIDE_CTAS_PER_SM") ? atoi(getenv("SWB_WIDE_CTAS_PER_SM")) : 0;
    if (P.n_symbols <= 8) {
        // profile of 4 or 8 codes; with 4 (DNA) seven 4-warp CTAs fit per SM: cfg3's 3,910 band items then run in ONE
        // wave (4,144 warp slots) instead of 1.65 waves on 2,368 slots, whose second wave is 35 % idle
        const int nc = P.n_symbols <= 4 ? 4 : 8;
        const size_t smem = (size_t)4 * nc * KL * WL * sizeof(int32_t);      // 4 warps x codes x BH rows
        const int per_sm = env_ctas > 0 ? env_ctas : (nc == 4 ? 7 : 4);
        const int ctas = (int)std::min<int64_t>(((int64_t)n_items + 3) / 4, (int64_t)sm_count * per_sm);
        static PerDeviceOnce attr;
        if (attr.need()) {
            e = cudaFuncSetAttribute(wide_fill_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)4 * 8 * KL * WL * sizeof(int32_t)));
            if (e != cudaSuccess) return e;
        }
        if (nc == 4) wide_fill_kernel<true, 4><<<ctas, 128, smem, st>>>(P, items, n_items, ticket, 1, lag);
        else         wide_fill_kernel<true, 8><<<ctas, 128, smem, st>>>(P, items, n_items, ticket, 1, lag);
    } else {
        const int per_sm = env_ctas > 0 ? env_ctas : 4;
        const int ctas = (int)std::min<int64_t>(((int64_t)n_items + 3) / 4, (int64_t)sm_count * per_sm);
        wide_fill_kernel<false, 8><<<ctas, 128, 0, st>>>(P, items, n_items, ticket, 1, lag);
    }
    return cudaGetLastError();
}

cudaError_t launch_wide_flag(const WideParams &P, int64_t total_blocks, WideTask *tasks, uint32_t cap, uint32_t *count,
                             cudaStream_t st)
{
    if (total_blocks == 0) return cudaSuccess;
    const int threads = 128;
    const int64_t blocks = std::min<int64_t>((total_blocks + threads - 1) / threads, 1 << 20);
    wide_flag_kernel<<<(unsigned)blocks, threads, 0, st>>>(P, total_blocks, tasks, cap, count);
    return cudaGetLastError();
}

cudaError_t launch_wide_locate(const WideParams &P, const WideTask *tasks, const uint32_t *n_tasks, uint32_t cap_tasks,
                               uint64_t *keys, uint32_t cap, uint32_t *count, int sm_count, cudaStream_t st)
{
    const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(((int64_t)cap_tasks + 3) / 4, (int64_t)sm_count * 8));
    wide_locate_kernel<<<(unsigned)ctas, 128, 0, st>>>(P, tasks, n_tasks, cap_tasks, keys, cap, count);
    return cudaGetLastError();
}

cudaError_t launch_wide_trace(const WideParams &P, const uint64_t *keys, uint32_t n_cells, int32_t *beginnings,
                              int32_t *op_lens, uint32_t *ops, int64_t ops_stride, int sm_count, cudaStream_t st)
{
    if (n_cells == 0) return cudaSuccess;
    const bool bytes = tile_trace_ok(P.match, P.mismatch, P.gap);        // same bound as the short path's byte tiles
    const size_t smem = (size_t)(WCB + 1) * WL * (bytes ? (KL + 1 + 3) / 4 : KL + 1) * sizeof(int32_t);
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(wide_trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((size_t)(WCB + 1) * WL * (KL + 1) * sizeof(int32_t)));
        if (e != cudaSuccess) return e;
    }
    const int per_sm = std::max(1, std::min(16, (int)((220 * 1024) / (smem + 1024))));   // CTAs (warps) per SM that fit in shared memory
    const int64_t ctas = std::min<int64_t>(n_cells, (int64_t)sm_count * per_sm);
    if (bytes) wide_trace_kernel<true><<<(unsigned)ctas, 32, smem, st>>>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride);
    else       wide_trace_kernel<false><<<(unsigned)ctas, 32, smem, st>>>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride);
    return cudaGetLastError();
}

}  // namespace swb
