// swb_trace.cu -- maximum-cell enumeration and traceback of the short-read path.
//
// The fill (swb_fill.cu) leaves, per pair, the maximum score, per-(lane, block) tile
// maxima and per-block register checkpoints.  Here:
//   flag_tiles : tiles whose maximum equals the pair's maximum (and is > 0)
//   locate     : recompute each flagged block from its checkpoint, emit every cell == max
//                -> the reference's max-cell list (ScoreMatrix.call, SmithWaterman.java:176-185);
//                keys (pair, i, j) are radix-sorted, which IS the row-major list order
//   trace      : per max cell, GetAlignment.call (SmithWaterman.java:354-436): walk while
//                the score is positive; the type of a positive cell is re-derived from the
//                scores with the priority of the ">=" cascade (:228,:236,:245): alignment,
//                then insertion, then deletion.  Blocks are recomputed lazily, right to
//                left, each from its own checkpoint, so every H the walk reads is exact.
// The recompute is unpacked int32 (DPX s32 ops); one 8-lane group per task, same lane/row/
// step geometry as the fill so the checkpoints drop straight into registers.
#include "swb_internal.h"
#include "swb_device.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace swb {

// ---------------------------------------------------------------------------------------
__global__ void flag_tiles_kernel(const BatchParams P, TileTask *tasks, uint32_t cap, uint32_t *count)
{
    // one thread per (read pair, sorted ref, block)
    const int64_t total = (int64_t)P.n_rp * P.blocks_per_rp;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int rp = (int)(idx / P.blocks_per_rp);
        const int64_t gb = idx - (int64_t)rp * P.blocks_per_rp;
        // sorted ref that owns global block gb: binary search in ref_blk_off
        int lo = 0, hi = P.n_refs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.ref_blk_off[mid] <= gb) lo = mid; else hi = mid; }
        const int ref = lo;
        const int b = (int)(gb - P.ref_blk_off[ref]);
        const int64_t ro = P.ref_orig[ref];
        const int ra = P.rp_reads[2 * rp], rb = P.rp_reads[2 * rp + 1];
        const int sa = P.scores[ro * P.n_reads + ra];
        const int sb = rb >= 0 ? P.scores[ro * P.n_reads + rb] : 0;
        if (sa <= 0 && sb <= 0) continue;
        const uint4 *tm = reinterpret_cast<const uint4 *>(P.tmx + idx * GL);
        const uint4 v0 = tm[0], v1 = tm[1];
        const uint32_t w[GL] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t ma = 0, mb = 0;
#pragma unroll
        for (int t = 0; t < GL; ++t) {
            if (sa > 0 && half_of(w[t], 0) == sa) ma |= 1u << t;
            if (sb > 0 && half_of(w[t], 1) == sb) mb |= 1u << t;
        }
        if (ma) {
            const uint32_t k = atomicAdd(count, 1u);
            if (k < cap) tasks[k] = TileTask{(uint32_t)rp * 2u, (uint32_t)ref, (uint32_t)b, ma};
        }
        if (mb) {
            const uint32_t k = atomicAdd(count, 1u);
            if (k < cap) tasks[k] = TileTask{(uint32_t)rp * 2u + 1u, (uint32_t)ref, (uint32_t)b, mb};
        }
    }
}

cudaError_t launch_flag_tiles(const BatchParams &P, TileTask *tasks, uint32_t cap, uint32_t *count, cudaStream_t st)
{
    const int64_t total = (int64_t)P.n_rp * P.blocks_per_rp;
    const int threads = 256;
    const int64_t blocks = std::min<int64_t>((total + threads - 1) / threads, 1 << 20);
    if (blocks == 0) return cudaSuccess;
    flag_tiles_kernel<<<(unsigned)blocks, threads, 0, st>>>(P, tasks, cap, count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// One block (CB steps) of the wavefront for one group, unpacked int32.
// State in/out: H[K] (this lane's column), diag.  sink(u, top, H, valid, j) after each step.
template <int K>
struct GroupCtx {
    const uint32_t *ref_words;   // this ref's packed codes
    int n;                       // ref length
    int m;                       // read length
    int rc[K];                   // read codes of this lane's rows (0xFE beyond the read)
    int match, mismatch, gap;
};

template <int K>
__device__ __forceinline__ void load_state(const BatchParams &P, int64_t blk0, int b, int half, int t,
                                           int (&H)[K], int &diag)
{
    if (b == 0) {
#pragma unroll
        for (int r = 0; r < K; ++r) H[r] = 0;
        diag = 0;
        return;
    }
    const uint32_t *blk = P.ck + (blk0 + b) * (int64_t)(Geo<K>::KW * GL);
#pragma unroll
    for (int r = 0; r < K; ++r) H[r] = half_of(load_checkpoint_word<K>(blk, t, r), half);
    diag = half_of(load_checkpoint_word<K>(blk, t, K), half);
}

template <int K, class Sink>
__device__ __forceinline__ void run_block(const GroupCtx<K> &C, int b, int t, unsigned gmask,
                                          int (&H)[K], int &diag, Sink &&sink)
{
#pragma unroll 1
    for (int u = 0; u < CB; ++u) {
        const int s = b * CB + u;
        int top = __shfl_up_sync(gmask, H[K - 1], 1, GL);
        if (t == 0) top = 0;
        const int j = s - t + 1;
        const bool valid = (j >= 1) && (j <= C.n);
        if (valid) {
            const int col = j - 1;
            const int c = (int)((__ldg(C.ref_words + (col >> 4)) >> (2 * (col & 15))) & 3u);
            int nw = diag, nn = top;
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const int sc = (C.rc[r] == c) ? C.match : C.mismatch;
                const int pre = __viaddmax_s32_relu(H[r], C.gap, nw + sc);   // max(W+gap, NW+s, 0)
                nw = H[r];
                H[r] = __viaddmax_s32(nn, C.gap, pre);                        // max(N+gap, pre)
                nn = H[r];
            }
        }
        sink(u, top, H, valid, j);
        diag = top;
    }
}

template <int K>
__device__ __forceinline__ void init_group(const BatchParams &P, int rp, int half, int ref, int t, GroupCtx<K> &C,
                                           int64_t &blk0, int64_t &pair, int &read_idx)
{
    read_idx = P.rp_reads[2 * rp + half];
    const int64_t off = P.read_off[read_idx];
    C.m = (int)(P.read_off[read_idx + 1] - off);
    C.n = P.ref_len[ref];
    C.ref_words = P.ref_words + P.ref_word_off[ref];
    C.match = P.match; C.mismatch = P.mismatch; C.gap = P.gap;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const int row = t * K + r;
        C.rc[r] = (row < C.m) ? (int)P.read_codes[off + row] : 0xFE;
    }
    blk0 = (int64_t)rp * P.blocks_per_rp + P.ref_blk_off[ref];
    pair = (int64_t)P.ref_orig[ref] * P.n_reads + read_idx;
}

// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) locate_kernel(const BatchParams P, const TileTask *tasks,
                                                      const uint32_t *n_tasks_ptr, uint32_t cap_tasks,
                                                      uint64_t *keys, uint32_t cap, uint32_t *count)
{
    const uint32_t n_tasks = min(*n_tasks_ptr, cap_tasks);        // written by flag_tiles on the same stream
    const int lane = threadIdx.x & 31, t = lane & (GL - 1), g = lane >> 3;
    const unsigned gmask = 0xffu << (8 * g);
    const uint32_t n_groups = gridDim.x * (blockDim.x >> 3);
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    // all 4 groups of a warp iterate the same number of times (tasks padded with idle turns)
    const uint32_t iters = (n_tasks + n_groups - 1) / n_groups;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t task = it * n_groups + gid;
        const bool live = task < n_tasks;
        TileTask T = live ? tasks[task] : TileTask{0, 0, 0, 0};
        GroupCtx<K> C;
        int64_t blk0 = 0, pair = 0; int read_idx = 0;
        int H[K], diag = 0;
        int S = 0;
        uint64_t pkey = 0;
        if (live) {
            init_group<K>(P, (int)(T.rp_half >> 1), (int)(T.rp_half & 1), (int)T.ref_sorted, t, C, blk0, pair, read_idx);
            load_state<K>(P, blk0, (int)T.block, (int)(T.rp_half & 1), t, H, diag);
            S = P.scores[pair];
            pkey = (uint64_t)T.rp_half * (uint64_t)P.n_refs + (uint64_t)P.ref_orig[T.ref_sorted];
        } else {
            C.n = 0; C.m = 0; C.ref_words = P.ref_words; C.match = C.mismatch = C.gap = 0;
#pragma unroll
            for (int r = 0; r < K; ++r) { H[r] = 0; C.rc[r] = 0xFE; }
        }
        const bool mine = live && ((T.lane_mask >> t) & 1u);
        run_block<K>(C, (int)T.block, t, gmask, H, diag,
                     [&](int, int, const int (&Hc)[K], bool valid, int j) {
                         if (!(valid && mine)) return;
#pragma unroll
                         for (int r = 0; r < K; ++r) {
                             const int i = t * K + r + 1;
                             if (Hc[r] == S && i <= C.m) {
                                 const uint32_t k = atomicAdd(count, 1u);
                                 if (k < cap) keys[k] = make_key(pkey, (uint32_t)i, (uint32_t)j);
                             }
                         }
                     });
    }
}

template <int K>
static cudaError_t launch_locate_k(const BatchParams &P, const TileTask *tasks, const uint32_t *n_tasks,
                                   uint32_t cap_tasks, uint64_t *keys, uint32_t cap, uint32_t *count, int sm_count,
                                   cudaStream_t st)
{
    const int threads = 128;                      // 16 groups per CTA; the task count is read on the device
    int64_t ctas = std::min<int64_t>(((int64_t)cap_tasks + 15) / 16, (int64_t)sm_count * 16);
    locate_kernel<K><<<(unsigned)std::max<int64_t>(ctas, 1), threads, 0, st>>>(P, tasks, n_tasks, cap_tasks, keys, cap, count);
    return cudaGetLastError();
}

cudaError_t launch_locate(int K, const BatchParams &P, const TileTask *tasks, const uint32_t *n_tasks,
                          uint32_t cap_tasks, uint64_t *keys, uint32_t cap, uint32_t *count, int sm_count,
                          cudaStream_t st)
{
    switch (K) {
        case 4:  return launch_locate_k<4>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 8:  return launch_locate_k<8>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 13: return launch_locate_k<13>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 16: return launch_locate_k<16>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 19: return launch_locate_k<19>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 25: return launch_locate_k<25>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
        case 32: return launch_locate_k<32>(P, tasks, n_tasks, cap_tasks, keys, cap, count, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

template <int K> struct TraceGeo {
    // lanes of a half's tile: a walk of CB steps climbs at most CB rows (plus rare insertion runs,
    // which simply end the block early), i.e. ceil(CB / K) lanes above the current one
    static constexpr int NLW = ((CB + K - 1) / K + 1) < GL ? ((CB + K - 1) / K + 1) : GL;
    static constexpr int KW = Geo<K>::KW;
    static constexpr int HALF_WORDS = NLW * (CB + 1) * KW;            // one half's tile: [slot][column][KW packed words]
    static constexpr int CODE_WORDS = ((2 * NLW * (CB + 1) + 1 + 15) / 16) * 4;   // [half][slot][column] bytes + 1 trash byte
    static constexpr int GROUP_WORDS = 2 * HALF_WORDS + CODE_WORDS + KW;         // + one trash column
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one tile column: word 0 = boundary row received from the lane above, words 1..K = this lane's rows
template <int K>
__device__ __forceinline__ void store_col(uint32_t *dst, uint32_t top, const uint32_t (&H)[K])
{
    constexpr int KW = Geo<K>::KW;
#pragma unroll
    for (int q = 0; q < KW / 4; ++q) {
        uint4 v;
        v.x = (4 * q == 0) ? top : ((4 * q - 1 < K) ? H[(4 * q - 1 < K && 4 * q >= 1) ? 4 * q - 1 : 0] : 0u);
        v.y = (4 * q + 0 < K) ? H[(4 * q + 0 < K) ? 4 * q + 0 : 0] : 0u;
        v.z = (4 * q + 1 < K) ? H[(4 * q + 1 < K) ? 4 * q + 1 : 0] : 0u;
        v.w = (4 * q + 2 < K) ? H[(4 * q + 2 < K) ? 4 * q + 2 : 0] : 0u;
        reinterpret_cast<uint4 *>(dst)[q] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Traceback, packed: one 8-lane group walks TWO max cells at once, one per s16 half.
// Keys are sorted (read slot, ref, i, j), so a run of cells shares its read: the two halves
// align the same read rows against two (usually different) references, and the score
// profile depends on the read only:
//     prof2[cA*5 + cB][lane][row] = pack(s(read[row], cA), s(read[row], cB)),  code 4 = "no column"
// (S_PAD: a column left of the matrix stays all-zero, one right of it is never read).
// Each half restarts from ITS block's checkpoint (own ref, own block index); the group then
// runs the same three-op cell as the (unbiased) fill for CB steps.  Only the lanes a walk of CB
// steps can reach are kept, [tc - NLW + 1, tc] around the current cell's lane tc, per half:
//     tile[half][slot][c][w]   c = 0..CB (c = 0: the checkpointed column), w = 0: boundary row
//                              received from lane-1 (matrix row lane*K), w = 1..K: the lane's rows
// so W, N and NW of a cell with c >= 1 lie in the same slot.  Lanes outside a window store to a
// trash column (no branches in the step loop).  Lane h of the group walks half h
// (GetAlignment.call, SmithWaterman.java:380-409) until the path leaves what the block holds, then
// the block that holds the current cell is recomputed -- exact, because every block starts
// from the fill's own registers.
template <int K>
__global__ void __launch_bounds__(128) trace_kernel(const BatchParams P, const uint64_t *keys, uint32_t n_cells,
                                                     int chunk, int32_t *beginnings, int32_t *op_lens, uint32_t *ops,
                                                     int ops_stride, uint32_t zero)
{
    using G = Geo<K>;
    using TG = TraceGeo<K>;
    constexpr int KW = G::KW;
    constexpr int NLW = TG::NLW;
    constexpr int HALF_WORDS = TG::HALF_WORDS;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t seg_next;
    uint32_t *prof2 = smem;                                         // [25][GL][KS]
    const int lane = threadIdx.x & 31, t = lane & (GL - 1), g = lane >> 3, warp = threadIdx.x >> 5;
    uint32_t *gbase = smem + 25 * G::CSTRIDE + (size_t)(warp * 4 + g) * TG::GROUP_WORDS;
    uint32_t *tile = gbase;
    uint8_t *codes = reinterpret_cast<uint8_t *>(gbase + 2 * HALF_WORDS);   // [half][slot][CB + 1]
    uint8_t *trash_code = codes + 2 * NLW * (CB + 1);
    uint32_t *trash = gbase + 2 * HALF_WORDS + TG::CODE_WORDS;
    uint8_t *rcodes_s = reinterpret_cast<uint8_t *>(smem + 25 * G::CSTRIDE + (size_t)(blockDim.x >> 3) * TG::GROUP_WORDS);   // [GL*K]
    const uint32_t g2 = pack2(P.gap, P.gap);
    const int my_prof = t * G::KS;

    const uint32_t c_lo = blockIdx.x * (uint32_t)chunk;
    const uint32_t c_hi = min(n_cells, c_lo + (uint32_t)chunk);
    uint32_t seg_lo = c_lo;
    while (seg_lo < c_hi) {
        // ---- segment = run of cells of one read (slot) -----------------------------------
        const uint32_t slot = (uint32_t)(key_pair(keys[seg_lo]) / (uint64_t)P.n_refs);
        uint32_t seg_hi;
        {
            const uint64_t target = make_key((uint64_t)(slot + 1) * (uint64_t)P.n_refs, 0, 0);
            uint32_t lo = seg_lo, hi = c_hi;
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < target) lo = mid + 1; else hi = mid; }
            seg_hi = lo;
        }
        const int read_idx = P.rp_reads[slot];
        const int rh = (int)(slot & 1u);                             // the read's half in the fill's checkpoints
        const int64_t roff = P.read_off[read_idx];
        const int m = (int)(P.read_off[read_idx + 1] - roff);
        __syncthreads();
        for (int idx = threadIdx.x; idx < 25 * GL * K; idx += blockDim.x) {
            const int cc = idx / (GL * K), rem = idx - cc * GL * K;
            const int tt = rem / K, r = rem - tt * K;
            const int row = tt * K + r;
            const int ca = cc / 5, cb = cc - ca * 5;
            int lo = S_PAD, hi = S_PAD;
            if (row < m) {
                const int q = P.read_codes[roff + row];
                if (ca < 4) lo = (q == ca) ? P.match : P.mismatch;
                if (cb < 4) hi = (q == cb) ? P.match : P.mismatch;
            }
            prof2[cc * G::CSTRIDE + tt * G::KS + r] = pack2(lo, hi);
        }
        for (int idx = threadIdx.x; idx < GL * K; idx += blockDim.x)
            rcodes_s[idx] = idx < m ? P.read_codes[roff + idx] : (uint8_t)0xFE;
        if (threadIdx.x == 0) seg_next = seg_lo;
        __syncthreads();

        // ---- per-half state, replicated in the 8 lanes of the group ----------------------
        bool busy0 = false, busy1 = false;
        int ci0 = 0, cj0 = 0, ci1 = 0, cj1 = 0, n0 = 0, n1 = 0;
        const uint32_t *rw0 = P.ref_words, *rw1 = P.ref_words;
        int64_t bk0 = 0, bk1 = 0;
        // walker state, replicated in the 4 lanes of a half's walk team (lanes 0-3: half 0, 4-7: half 1);
        // the team leader (lane 0 / 4) writes the results
        uint32_t w_cell = 0, w_opword = 0;
        int w_h = 0, w_beg = 0, w_len = 0;

        for (;;) {
            // fetch a new cell for every idle half
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool idle = h ? !busy1 : !busy0;
                uint32_t e = 0xffffffffu;
                if (idle && t == 4 * h) e = atomicAdd(&seg_next, 1u);
                e = __shfl_sync(FULL, e, 4 * h, GL);
                if (idle && e < seg_hi) {
                    const uint64_t key = keys[e];
                    const int ro = (int)(key_pair(key) - (uint64_t)slot * (uint64_t)P.n_refs);
                    const int ref = P.ref_sorted_of[ro];
                    const int ii = (int)key_i(key), jj = (int)key_j(key);
                    if (h == 0) { busy0 = true; ci0 = ii; cj0 = jj; n0 = P.ref_len[ref]; rw0 = P.ref_words + P.ref_word_off[ref];
                                  bk0 = (int64_t)(slot >> 1) * P.blocks_per_rp + P.ref_blk_off[ref]; }
                    else        { busy1 = true; ci1 = ii; cj1 = jj; n1 = P.ref_len[ref]; rw1 = P.ref_words + P.ref_word_off[ref];
                                  bk1 = (int64_t)(slot >> 1) * P.blocks_per_rp + P.ref_blk_off[ref]; }
                    if ((t >> 2) == h) {                             // the 4 lanes of this half's walk team
                        w_cell = e; w_opword = 0; w_beg = 0; w_len = 0;
                        w_h = P.scores[(int64_t)ro * P.n_reads + read_idx];
                    }
                }
            }
            if (!__any_sync(FULL, busy0 || busy1)) break;

            // ---- block of each half, checkpoint -> registers ---------------------------------
            const int tca = (ci0 - 1) / K, tcb = (ci1 - 1) / K;
            const int b0 = busy0 ? (cj0 - 1 + tca) / CB : 0;
            const int b1 = busy1 ? (cj1 - 1 + tcb) / CB : 0;
            // this lane's slot in each half's window [tc - NLW + 1, tc]; outside: the trash column
            const int sa = busy0 ? t - (tca - NLW + 1) : -1;
            const int sb = busy1 ? t - (tcb - NLW + 1) : -1;
            const bool ina = sa >= 0 && sa < NLW, inb = sb >= 0 && sb < NLW;
            // ONE packed column store per lane and step (shared-memory bandwidth is this kernel's tightest
            // resource): physical slots 0..NLW-1 hold half 0's window, NLW..2NLW-1 the lanes that are only in
            // half 1's window; a lane in both windows lives in half 0's slot and half 1's walker looks it up there
            const int phys = ina ? sa : (inb ? NLW + sb : -1);
            uint32_t *pa = phys >= 0 ? tile + (size_t)phys * (CB + 1) * KW : trash;
            uint8_t *qa = ina ? codes + sa * (CB + 1) : trash_code;
            uint8_t *qb = inb ? codes + (NLW + sb) * (CB + 1) : trash_code;
            const int inca = phys >= 0 ? KW : 0, incqa = ina ? 1 : 0, incqb = inb ? 1 : 0;

            uint32_t H[K], diag = 0;
            {
                const uint4 *ckA = reinterpret_cast<const uint4 *>(P.ck + (bk0 + b0) * (int64_t)(KW * GL)) + t;
                const uint4 *ckB = reinterpret_cast<const uint4 *>(P.ck + (bk1 + b1) * (int64_t)(KW * GL)) + t;
                const bool la = busy0 && b0 > 0, lb = busy1 && b1 > 0;
                const uint32_t sel = rh ? 0x7632u : 0x5410u;        // (A.half rh) | (B.half rh) << 16
                const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < KW / 4; ++q) {
                    const uint4 a = la ? __ldg(ckA + q * GL) : z;
                    const uint4 bq = lb ? __ldg(ckB + q * GL) : z;
                    const uint32_t v[4] = {__byte_perm(a.x, bq.x, sel), __byte_perm(a.y, bq.y, sel),
                                           __byte_perm(a.z, bq.z, sel), __byte_perm(a.w, bq.w, sel)};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int w = 4 * q + e;
                        if (w < K) H[w < K ? w : 0] = v[e];
                        else if (w == K) diag = v[e];
                    }
                }
            }
            // the path almost always continues into the block to the left: pull its checkpoints towards L2
            if (t < KW / 4) {
                if (busy0 && b0 > 1) prefetch_l2(P.ck + (bk0 + b0 - 1) * (int64_t)(KW * GL) + t * (GL * 4));
                if (busy1 && b1 > 1) prefetch_l2(P.ck + (bk1 + b1 - 1) * (int64_t)(KW * GL) + t * (GL * 4));
            }
            // reference-code windows of this lane for the block: 0-based columns j0 .. j0+15
            uint32_t win0, win1; int ulo0, uhi0, ulo1, uhi1;
            {
                const int j0 = b0 * CB - t, j1 = b1 * CB - t;
                const int wi0 = j0 >> 4, wi1 = j1 >> 4;                   // floor (j may be negative in block 0)
                const uint32_t a0 = (busy0 && wi0 >= 0 && wi0 * 16 < n0) ? __ldg(rw0 + wi0) : 0u;
                const uint32_t a1 = (busy0 && wi0 + 1 >= 0 && (wi0 + 1) * 16 < n0) ? __ldg(rw0 + wi0 + 1) : 0u;
                const uint32_t c0 = (busy1 && wi1 >= 0 && wi1 * 16 < n1) ? __ldg(rw1 + wi1) : 0u;
                const uint32_t c1 = (busy1 && wi1 + 1 >= 0 && (wi1 + 1) * 16 < n1) ? __ldg(rw1 + wi1 + 1) : 0u;
                win0 = __funnelshift_r(a0, a1, 2 * (j0 & 15));
                win1 = __funnelshift_r(c0, c1, 2 * (j1 & 15));
                // step u computes 0-based column j0 + u: valid iff 0 <= j0 + u < n
                ulo0 = -j0; uhi0 = busy0 ? n0 - j0 - 1 : -1;
                ulo1 = -j1; uhi1 = busy1 ? n1 - j1 - 1 : -1;
            }
            // halo column c = 0
            store_col<K>(pa, diag, H);
            pa += inca; qa += incqa; qb += incqb;

            // interior block: every step of every lane / half is a real column -> no per-step validity tests
            const bool interior = ulo0 <= 0 && uhi0 >= CB - 1 && ulo1 <= 0 && uhi1 >= CB - 1;
            if (__all_sync(FULL, interior)) {
#pragma unroll 4
                for (int u = 0; u < CB; ++u) {
                    uint32_t top = __shfl_up_sync(FULL, H[K - 1], 1, GL);
                    if (t == 0) top = 0;
                    const uint32_t ca = (win0 >> (2 * u)) & 3u, cb = (win1 >> (2 * u)) & 3u;
                    uint32_t sv[G::KP];
                    load_profile<G::KP>(prof2 + (ca * 5 + cb) * G::CSTRIDE + my_prof, sv);
                    uint32_t nw = diag, nn = top;
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const uint32_t x = viaddmax_relu(nw, sv[r], zero);
                        const uint32_t pre = viaddmax(H[r], g2, x);
                        nw = H[r];
                        H[r] = viaddmax(nn, g2, pre);
                        nn = H[r];
                    }
                    diag = top;
                    store_col<K>(pa, top, H);
                    *qa = (uint8_t)ca; *qb = (uint8_t)cb;
                    pa += inca; qa += incqa; qb += incqb;
                }
            } else {
#pragma unroll 1
                for (int u = 0; u < CB; ++u) {
                    uint32_t top = __shfl_up_sync(FULL, H[K - 1], 1, GL);
                    if (t == 0) top = 0;
                    const uint32_t ca = (u >= ulo0 && u <= uhi0) ? ((win0 >> (2 * u)) & 3u) : 4u;
                    const uint32_t cb = (u >= ulo1 && u <= uhi1) ? ((win1 >> (2 * u)) & 3u) : 4u;
                    uint32_t sv[G::KP];
                    load_profile<G::KP>(prof2 + (ca * 5 + cb) * G::CSTRIDE + my_prof, sv);
                    uint32_t nw = diag, nn = top;
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        const uint32_t x = viaddmax_relu(nw, sv[r], zero);
                        const uint32_t pre = viaddmax(H[r], g2, x);
                        nw = H[r];
                        H[r] = viaddmax(nn, g2, pre);
                        nn = H[r];
                    }
                    diag = top;
                    store_col<K>(pa, top, H);
                    *qa = (uint8_t)ca; *qb = (uint8_t)cb;
                    pa += inca; qa += incqa; qb += incqb;
                }
            }
            __syncwarp();

            // ---- walk (GetAlignment.call, SmithWaterman.java:380-409), four lanes per half ----------------
            // Lane q of a team looks at the cell q steps up the diagonal from the current one and derives ITS
            // move from the tile.  Alignment paths are mostly diagonal, so a ballot finds the run of leading
            // "alignment" moves and the whole run (plus the first gap move after it) is applied in one round:
            // ~3 path steps per round instead of one load -> compare -> select chain per step.
            const int h = t >> 2, q = t & 3;
            const bool mybusy = h ? busy1 : busy0;
            int ci = h ? ci1 : ci0, cj = h ? cj1 : cj0;
            const int b = h ? b1 : b0;
            const int lane_lo = (h ? tcb : tca) - NLW + 1;              // first lane held in this half's window
            const int a_lo = tca - NLW + 1;
            const int16_t *T16 = reinterpret_cast<const int16_t *>(tile);   // halfword 2*w + h of packed word w
            const uint8_t *tcodes = codes + h * NLW * (CB + 1);
            auto phys_of = [&](int s_) -> int {                           // physical slot of window position s_
                if (h == 0) return s_;
                const int in_a = (lane_lo + s_) - a_lo;
                return (busy0 && in_a >= 0 && in_a < NLW) ? in_a : NLW + s_;
            };
            // position of the current cell inside the tile, kept incrementally
            const int tc0 = (ci - 1) / K;
            int r = ci - tc0 * K;                                        // 1..K
            int c = cj - (b * CB - tc0);
            int sl = tc0 - lane_lo;
            bool walking = mybusy;                                        // this half still walks in this block
            int done = 0;
            const int team = lane & ~3;
            while (__any_sync(FULL, walking)) {
                // my cell: q steps up the diagonal
                int r_q = r - q, sl_q = sl, c_q = c - q;
                if (r_q <= 0) { r_q += K; --sl_q; --c_q; }
                const int ci_q = ci - q, cj_q = cj - q;
                const bool in = walking && ci_q >= 1 && cj_q >= 1 && c_q >= 1 && sl_q >= 0;
                int hc = 0, hw = 0, hn = 0, hnw = 0, rc = 0, qc = 1;
                if (in) {
                    const int idx = 2 * ((phys_of(sl_q) * (CB + 1) + c_q) * KW + r_q) + h;
                    hc = T16[idx]; hw = T16[idx - 2 * KW]; hn = T16[idx - 2]; hnw = T16[idx - 2 * KW - 2];
                    rc = tcodes[sl_q * (CB + 1) + c_q];
                    qc = rcodes_s[ci_q - 1];
                }
                const int sc = (qc == rc) ? P.match : P.mismatch;
                // type of a positive cell = first of (alignment, insertion, deletion) whose candidate equals H:
                // the ">=" cascade of GetCellScore.call; SWB_F_TIE_GT: first of (deletion, insertion, alignment)
                const bool eq_a = (hnw + sc == hc), eq_i = (hn + P.gap == hc), eq_d = (hw + P.gap == hc);
                const uint32_t op = P.tie_gt ? (eq_d ? 3u : (eq_i ? 2u : 1u)) : (eq_a ? 1u : (eq_i ? 2u : 3u));
                const bool alive = in && hc > 0;
                const int hnext = op == 1u ? hnw : (op == 2u ? hn : hw);
                const uint32_t bal = (__ballot_sync(FULL, alive && op == 1u) >> team) & 0xFu;
                const int n_a = __ffs((int)(~bal & 0x1Fu)) - 1;         // leading alignment moves: 0..4
                const int src = team + min(n_a, 3);
                const bool alive_s = __shfl_sync(FULL, (int)alive, src) != 0;
                const uint32_t op_s = __shfl_sync(FULL, op, src);
                const int hnext_s = __shfl_sync(FULL, hnext, src);
                const int hnw_prev = __shfl_sync(FULL, hnw, team + max(n_a - 1, 0));
                const bool extra = n_a < 4 && alive_s;                   // lane n_a's own (gap) move is applied too
                const int steps = n_a + (extra ? 1 : 0);
                if (walking) {
                    if (steps == 0) {
                        // the current cell is outside what this block holds (or its score is 0: cannot happen here)
                        walking = false;
                    } else {
                        const int up = extra ? (op_s != 3u) : 0, left = extra ? (op_s != 2u) : 0;
                        w_beg = extra ? cj - n_a : cj - n_a + 1;        // column of the last cell that made a move
                        const uint64_t bits = (n_a ? (0x55ull >> (8 - 2 * n_a)) : 0ull) | (extra ? (uint64_t)op_s << (2 * n_a) : 0ull);
                        const uint64_t acc = (uint64_t)w_opword | (bits << (2 * (w_len & 15)));
                        const int nl = w_len + steps;
                        if ((nl >> 4) != (w_len >> 4)) {                 // a 16-column word is complete
                            if (q == 0) ops[(size_t)w_cell * ops_stride + (w_len >> 4)] = (uint32_t)acc;
                            w_opword = (uint32_t)(acc >> 32);
                        } else {
                            w_opword = (uint32_t)acc;
                        }
                        w_len = nl;
                        w_h = extra ? hnext_s : ((n_a == 4) ? hnext_s : hnw_prev);
                        ci -= n_a + up; cj -= n_a + left;
                        r -= n_a + up; c -= n_a + left;
                        if (r <= 0) { r += K; --sl; --c; }
                        if (r <= 0) { r += K; --sl; --c; }
                        if (w_h <= 0) { walking = false; done = 1; }
                    }
                }
            }
            if (done && q == 0) {
                if (w_len & 15) ops[(size_t)w_cell * ops_stride + (w_len >> 4)] = w_opword;
                beginnings[w_cell] = w_beg;
                op_lens[w_cell] = w_len;
            }
            {
                const int d0 = __shfl_sync(FULL, done, 0, GL), d1 = __shfl_sync(FULL, done, 4, GL);
                const int i0 = __shfl_sync(FULL, ci, 0, GL), j0 = __shfl_sync(FULL, cj, 0, GL);
                const int i1 = __shfl_sync(FULL, ci, 4, GL), j1 = __shfl_sync(FULL, cj, 4, GL);
                if (busy0) { ci0 = i0; cj0 = j0; if (d0) busy0 = false; }
                if (busy1) { ci1 = i1; cj1 = j1; if (d1) busy1 = false; }
            }
            __syncwarp();
        }
        seg_lo = seg_hi;
    }
}

template <int K>
static cudaError_t launch_trace_k(const BatchParams &P, const uint64_t *keys, uint32_t n_cells, int32_t *beginnings,
                                  int32_t *op_lens, uint32_t *ops, int ops_stride, int sm_count, cudaStream_t st)
{
    using G = Geo<K>;
    if (n_cells == 0) return cudaSuccess;
    const size_t per_group = (size_t)TraceGeo<K>::GROUP_WORDS * sizeof(uint32_t);
    const size_t prof_bytes = (size_t)25 * G::CSTRIDE * sizeof(uint32_t);
    const size_t tail_bytes = (((size_t)GL * K + 15) / 16) * 16;
    int warps = 4;
    while (warps > 1 && prof_bytes + per_group * 4 * warps + tail_bytes > 112 * 1024) --warps;   // two CTAs per SM
    const size_t smem = prof_bytes + per_group * 4 * warps + tail_bytes;
    static bool attr_set[64] = {false};
    if (!attr_set[K]) {
        cudaError_t e = cudaFuncSetAttribute(trace_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[K] = true;
    }
    // chunks of the sorted cell list: several per SM for balance, each long enough to amortise its profile
    int chunk = (int)std::max<int64_t>(256, ((int64_t)n_cells + sm_count * 4 - 1) / (sm_count * 4));
    const int64_t ctas = ((int64_t)n_cells + chunk - 1) / chunk;
    trace_kernel<K><<<(unsigned)ctas, warps * 32, smem, st>>>(P, keys, n_cells, chunk, beginnings, op_lens, ops, ops_stride, 0u);
    return cudaGetLastError();
}

cudaError_t launch_trace(int K, const BatchParams &P, const uint64_t *keys, uint32_t n_cells, int32_t *beginnings,
                         int32_t *op_lens, uint32_t *ops, int ops_stride_words, int sm_count, cudaStream_t st)
{
    switch (K) {
        case 4:  return launch_trace_k<4>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 8:  return launch_trace_k<8>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 13: return launch_trace_k<13>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 16: return launch_trace_k<16>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 19: return launch_trace_k<19>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 25: return launch_trace_k<25>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
        case 32: return launch_trace_k<32>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride_words, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------
// offsets[k] = first sorted key whose pair id >= pair_ids[k]  (pair_ids ascending; last = +inf)
__global__ void cell_offsets_kernel(const uint64_t *keys, uint32_t n_cells, const int64_t *pair_ids, int64_t n_pairs,
                                    int64_t *offsets)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n_pairs) return;
    if (k == n_pairs) { offsets[k] = n_cells; return; }
    const uint64_t target = make_key((uint64_t)pair_ids[k], 0, 0);
    uint32_t lo = 0, hi = n_cells;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    offsets[k] = lo;
}

cudaError_t launch_cell_offsets(const uint64_t *keys, uint32_t n_cells, const int64_t *pair_ids, int64_t n_pairs,
                                int64_t *offsets, cudaStream_t st)
{
    const int threads = 256;
    const int64_t blocks = (n_pairs + 1 + threads - 1) / threads;
    cell_offsets_kernel<<<(unsigned)blocks, threads, 0, st>>>(keys, n_cells, pair_ids, n_pairs, offsets);
    return cudaGetLastError();
}

// per-reference wrapping int32 total over reads (Distribution.java:424): one warp per ref
__global__ void ref_totals_kernel(const int32_t *scores, int64_t n_refs, int64_t n_reads, int32_t *totals)
{
    const int64_t ref = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ref >= n_refs) return;
    uint32_t acc = 0;
    for (int64_t q = lane; q < n_reads; q += 32) acc += (uint32_t)scores[ref * n_reads + q];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) totals[ref] = (int32_t)acc;
}

cudaError_t launch_ref_totals(const int32_t *scores, int64_t n_refs, int64_t n_reads, int32_t *totals, cudaStream_t st)
{
    if (n_refs == 0) return cudaSuccess;
    const int threads = 256;
    const int64_t blocks = (n_refs * 32 + threads - 1) / threads;
    ref_totals_kernel<<<(unsigned)blocks, threads, 0, st>>>(scores, n_refs, n_reads, totals);
    return cudaGetLastError();
}

// per-read best reference: highest score, lowest ref index on ties. best[4q] = score, [4q+1] = ref
__global__ void best_hits_kernel(const int32_t *scores, int64_t n_refs, int64_t n_reads, int32_t *best)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_reads) return;
    int bs = -1, br = -1;
    for (int64_t r = 0; r < n_refs; ++r) {
        const int s = scores[r * n_reads + q];
        if (s > bs) { bs = s; br = (int)r; }
    }
    best[4 * q] = bs < 0 ? 0 : bs; best[4 * q + 1] = br; best[4 * q + 2] = 0; best[4 * q + 3] = 0;
}

cudaError_t launch_best_hits(const int32_t *scores, int64_t n_refs, int64_t n_reads, int32_t *best, cudaStream_t st)
{
    if (n_reads == 0) return cudaSuccess;
    const int threads = 128;
    best_hits_kernel<<<(unsigned)((n_reads + threads - 1) / threads), threads, 0, st>>>(scores, n_refs, n_reads, best);
    return cudaGetLastError();
}

// fill (i, j) of each read's best hit: first key of pair (best ref, read) in the read's batch
__global__ void best_cells_kernel(int32_t *best, int64_t n_reads, int64_t n_refs, const int32_t *read_batch,
                                  const int32_t *read_slot, const uint64_t *const *batch_keys, const uint32_t *batch_n)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_reads) return;
    const int b = read_batch[q];
    const int ref = best[4 * q + 1];
    if (b < 0 || ref < 0 || best[4 * q] <= 0) return;
    const uint64_t *keys = batch_keys[b];
    const uint64_t p = (uint64_t)read_slot[q] * (uint64_t)n_refs + (uint64_t)ref;   // key order: (read slot, ref)
    const uint64_t target = make_key(p, 0, 0);
    uint32_t lo = 0, hi = batch_n[b];
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    if (lo < batch_n[b] && key_pair(keys[lo]) == p) {
        best[4 * q + 2] = (int32_t)key_i(keys[lo]);
        best[4 * q + 3] = (int32_t)key_j(keys[lo]);
    }
}

cudaError_t launch_best_cells(int32_t *best, int64_t n_reads, int64_t n_refs, const int32_t *read_batch,
                              const int32_t *read_slot, const uint64_t *const *batch_keys, const uint32_t *batch_n,
                              cudaStream_t st)
{
    if (n_reads == 0) return cudaSuccess;
    const int threads = 128;
    best_cells_kernel<<<(unsigned)((n_reads + threads - 1) / threads), threads, 0, st>>>(best, n_reads, n_refs, read_batch,
                                                                                        read_slot, batch_keys, batch_n);
    return cudaGetLastError();
}

size_t sort_keys_tmp_bytes(uint32_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)n);
    return bytes;
}

cudaError_t sort_keys(uint64_t *keys_in, uint64_t *keys_out, uint32_t n, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_in, keys_out, (int)n, 0, 64, st);
}

}  // namespace swb
