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
__global__ void __launch_bounds__(128) locate_kernel(const BatchParams P, const TileTask *tasks, uint32_t n_tasks,
                                                      uint64_t *keys, uint32_t cap, uint32_t *count)
{
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
        if (live) {
            init_group<K>(P, (int)(T.rp_half >> 1), (int)(T.rp_half & 1), (int)T.ref_sorted, t, C, blk0, pair, read_idx);
            load_state<K>(P, blk0, (int)T.block, (int)(T.rp_half & 1), t, H, diag);
            S = P.scores[pair];
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
                                 if (k < cap) keys[k] = make_key((uint64_t)pair, (uint32_t)i, (uint32_t)j);
                             }
                         }
                     });
    }
}

template <int K>
static cudaError_t launch_locate_k(const BatchParams &P, const TileTask *tasks, uint32_t n_tasks, uint64_t *keys,
                                   uint32_t cap, uint32_t *count, int sm_count, cudaStream_t st)
{
    if (n_tasks == 0) return cudaSuccess;
    const int threads = 128;                      // 16 groups per CTA
    int64_t ctas = ((int64_t)n_tasks + 15) / 16;
    ctas = std::min<int64_t>(ctas, (int64_t)sm_count * 16);
    locate_kernel<K><<<(unsigned)ctas, threads, 0, st>>>(P, tasks, n_tasks, keys, cap, count);
    return cudaGetLastError();
}

cudaError_t launch_locate(int K, const BatchParams &P, const TileTask *tasks, uint32_t n_tasks, uint64_t *keys,
                          uint32_t cap, uint32_t *count, int sm_count, cudaStream_t st)
{
    switch (K) {
        case 4:  return launch_locate_k<4>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 8:  return launch_locate_k<8>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 13: return launch_locate_k<13>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 16: return launch_locate_k<16>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 19: return launch_locate_k<19>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 25: return launch_locate_k<25>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
        case 32: return launch_locate_k<32>(P, tasks, n_tasks, keys, cap, count, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------
// Traceback.  Tile of one group in shared memory (int16):
//   tile[t][c][r], c = 0..CB, r = 0..K.   c = 0 is the checkpointed column of lane t,
//   c >= 1 the column computed at step u = c-1; r = 0 is the boundary row received from
//   lane t-1 (row t*K of the matrix), r >= 1 the lane's own rows.  With that halo every
//   neighbour (W, N, NW) of a cell with c >= 1, r >= 1 lies in the same lane's tile.
template <int K>
__global__ void __launch_bounds__(128) trace_kernel(const BatchParams P, const uint64_t *keys, uint32_t n_cells,
                                                     int32_t *beginnings, int32_t *op_lens, uint32_t *ops,
                                                     int ops_stride, int groups_per_cta)
{
    using G = Geo<K>;
    extern __shared__ __align__(16) int16_t tiles[];
    const int lane = threadIdx.x & 31, t = lane & (GL - 1), g = lane >> 3;
    const int warp = threadIdx.x >> 5;
    const unsigned gmask = 0xffu << (8 * g);
    const int gl = warp * 4 + g;                                   // group within CTA
    int16_t *tile = tiles + (size_t)gl * G::TILE_HALFWORDS;
    int16_t *mytile = tile + (size_t)t * (CB + 1) * G::RS;
    const uint32_t n_groups = gridDim.x * groups_per_cta;
    const uint32_t gid = blockIdx.x * groups_per_cta + gl;

    // per-group state (identical in all 8 lanes; lane 0 walks and broadcasts)
    bool busy = false;
    uint32_t cell = 0, next_cell = gid;
    int ci = 0, cj = 0, hcur = 0, beginning = 0, oplen = 0, half = 0, b = 0;
    uint32_t opword = 0;
    GroupCtx<K> C;
    int64_t blk0 = 0, pair = 0; int read_idx = 0;
    C.n = 0; C.m = 0; C.ref_words = P.ref_words; C.match = C.mismatch = C.gap = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) C.rc[r] = 0xFE;

    for (;;) {
        if (!busy && next_cell < n_cells) {
            cell = next_cell; next_cell += n_groups;
            const uint64_t key = keys[cell];
            pair = (int64_t)key_pair(key);
            ci = (int)key_i(key); cj = (int)key_j(key);
            const int64_t ro = pair / P.n_reads;
            const int rd = (int)(pair - ro * P.n_reads);
            const int slot = P.read_slot[rd];
            const int ref = P.ref_sorted_of[ro];
            half = slot & 1;
            init_group<K>(P, slot >> 1, half, ref, t, C, blk0, pair, read_idx);
            hcur = P.scores[pair];
            beginning = 0; oplen = 0; opword = 0;
            busy = true;
        }
        if (!__any_sync(0xffffffffu, busy)) break;

        // ---- recompute the block that holds the current cell ---------------------------
        {
            const int tc = (ci - 1) / K;
            b = busy ? (cj - 1 + tc) / CB : 0;
        }
        int H[K], diag;
        if (busy) load_state<K>(P, blk0, b, half, t, H, diag);
        else {
#pragma unroll
            for (int r = 0; r < K; ++r) H[r] = 0;
            diag = 0;
        }
        // halo column c = 0
        {
            mytile[0] = (int16_t)diag;
#pragma unroll
            for (int r = 0; r < K; ++r) mytile[r + 1] = (int16_t)H[r];
        }
        run_block<K>(C, b, t, gmask, H, diag,
                     [&](int u, int top, const int (&Hc)[K], bool, int) {
                         int16_t *col = mytile + (u + 1) * G::RS;
                         col[0] = (int16_t)top;
#pragma unroll
                         for (int r = 0; r < K; ++r) col[r + 1] = (int16_t)Hc[r];
                     });
        __syncwarp();

        // ---- walk inside the tile (lane 0 of the group) --------------------------------
        int done = 0;
        if (busy && t == 0) {
            uint32_t *myops = ops + (size_t)cell * ops_stride;
            while (hcur > 0) {
                const int tc = (ci - 1) / K;
                const int r = ci - tc * K;                          // 1..K
                const int c = cj - (b * CB - tc);                   // column index in lane tc's tile
                if (c < 1 || c > CB) break;                         // the cell belongs to an earlier block
                const int16_t *lt = tile + ((size_t)tc * (CB + 1) + c) * G::RS + r;
                const int hw = lt[-G::RS];                          // W  = tile[tc][c-1][r]
                const int hn = lt[-1];                              // N  = tile[tc][c][r-1]
                const int hnw = lt[-G::RS - 1];                     // NW = tile[tc][c-1][r-1]
                const int col = cj - 1;
                const int rc = (int)((__ldg(C.ref_words + (col >> 4)) >> (2 * (col & 15))) & 3u);
                const int qc = (int)P.read_codes[P.read_off[read_idx] + ci - 1];
                const int sc = (qc == rc) ? C.match : C.mismatch;
                beginning = cj;
                uint32_t op;
                if (hnw + sc == hcur)        { op = 1; --ci; --cj; hcur = hnw; }   // alignment
                else if (hn + C.gap == hcur) { op = 2; --ci;       hcur = hn;  }   // insertion
                else                         { op = 3;       --cj; hcur = hw;  }   // deletion
                opword |= op << (2 * (oplen & 15));
                ++oplen;
                if ((oplen & 15) == 0) { myops[(oplen >> 4) - 1] = opword; opword = 0; }
            }
            if (hcur <= 0) {
                if (oplen & 15) myops[oplen >> 4] = opword;
                beginnings[cell] = beginning;
                op_lens[cell] = oplen;
                done = 1;
            }
        }
        // broadcast the walker's state to the group
        done = __shfl_sync(gmask, done, 0, GL);
        ci = __shfl_sync(gmask, ci, 0, GL);
        cj = __shfl_sync(gmask, cj, 0, GL);
        if (done) busy = false;
        __syncwarp();
    }
}

template <int K>
static cudaError_t launch_trace_k(const BatchParams &P, const uint64_t *keys, uint32_t n_cells, int32_t *beginnings,
                                  int32_t *op_lens, uint32_t *ops, int ops_stride, int sm_count, cudaStream_t st)
{
    using G = Geo<K>;
    if (n_cells == 0) return cudaSuccess;
    const size_t per_group = (size_t)G::TILE_HALFWORDS * sizeof(int16_t);
    int warps = 4;
    while (warps > 1 && per_group * 4 * warps > 200 * 1024) --warps;
    const size_t smem = per_group * 4 * warps;
    static bool attr_set[64] = {false};
    if (!attr_set[K]) {
        cudaError_t e = cudaFuncSetAttribute(trace_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[K] = true;
    }
    const int gpc = warps * 4;
    int64_t ctas = ((int64_t)n_cells + gpc - 1) / gpc;
    ctas = std::min<int64_t>(ctas, (int64_t)sm_count * 2);
    trace_kernel<K><<<(unsigned)ctas, warps * 32, smem, st>>>(P, keys, n_cells, beginnings, op_lens, ops, ops_stride, gpc);
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
__global__ void best_cells_kernel(int32_t *best, int64_t n_reads, const int32_t *read_batch,
                                  const uint64_t *const *batch_keys, const uint32_t *batch_n)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_reads) return;
    const int b = read_batch[q];
    const int ref = best[4 * q + 1];
    if (b < 0 || ref < 0 || best[4 * q] <= 0) return;
    const uint64_t *keys = batch_keys[b];
    const uint64_t p = (uint64_t)ref * (uint64_t)n_reads + (uint64_t)q;
    const uint64_t target = make_key(p, 0, 0);
    uint32_t lo = 0, hi = batch_n[b];
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    if (lo < batch_n[b] && key_pair(keys[lo]) == p) {
        best[4 * q + 2] = (int32_t)key_i(keys[lo]);
        best[4 * q + 3] = (int32_t)key_j(keys[lo]);
    }
}

cudaError_t launch_best_cells(int32_t *best, int64_t n_reads, const int32_t *read_batch,
                              const uint64_t *const *batch_keys, const uint32_t *batch_n, cudaStream_t st)
{
    if (n_reads == 0) return cudaSuccess;
    const int threads = 128;
    best_cells_kernel<<<(unsigned)((n_reads + threads - 1) / threads), threads, 0, st>>>(best, n_reads, read_batch,
                                                                                        batch_keys, batch_n);
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
