// swb_host.h -- host-side structures of libswb200 shared by swb_api.cu and swb_wide_host.cu.
#pragma once
#include "swb_internal.h"

#include <algorithm>
#include <atomic>
#include <map>
#include <tuple>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>

namespace swbh {
using namespace swb;

std::string &last_error();
inline int fail(int code, const std::string &msg) { last_error() = msg; return code; }
inline int cuda_fail(cudaError_t e, const char *what)
{
    last_error() = std::string(what) + ": " + cudaGetErrorString(e);
    return SWB_E_CUDA;
}
#define CU(x)                                                                  \
    do {                                                                       \
        cudaError_t e_ = (x);                                                  \
        if (e_ != cudaSuccess) return swbh::cuda_fail(e_, #x);                 \
    } while (0)

// Large device blocks (>= 4 MiB) are recycled by EXACT size per stream, outside the CUDA pool: the stream-ordered
// pool splits and merges its free blocks, so a 730 MB result array released by one call is carved up by the next
// call's smaller requests and the pool has to be reshaped when the 730 MB request comes back -- measured on B200 as
// single cudaMallocAsync calls of 0.2 - 1.4 s, 15 % of a 100,000-read job.  A block cached here is handed out again
// only for the same (rounded) size on the same stream, which keeps stream order: whatever still uses it was issued
// on that stream before.  The cache is bounded; overflow goes back to the pool.
struct BlockCache {
    std::mutex mu;
    std::multimap<std::pair<cudaStream_t, size_t>, void *> free_blocks;
    size_t cached = 0;
    static constexpr size_t MIN_BYTES = (size_t)4 << 20, MAX_CACHED = (size_t)24 << 30;
    static BlockCache &get() { static BlockCache c; return c; }
    void *take(cudaStream_t st, size_t bytes)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_blocks.find({st, bytes});
        if (it == free_blocks.end()) return nullptr;
        void *p = it->second;
        free_blocks.erase(it);
        cached -= bytes;
        return p;
    }
    bool give(cudaStream_t st, size_t bytes, void *p)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cached + bytes > MAX_CACHED) return false;
        free_blocks.emplace(std::make_pair(st, bytes), p);
        cached += bytes;
        return true;
    }
    // the pool ran dry: hand this stream's cached blocks back (stream-ordered), the caller retries its allocation
    size_t spill(cudaStream_t st)
    {
        std::lock_guard<std::mutex> lk(mu);
        size_t freed = 0;
        for (auto it = free_blocks.begin(); it != free_blocks.end();)
            if (it->first.first == st) { cudaFreeAsync(it->second, st); freed += it->first.second; cached -= it->first.second; it = free_blocks.erase(it); } else ++it;
        return freed;
    }
    // a context goes away: its streams' blocks return to the pool (the caller has synchronised the streams)
    void purge(cudaStream_t st)
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto it = free_blocks.begin(); it != free_blocks.end();)
            if (it->first.first == st) { cudaFree(it->second); cached -= it->first.second; it = free_blocks.erase(it); } else ++it;
    }
};

// Device buffer from the stream-ordered pool (cudaMallocAsync): allocation and release are
// ordered on the engine's stream and reuse pool memory, so the align loop never hits the
// synchronising cudaMalloc/cudaFree.
template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0; cudaStream_t st = nullptr;
    bool owned = true;                          // false: a view into somebody else's allocation
    size_t cap_bytes = 0;                       // what was asked from the allocator (size class)
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), st(o.st), owned(o.owned), cap_bytes(o.cap_bytes) { o.p = nullptr; o.n = 0; o.owned = true; o.cap_bytes = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; owned = o.owned; cap_bytes = o.cap_bytes; o.p = nullptr; o.n = 0; o.owned = true; o.cap_bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p && owned && !(cap_bytes >= BlockCache::MIN_BYTES && BlockCache::get().give(st, cap_bytes, p))) cudaFreeAsync(p, st);
        p = nullptr; n = 0; owned = true; cap_bytes = 0;
    }
    void view(T *ptr, size_t count) { release(); p = ptr; n = count; owned = false; }
    cudaError_t alloc(size_t count, cudaStream_t stream)
    {
        release();
        if (count == 0) count = 1;
        st = stream;
        // size classes (eight per octave, <= 12.5 % over) above 1 MiB: the arrays of consecutive calls differ by a
        // percent or so (max-cell counts) and land in the same class
        size_t bytes = count * sizeof(T);
        if (bytes >= ((size_t)1 << 20)) {
            size_t step = (size_t)1 << 17;
            while ((step << 4) <= bytes) step <<= 1;               // step = 2^(floor(log2 bytes) - 3)
            bytes = (bytes + step - 1) / step * step;
        }
        cap_bytes = bytes;
        if (bytes >= BlockCache::MIN_BYTES) {
            p = static_cast<T *>(BlockCache::get().take(stream, bytes));
            if (p) { n = count; return cudaSuccess; }
        }
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e == cudaErrorMemoryAllocation && BlockCache::get().spill(stream) > 0) {
            cudaGetLastError();
            e = cudaMallocAsync(&p, bytes, stream);                // once more with the cached blocks back in the pool
        }
        if (e == cudaSuccess) n = count; else { p = nullptr; cap_bytes = 0; }
        return e;
    }
    cudaError_t reserve(size_t count, cudaStream_t stream) { return n >= count && p ? cudaSuccess : alloc(count + count / 8, stream); }
};

inline int upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

}  // namespace swbh

using swbh::DevBuf;
using swb::TileTask;
using swb::WideTask;

struct swb_ctx {
    int device = 0;
    int sm_count = 0;
    int64_t ws_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_fill = nullptr;         // second stream: fill of batch k+1 runs beside the traceback of batch k
    cudaStream_t stream_copy = nullptr;         // results of reference part k cross PCIe while part k+1 computes
    bool pipeline = true;
    DevBuf<uint32_t> ck2[2], tmx2[2];           // ping-pong checkpoint workspaces
    DevBuf<int32_t> rp2[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::recursive_mutex mu;
    std::atomic<int> live{0};                   // handles (refsets, read batches, results) that still point here
    std::atomic<bool> dying{false};             // swb_destroy was called while handles were alive: the last free destroys
    // grow-only scratch reused by every align call on this context
    DevBuf<uint32_t> ck, tmx, counters;
    DevBuf<int32_t> rp, slot;
    DevBuf<int32_t> v_ref, v_c0, v_len, v_skip, v_end;   // fill work units (reference segments) of the current class
    DevBuf<TileTask> tasks;
    DevBuf<uint8_t> task_hits;                 // per candidate tile: exact maximum + buffered cells (locate scan -> emit)
    DevBuf<uint64_t> keys_tmp;
    DevBuf<uint8_t> sort_tmp;
    // wide (int32, long-pair) path scratch
    DevBuf<int32_t> w_brow, w_rec, w_tmx, w_prog, w_pair_ref, w_pair_read, w_wreads;
    DevBuf<uint8_t> w_rpad;                     // wide reads, aligned + padded (0xFE) to whole bands
    DevBuf<int64_t> w_rpad_off, w_rpad_len;
    DevBuf<unsigned long long> w_dbg;
    DevBuf<int32_t> w_mail;                     // tokens of the pipelined CTA-wide traceback
    DevBuf<int64_t> w_band_off, w_blk_off, w_brow_off;
    DevBuf<int2> w_items;
    DevBuf<WideTask> w_tasks;
    // pinned host buffers, recycled between results (cudaHostAlloc is slow; D2H into pageable memory too)
    struct PinBuf { void *p = nullptr; size_t bytes = 0; };
    std::vector<PinBuf> pin_free;
    // Pinned host buffers come in power-of-two size classes (>= 64 KiB): a step whose result is a few percent larger
    // than the previous one's finds its buffers in the pool instead of paying cudaHostAlloc (milliseconds per 100 MB).
    static size_t pin_class(size_t bytes)
    {
        size_t c = (size_t)64 << 10;
        while (c < bytes) c <<= 1;
        return c;
    }
    int64_t pin_allocs = 0;                     // cudaHostAlloc calls so far (SWB_TIMELINE prints them)
    PinBuf pin_get(size_t bytes)
    {
        const size_t want = pin_class(bytes);
        size_t best = pin_free.size();
        for (size_t k = 0; k < pin_free.size(); ++k)
            if (pin_free[k].bytes >= want && (best == pin_free.size() || pin_free[k].bytes < pin_free[best].bytes)) best = k;
        if (best != pin_free.size()) { PinBuf b = pin_free[best]; pin_free.erase(pin_free.begin() + best); return b; }
        PinBuf b;
        b.bytes = want;
        ++pin_allocs;
        if (cudaHostAlloc(&b.p, b.bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; b.bytes = 0; }
        return b;
    }
    void pin_put(PinBuf b) { if (b.p) pin_free.push_back(b); }
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaError_t next_event(cudaEvent_t *e)
    {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t n;
            cudaError_t rc = cudaEventCreate(&n);
            if (rc != cudaSuccess) return rc;
            ev_pool.push_back(n);
        }
        *e = ev_pool[ev_used++];
        return cudaSuccess;
    }
};

struct swb_refset {
    swb_ctx *ctx = nullptr;
    int64_t n_refs = 0, total_bases = 0, blocks_per_rp = 0;
    int32_t max_len = 0;
    int n_symbols = 0;
    uint8_t code_of[128];                       // upper-cased ASCII byte -> code (rank among the set's symbols), 0xFF = absent
    bool two_bit_ok = false;                    // <= 4 symbols: the 2-bit packed short-read path is available
    DevBuf<uint8_t> codes8;                     // 1 byte per base, original order (wide path)
    DevBuf<int64_t> off8;                       // [n_refs + 1]
    std::vector<int32_t> len_orig;
    std::vector<int32_t> len_sorted;            // lengths in the device (descending-length) order
    DevBuf<uint32_t> words, word_off;
    DevBuf<int32_t> len, orig, sorted_of;
    DevBuf<int64_t> blk_off;
    // A large set is cut into PARTS: contiguous ranges of the caller's reference indices, each a complete packed set
    // of its own.  An align call walks the parts in order; the results of part k are a contiguous segment of every
    // ABI array (pair = ref * n_reads + read), so they cross PCIe while part k + 1 computes.  The parent keeps the
    // alphabet, the lengths and the totals; `parts` is empty for a part (and for a set small enough to be one).
    // fill work units (reference segments, make_segments) per (K, scores): device tables built on first use.  Guarded by
    // the context mutex like every align call.
    struct SegKey {
        int K, match, mismatch, gap, few;
        bool operator<(const SegKey &o) const { return std::tie(K, match, mismatch, gap, few) < std::tie(o.K, o.match, o.mismatch, o.gap, o.few); }
    };
    struct SegTables { size_t nv = 0; DevBuf<int32_t> ref, c0, len, skip, end; };
    mutable std::map<SegKey, std::shared_ptr<SegTables>> seg_cache;
    std::vector<swb_refset *> parts;
    std::vector<int64_t> part_first;            // first global reference index of each part
    swb_refset *whole = nullptr;                // a set with parts: the same references as one set (device-resident / scores-only calls)
    const swb_refset *parent = nullptr;
};

struct swb_reads {
    swb_ctx *ctx = nullptr;
    const swb_refset *rs = nullptr;
    int64_t n_reads = 0;
    std::vector<int32_t> len;
    DevBuf<uint8_t> codes;
    DevBuf<int64_t> off;
};

struct BatchOut {
    int K = 0;
    bool wide = false;                           // produced by the int32 long-pair path (different key layout)
    std::vector<int64_t> wide_pair_p;            // wide: local pair -> ABI pair index p
    uint32_t n_cells = 0;
    int64_t ops_stride = 0;
    DevBuf<uint64_t> keys;
    DevBuf<int32_t> beg, oplen;
    DevBuf<uint32_t> ops;
    std::vector<int32_t> slot_read;              // read slot of this batch -> read index of the call
    DevBuf<int32_t> d_slot_read;                 // the same on the device, for the assembly
    DevBuf<int64_t> d_pair_map;                  // wide path: local pair -> p, on the device
};

struct swb_result {
    swb_ctx *ctx = nullptr;
    int64_t n_refs = 0, n_reads = 0;
    uint32_t flags = 0;
    bool fetched = false;
    std::vector<int32_t> ref_len, read_len;
    DevBuf<int32_t> d_scores, d_totals, d_best;
    std::vector<BatchOut> batches;               // released once the result is assembled
    // assembled on the device (swb_assemble.cu), ABI order
    uint32_t total_cells = 0;
    int64_t total_words = 0;
    DevBuf<int32_t> f_cells, f_beg, f_len;
    DevBuf<int64_t> f_ops_off, f_cell_off;
    DevBuf<uint32_t> f_ops;
    // host views into pinned buffers (after fetch)
    std::vector<swb_ctx::PinBuf> pins;
    const int32_t *scores = nullptr, *totals = nullptr, *best = nullptr;
    const int64_t *cell_off = nullptr, *ops_off = nullptr;
    const int32_t *cells = nullptr, *beginnings = nullptr, *op_lens = nullptr;
    const uint32_t *ops = nullptr;
    double stats[12] = {0};
    std::vector<char> own;                      // host arrays of a 1 x 1 result cut out of a coalesced batch (swb_queue.cu)
    // result of a multi-part reference set: one sub-result per part (device arrays), stitched into the host arrays
    std::vector<swb_result *> subs;
    std::vector<int64_t> sub_first;
    size_t parts_total = 0;                     // parts the call will produce (capacity guess of the host arrays)
    std::vector<char> sub_copied;               // part k's arrays are already on their way to the host buffers
    swb_ctx::PinBuf h_scores, h_totals, h_best, h_cell_off, h_cells, h_beg, h_len, h_ops_off, h_ops;
    uint64_t cells_done = 0;                    // cells / words of the parts copied so far
    int64_t words_done = 0;
};


#include <functional>
namespace swbh {
int run_wide_path(swb_ctx *ctx, const swb_refset *rs, const swb_reads *rd, const std::vector<int32_t> &reads,
                  int match, int mismatch, int gap, uint32_t flags, swb_result *res, int *launches, int *n_batches,
                  double *ck_bytes, const std::function<cudaError_t(int)> &tic, const std::function<cudaError_t()> &toc);
}
