// swb_multi.cu -- multi-GPU inside the C ABI (include/swb200.h, "multi-GPU" section).
//
// The reference partitions the pair set by REFERENCE (`sc.parallelize(refs)` + map(MapRef),
// Distribution.java:337-338).  Here every device keeps a length-balanced shard of the reference set
// resident in its HBM, aligns all reads against it (swb_align on its own context), and only the
// per-read best-hit records (score, global ref, i, j) cross NVLink: one ncclAllGather issued on each
// device's engine stream after the best-hit kernel, then a merge kernel on the device.
//   swb_multi_*  one process, n devices, one host thread per device inside swb_multi_align
//   swb_comm_*   one process per device (torchrun / MPI); the host ships the 128-byte NCCL id
// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it and a
// single-device host never loads it.
#include "swb_host.h"

#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only; the entry points are looked up with dlsym

#include <chrono>
#include <thread>

using namespace swbh;

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("SWB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("NCCL not found (set SWB_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return; }
#define SWB_NCCL_SYM(field, sym)                                                       \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));     \
        if (!api.field) { api.error = std::string("NCCL symbol missing: ") + sym; return; }
        SWB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        SWB_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        SWB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        SWB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        SWB_NCCL_SYM(AllGather, "ncclAllGather")
        SWB_NCCL_SYM(GroupStart, "ncclGroupStart")
        SWB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        SWB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SWB_NCCL_SYM
    });
    return &api;
}

int nccl_fail(ncclResult_t r, const char *what)
{
    NcclApi *a = nccl_api();
    return fail(SWB_E_CUDA, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}
#define NC(x)                                                                  \
    do {                                                                       \
        ncclResult_t r_ = (x);                                                 \
        if (r_ != ncclSuccess) return nccl_fail(r_, #x);                       \
    } while (0)

// shard-local reference index -> global index (ids), -1 stays -1 (a shard without references)
__global__ void best_to_global_kernel(const int32_t *best, const int64_t *ids, int64_t n_ids, int64_t n_reads, int32_t *out)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_reads) return;
    int4 b = reinterpret_cast<const int4 *>(best)[q];
    b.y = (b.y >= 0 && b.y < n_ids) ? (int32_t)ids[b.y] : -1;
    reinterpret_cast<int4 *>(out)[q] = b;
}

// gathered[w][q] -> merged[q]: highest score, then lowest global reference index; a record without a reference
// (ref < 0, an empty shard) loses against every real one
__global__ void merge_best_kernel(const int32_t *gathered, int world, int64_t n_reads, int32_t *merged)
{
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_reads) return;
    int4 best = make_int4(0, -1, 0, 0);
    for (int w = 0; w < world; ++w) {
        const int4 b = reinterpret_cast<const int4 *>(gathered)[(int64_t)w * n_reads + q];
        if (b.y < 0) continue;
        if (best.y < 0 || b.x > best.x || (b.x == best.x && b.y < best.y)) best = b;
    }
    reinterpret_cast<int4 *>(merged)[q] = best;
}

// send/gather/merge buffers of one device
struct GatherBufs {
    DevBuf<int32_t> send, gathered, merged;
    DevBuf<int64_t> ids;
    int64_t n_ids = 0;
    void release() { send.release(); gathered.release(); merged.release(); ids.release(); }
};

// localize -> allgather -> merge on `st` of the current device.  `grouped`: the caller brackets several devices'
// collectives with ncclGroupStart / ncclGroupEnd (single-process form) and launches the merge afterwards.
cudaError_t enqueue_localize(GatherBufs &g, const int32_t *d_best, int64_t n_reads, int world, cudaStream_t st)
{
    cudaError_t e = g.send.reserve((size_t)n_reads * 4, st);
    if (e == cudaSuccess) e = g.gathered.reserve((size_t)n_reads * 4 * world, st);
    if (e == cudaSuccess) e = g.merged.reserve((size_t)n_reads * 4, st);
    if (e != cudaSuccess || n_reads == 0) return e;
    const int threads = 256;
    best_to_global_kernel<<<(unsigned)((n_reads + threads - 1) / threads), threads, 0, st>>>(d_best, g.ids.p, g.n_ids, n_reads, g.send.p);
    return cudaGetLastError();
}
cudaError_t enqueue_merge(GatherBufs &g, int64_t n_reads, int world, cudaStream_t st)
{
    if (n_reads == 0) return cudaSuccess;
    const int threads = 256;
    merge_best_kernel<<<(unsigned)((n_reads + threads - 1) / threads), threads, 0, st>>>(g.gathered.p, world, n_reads, g.merged.p);
    return cudaGetLastError();
}

}  // namespace

struct swb_comm {
    swb_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    GatherBufs g;
    int64_t n_reads = 0;
};

struct swb_multi {
    std::vector<int> devices;
    std::vector<swb_ctx *> ctx;
    std::vector<swb_refset *> rs;
    std::vector<std::vector<int64_t>> ids;       // per shard: global reference indices, ascending
    std::vector<int32_t> shard_of;               // global reference -> shard
    std::vector<int64_t> local_of;               // global reference -> index in the shard
    std::vector<ncclComm_t> comms;
    std::vector<GatherBufs> g;
    int64_t n_refs = 0;
    std::mutex mu;
};

struct swb_multi_result {
    swb_multi *m = nullptr;
    std::vector<swb_result *> shard;
    std::vector<int32_t> merged;
    int64_t n_reads = 0;
    double stats[3] = {0, 0, 0};
};

extern "C" {

int swb_comm_unique_id(char *id128)
{
    if (!id128) return fail(SWB_E_INVALID, "swb_comm_unique_id: null");
    NcclApi *a = nccl_api();
    if (!a->handle || !a->GetUniqueId) return fail(SWB_E_UNSUPPORTED, "swb_comm_unique_id: " + a->error);
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI ships the NCCL id as 128 bytes");
    ncclUniqueId id;
    NC(a->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SWB_OK;
}

int swb_comm_create(swb_ctx *ctx, const char *id128, int32_t rank, int32_t world, swb_comm **out)
{
    if (!ctx || !out || (world > 1 && !id128)) return fail(SWB_E_INVALID, "swb_comm_create: null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(SWB_E_INVALID, "swb_comm_create: bad rank / world");
    std::unique_ptr<swb_comm> c(new swb_comm());
    c->ctx = ctx; c->rank = rank; c->world = world;
    if (world > 1) {
        NcclApi *a = nccl_api();
        if (!a->handle || !a->CommInitRank) return fail(SWB_E_UNSUPPORTED, "swb_comm_create: " + a->error);
        CU(cudaSetDevice(ctx->device));
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        NC(a->CommInitRank(&c->comm, world, id, rank));
    }
    *out = c.release();
    return SWB_OK;
}

void swb_comm_destroy(swb_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    c->g.release();
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    delete c;
}

int swb_comm_allgather_best(swb_comm *c, const swb_result *res, const int64_t *global_ids, int64_t n_local_refs,
                            int32_t *merged_host)
{
    if (!c || !res) return fail(SWB_E_INVALID, "swb_comm_allgather_best: null argument");
    if (res->ctx != c->ctx) return fail(SWB_E_INVALID, "swb_comm_allgather_best: the result belongs to another context");
    if (n_local_refs != res->n_refs) return fail(SWB_E_INVALID, "swb_comm_allgather_best: global_ids must cover the shard's references");
    if (n_local_refs > 0 && !global_ids) return fail(SWB_E_INVALID, "swb_comm_allgather_best: global_ids is null");
    swb_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n_reads = res->n_reads;
    c->n_reads = n_reads;
    CU(c->g.ids.reserve((size_t)std::max<int64_t>(n_local_refs, 1), st));
    c->g.n_ids = n_local_refs;
    if (n_local_refs) CU(cudaMemcpyAsync(c->g.ids.p, global_ids, (size_t)n_local_refs * 8, cudaMemcpyHostToDevice, st));
    CU(enqueue_localize(c->g, res->d_best.p, n_reads, c->world, st));
    if (c->world > 1) {
        if (n_reads) NC(nccl_api()->AllGather(c->g.send.p, c->g.gathered.p, (size_t)n_reads * 4, ncclInt32, c->comm, st));
    } else if (n_reads) {
        CU(cudaMemcpyAsync(c->g.gathered.p, c->g.send.p, (size_t)n_reads * 16, cudaMemcpyDeviceToDevice, st));
    }
    CU(enqueue_merge(c->g, n_reads, c->world, st));
    if (merged_host && n_reads) CU(cudaMemcpyAsync(merged_host, c->g.merged.p, (size_t)n_reads * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));        // global_ids / merged_host are the caller's
    return SWB_OK;
}

int swb_comm_merged_device_ptr(swb_comm *c, void **ptr, int64_t *n_elems)
{
    if (!c || !ptr) return fail(SWB_E_INVALID, "swb_comm_merged_device_ptr: null");
    *ptr = c->g.merged.p;
    if (n_elems) *n_elems = c->n_reads * 4;
    return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------
int swb_multi_create(const int32_t *devices, int32_t n_devices, int64_t workspace_bytes, swb_multi **out)
{
    if (!devices || !out || n_devices < 1) return fail(SWB_E_INVALID, "swb_multi_create: bad arguments");
    *out = nullptr;
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return fail(SWB_E_INVALID, "swb_multi_create: a device is listed twice");
    std::unique_ptr<swb_multi> m(new swb_multi());
    auto cleanup = [&] { for (swb_ctx *c : m->ctx) swb_destroy(c); m->ctx.clear(); };
    for (int d = 0; d < n_devices; ++d) {
        swb_ctx *c = nullptr;
        const int rc = swb_create(devices[d], workspace_bytes, &c);
        if (rc) { cleanup(); return rc; }
        m->ctx.push_back(c);
        m->devices.push_back(devices[d]);
    }
    m->rs.assign((size_t)n_devices, nullptr);
    m->ids.resize((size_t)n_devices);
    m->g.resize((size_t)n_devices);
    if (n_devices > 1) {
        NcclApi *a = nccl_api();
        if (!a->handle || !a->CommInitAll) { cleanup(); return fail(SWB_E_UNSUPPORTED, "swb_multi_create: " + a->error); }
        m->comms.assign((size_t)n_devices, nullptr);
        const ncclResult_t r = a->CommInitAll(m->comms.data(), n_devices, m->devices.data());
        if (r != ncclSuccess) { cleanup(); return nccl_fail(r, "ncclCommInitAll"); }
    }
    *out = m.release();
    return SWB_OK;
}

void swb_multi_destroy(swb_multi *m)
{
    if (!m) return;
    for (size_t d = 0; d < m->ctx.size(); ++d) {
        cudaSetDevice(m->devices[d]);
        cudaStreamSynchronize(m->ctx[d]->stream);
        m->g[d].release();
        if (m->rs[d]) swb_refset_free(m->rs[d]);
        if (d < m->comms.size() && m->comms[d]) nccl_api()->CommDestroy(m->comms[d]);
    }
    for (swb_ctx *c : m->ctx) swb_destroy(c);
    delete m;
}

int32_t swb_multi_device_count(const swb_multi *m) { return m ? (int32_t)m->ctx.size() : 0; }
swb_ctx *swb_multi_ctx(swb_multi *m, int32_t d) { return (m && d >= 0 && d < (int32_t)m->ctx.size()) ? m->ctx[(size_t)d] : nullptr; }
int64_t swb_multi_ref_count(const swb_multi *m) { return m ? m->n_refs : 0; }

int swb_multi_refset_load(swb_multi *m, int64_t n_refs, const char *bytes, const int64_t *offsets)
{
    if (!m) return fail(SWB_E_INVALID, "swb_multi_refset_load: null");
    if (n_refs < 0 || !offsets) return fail(SWB_E_INVALID, "swb_multi_refset_load: bad arguments");
    for (int64_t k = 0; k < n_refs; ++k)
        if (offsets[k + 1] < offsets[k]) return fail(SWB_E_INVALID, "swb_multi_refset_load: offsets not monotone");
    std::lock_guard<std::mutex> lk(m->mu);
    const int n = (int)m->ctx.size();
    // descending length (stable), dealt in snake order: 0..n-1, n-1..0, ...
    std::vector<int64_t> order((size_t)n_refs);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b];
    });
    for (auto &v : m->ids) v.clear();
    m->shard_of.assign((size_t)n_refs, 0);
    m->local_of.assign((size_t)n_refs, 0);
    for (int64_t pos = 0; pos < n_refs; ++pos) {
        const int64_t round = pos / n, slot = pos % n;
        const int d = (int)((round % 2 == 0) ? slot : n - 1 - slot);
        m->ids[(size_t)d].push_back(order[(size_t)pos]);
    }
    for (int d = 0; d < n; ++d) {
        auto &v = m->ids[(size_t)d];
        std::sort(v.begin(), v.end());
        for (size_t k = 0; k < v.size(); ++k) { m->shard_of[(size_t)v[k]] = d; m->local_of[(size_t)v[k]] = (int64_t)k; }
    }
    m->n_refs = n_refs;
    // one loader thread per device: gather the shard's bytes, pack, upload
    std::vector<int> rcs((size_t)n, SWB_OK);
    std::vector<std::string> errs((size_t)n);
    std::vector<std::thread> th;
    for (int d = 0; d < n; ++d)
        th.emplace_back([&, d] {
            const auto &v = m->ids[(size_t)d];
            std::vector<int64_t> off(v.size() + 1, 0);
            for (size_t k = 0; k < v.size(); ++k) off[k + 1] = off[k] + (offsets[v[k] + 1] - offsets[v[k]]);
            std::string buf((size_t)off[v.size()], '\0');
            for (size_t k = 0; k < v.size(); ++k)
                memcpy(&buf[(size_t)off[k]], bytes + offsets[v[k]], (size_t)(off[k + 1] - off[k]));
            if (m->rs[(size_t)d]) { swb_refset_free(m->rs[(size_t)d]); m->rs[(size_t)d] = nullptr; }
            int rc = swb_refset_load(m->ctx[(size_t)d], (int64_t)v.size(), buf.data(), off.data(), &m->rs[(size_t)d]);
            if (rc == SWB_OK) {
                swb_ctx *c = m->ctx[(size_t)d];
                GatherBufs &g = m->g[(size_t)d];
                cudaError_t e = cudaSetDevice(c->device);
                if (e == cudaSuccess) e = g.ids.reserve(std::max<size_t>(v.size(), 1), c->stream);
                if (e == cudaSuccess && !v.empty()) e = cudaMemcpyAsync(g.ids.p, v.data(), v.size() * 8, cudaMemcpyHostToDevice, c->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
                g.n_ids = (int64_t)v.size();
                if (e != cudaSuccess) rc = cuda_fail(e, "swb_multi_refset_load: id table");
            }
            rcs[(size_t)d] = rc;
            if (rc) errs[(size_t)d] = swb_last_error();
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < n; ++d)
        if (rcs[(size_t)d]) return fail(rcs[(size_t)d], "shard " + std::to_string(d) + ": " + errs[(size_t)d]);
    return SWB_OK;
}

int swb_multi_ref_location(const swb_multi *m, int64_t global_ref, int32_t *shard, int64_t *local_ref)
{
    if (!m) return fail(SWB_E_INVALID, "swb_multi_ref_location: null");
    if (global_ref < 0 || global_ref >= m->n_refs) return fail(SWB_E_RANGE, "swb_multi_ref_location: bad reference index");
    if (shard) *shard = m->shard_of[(size_t)global_ref];
    if (local_ref) *local_ref = m->local_of[(size_t)global_ref];
    return SWB_OK;
}

const int64_t *swb_multi_shard_refs(const swb_multi *m, int32_t d, int64_t *n)
{
    if (!m || d < 0 || d >= (int32_t)m->ids.size()) { if (n) *n = 0; return nullptr; }
    if (n) *n = (int64_t)m->ids[(size_t)d].size();
    return m->ids[(size_t)d].data();
}

int swb_multi_align(swb_multi *m, int64_t n_reads, const char *read_bytes, const int64_t *read_offsets, int32_t match,
                    int32_t mismatch, int32_t gap, uint32_t flags, swb_multi_result **out)
{
    if (!m || !out) return fail(SWB_E_INVALID, "swb_multi_align: null argument");
    *out = nullptr;
    const int n = (int)m->ctx.size();
    for (int d = 0; d < n; ++d)
        if (!m->rs[(size_t)d]) return fail(SWB_E_INVALID, "swb_multi_align: no reference set loaded");
    std::lock_guard<std::mutex> lk(m->mu);
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<swb_multi_result> res(new swb_multi_result());
    res->m = m; res->n_reads = n_reads;
    res->shard.assign((size_t)n, nullptr);
    auto drop = [&] { for (swb_result *r : res->shard) if (r) swb_result_free(r); res->shard.clear(); };

    std::vector<int> rcs((size_t)n, SWB_OK);
    std::vector<std::string> errs((size_t)n);
    // phase 1: every device aligns all reads against its shard; the results stay in HBM
    {
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d)
            th.emplace_back([&, d] {
                rcs[(size_t)d] = swb_align(m->ctx[(size_t)d], m->rs[(size_t)d], n_reads, read_bytes, read_offsets, match, mismatch,
                                           gap, flags | SWB_F_NO_FETCH, &res->shard[(size_t)d]);
                if (rcs[(size_t)d]) errs[(size_t)d] = swb_last_error();
            });
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < n; ++d)
        if (rcs[(size_t)d]) { const int rc = rcs[(size_t)d]; const std::string e = errs[(size_t)d]; drop(); return fail(rc, "shard " + std::to_string(d) + ": " + e); }

    // phase 2: best hits -> global ids -> one all-gather over NVLink -> merge, on every device's engine stream
    const auto t1 = std::chrono::steady_clock::now();
    for (int d = 0; d < n; ++d) {
        swb_ctx *c = m->ctx[(size_t)d];
        cudaError_t e = cudaSetDevice(c->device);
        if (e == cudaSuccess) e = enqueue_localize(m->g[(size_t)d], res->shard[(size_t)d]->d_best.p, n_reads, n, c->stream);
        if (e != cudaSuccess) { drop(); return cuda_fail(e, "swb_multi_align: best hits -> global ids"); }
    }
    if (n > 1 && n_reads > 0) {
        NcclApi *a = nccl_api();
        ncclResult_t r = a->GroupStart();
        for (int d = 0; d < n && r == ncclSuccess; ++d)
            r = a->AllGather(m->g[(size_t)d].send.p, m->g[(size_t)d].gathered.p, (size_t)n_reads * 4, ncclInt32, m->comms[(size_t)d],
                             m->ctx[(size_t)d]->stream);
        const ncclResult_t r2 = a->GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { drop(); return nccl_fail(r, "swb_multi_align: ncclAllGather"); }
    }
    res->merged.assign((size_t)n_reads * 4, 0);
    for (int d = 0; d < n; ++d) {
        swb_ctx *c = m->ctx[(size_t)d];
        GatherBufs &g = m->g[(size_t)d];
        cudaError_t e = cudaSetDevice(c->device);
        if (e == cudaSuccess && n == 1 && n_reads)
            e = cudaMemcpyAsync(g.gathered.p, g.send.p, (size_t)n_reads * 16, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = enqueue_merge(g, n_reads, n, c->stream);
        if (e == cudaSuccess && d == 0 && n_reads)
            e = cudaMemcpyAsync(res->merged.data(), g.merged.p, (size_t)n_reads * 16, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) { drop(); return cuda_fail(e, "swb_multi_align: merge"); }
    }
    for (int d = 0; d < n; ++d) {
        cudaSetDevice(m->ctx[(size_t)d]->device);
        const cudaError_t e = cudaStreamSynchronize(m->ctx[(size_t)d]->stream);
        if (e != cudaSuccess) { drop(); return cuda_fail(e, "swb_multi_align: all-gather"); }
    }
    const auto t2 = std::chrono::steady_clock::now();
    // phase 3: results to the host (unless the caller keeps them in HBM), all devices at once
    if (!(flags & SWB_F_NO_FETCH)) {
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d)
            th.emplace_back([&, d] {
                rcs[(size_t)d] = swb_result_fetch(res->shard[(size_t)d]);
                if (rcs[(size_t)d]) errs[(size_t)d] = swb_last_error();
            });
        for (auto &t : th) t.join();
        for (int d = 0; d < n; ++d)
            if (rcs[(size_t)d]) { const int rc = rcs[(size_t)d]; const std::string e = errs[(size_t)d]; drop(); return fail(rc, "shard " + std::to_string(d) + ": " + e); }
    }
    const auto t3 = std::chrono::steady_clock::now();
    res->stats[0] = std::chrono::duration<double, std::milli>(t3 - t0).count();
    for (int d = 0; d < n; ++d) res->stats[1] = std::max(res->stats[1], res->shard[(size_t)d]->stats[5]);
    res->stats[2] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    *out = res.release();
    return SWB_OK;
}

void swb_multi_result_free(swb_multi_result *res)
{
    if (!res) return;
    for (swb_result *r : res->shard) if (r) swb_result_free(r);
    delete res;
}

swb_result *swb_multi_result_shard(swb_multi_result *res, int32_t d)
{
    return (res && d >= 0 && d < (int32_t)res->shard.size()) ? res->shard[(size_t)d] : nullptr;
}
const int32_t *swb_multi_result_best_hits(const swb_multi_result *res) { return res ? res->merged.data() : nullptr; }
int swb_multi_result_stats(const swb_multi_result *res, double *out, int n)
{
    if (!res || !out) return fail(SWB_E_INVALID, "swb_multi_result_stats: null");
    for (int k = 0; k < n && k < 3; ++k) out[k] = res->stats[k];
    return SWB_OK;
}

}  // extern "C"
