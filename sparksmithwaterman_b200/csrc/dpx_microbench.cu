// dpx_microbench.cu -- measures the integer / DPX issue rates that bound the
// Smith-Waterman cell update on sm_100a (R_int is measured,
// not assumed).  Each kernel runs NCHAIN independent dependency chains per thread
// so the pipe, not latency, limits it.  Result: warp-instructions per clock per SM
// and lane-ops per clock per SM for each instruction mix.
//
// Standalone:  dpx_microbench [iters]   -> one JSON object on stdout.
// Also linked into libswb200.so as swb_microbench_json().
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { uint32_t r; asm("add.s16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { uint32_t r; asm("max.s16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t max2relu(uint32_t a, uint32_t b) { uint32_t r; asm("max.s16x2.relu %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t r; asm("add.f16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

constexpr int NCHAIN = 8;
constexpr int UNROLL = 16;

enum Mix { VIADDMNMX16 = 0, VIMNMX3_16, VIADD16, IADD32, IMAD32, VIADDMNMX32, LOP, MIX_DPX_IMAD, MIX_DPX_LDS,
           MIX_SW_CELL, VIMNMX2_16, VIMNMX2_16_RELU, VIADDMNMX16_RELU, MIX_DPX_VIMNMX2, MIX_DPX_VIADD, MIX_VIMNMX2_IMAD,
           PRMT_OP, SHF_OP, MIX_DPX_SHFL, IMNMX32, MIX_VIADD_VIMNMX2,
           HMNMX2_OP, HADD2_OP, MIX_DPX_HMNMX2, MIX_DPX_HADD2, MIX_HCELL, MIX_DPX_CELL_AND_HCELL, NMIX };

template <int MIX>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t *out, long long *cyc, int iters, uint32_t a0, uint32_t b0, uint32_t one)
{
    __shared__ uint32_t sm[256];
    const long long c0 = clock64();
    sm[threadIdx.x] = a0 + threadIdx.x;
    __syncthreads();
    uint32_t x[NCHAIN], y[NCHAIN];
    asm volatile("" : "+r"(a0), "+r"(b0), "+r"(one));      // keep the operands in registers, not c[][]
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) { x[c] = a0 + c + threadIdx.x; y[c] = b0 * (c + 1); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < NCHAIN; ++c) {
                if (MIX == VIADDMNMX16) x[c] = max2(add2(x[c], b0), y[c]);
                else if (MIX == VIMNMX3_16) x[c] = max2(max2(x[c], b0), y[c]);
                else if (MIX == VIADD16) x[c] = add2(x[c], y[c]);
                else if (MIX == IADD32) { x[c] = x[c] + y[c]; y[c] = y[c] + x[c]; }
                else if (MIX == IMAD32) x[c] = x[c] * one + y[c];
                else if (MIX == VIADDMNMX32) x[c] = (uint32_t)__viaddmax_s32((int)x[c], (int)b0, (int)y[c]);
                else if (MIX == LOP) { x[c] = (x[c] & y[c]) ^ b0; y[c] = (y[c] | x[c]) ^ a0; }
                else if (MIX == MIX_DPX_IMAD) { x[c] = max2(add2(x[c], b0), y[c]); y[c] = y[c] * one + a0; }
                else if (MIX == MIX_DPX_LDS) { x[c] = max2(add2(x[c], b0), y[c]); if (c == 0) y[0] ^= sm[(threadIdx.x + u) & 255]; }
                else if (MIX == VIMNMX2_16) x[c] = max2(x[c], y[c]) ^ 0u, y[c] = max2(y[c], x[c]);
                else if (MIX == VIMNMX2_16_RELU) { x[c] = max2relu(x[c], y[c]); y[c] = max2relu(y[c], x[c]); }
                else if (MIX == VIADDMNMX16_RELU) x[c] = max2relu(add2(x[c], b0), y[c]);
                else if (MIX == MIX_DPX_VIMNMX2) { x[c] = max2(add2(x[c], b0), y[c]); y[c] = max2(y[c], x[c]); }
                else if (MIX == MIX_DPX_VIADD) { x[c] = max2(add2(x[c], b0), y[c]); y[c] = add2(y[c], a0); }
                else if (MIX == MIX_VIMNMX2_IMAD) { x[c] = max2(x[c], y[c]); y[c] = y[c] * one + x[c]; }
                else if (MIX == PRMT_OP) { x[c] = __byte_perm(x[c], y[c], 0x5410 + u); y[c] = __byte_perm(y[c], x[c], 0x3210 ^ u); }
                else if (MIX == SHF_OP) { x[c] = __funnelshift_r(x[c], y[c], 3); y[c] = __funnelshift_l(y[c], x[c], 5); }
                else if (MIX == MIX_DPX_SHFL) { x[c] = max2(add2(x[c], b0), y[c]); if (c == 0) y[0] ^= __shfl_up_sync(0xffffffffu, x[1], 1); }
                else if (MIX == IMNMX32) { x[c] = (uint32_t)max((int)x[c], (int)y[c]); y[c] = (uint32_t)min((int)y[c], (int)x[c]) + 1u; }
                else if (MIX == MIX_VIADD_VIMNMX2) { x[c] = add2(x[c], y[c]); y[c] = max2(y[c], x[c]); }
                else if (MIX == HMNMX2_OP) { x[c] = hmax2(x[c], y[c]); y[c] = hmax2(y[c], b0); }
                else if (MIX == HADD2_OP) { x[c] = hadd2(x[c], y[c]); y[c] = hadd2(y[c], b0); }
                else if (MIX == MIX_DPX_HMNMX2) { x[c] = max2(add2(x[c], b0), y[c]); y[c] = hmax2(y[c], a0); }
                else if (MIX == MIX_DPX_HADD2) { x[c] = max2(add2(x[c], b0), y[c]); y[c] = hadd2(y[c], a0); }
                else if (MIX == MIX_HCELL) {
                    // the linear-gap cell in f16x2 (values kept in [1024, 2048): same bits as biased int16): 5 ops
                    uint32_t t = hadd2(y[c], a0);
                    uint32_t pre = hmax2(hmax2(t, x[c]), b0);
                    y[c] = x[c];
                    x[c] = hmax2(hadd2(x[(c + 1) % NCHAIN], b0), pre);
                }
                else if (MIX == MIX_DPX_CELL_AND_HCELL) {
                    // chains 0..4: the DPX cell (IMAD + VIMNMX3 + VIADDMNMX); chains 5..7: the f16x2 cell (5 ops)
                    if (c < 5) {
                        uint32_t t = y[c] * one + a0;
                        uint32_t pre = max2(max2(t, x[c]), b0);
                        y[c] = x[c];
                        x[c] = max2(add2(x[(c + 1) % 5], b0), pre);
                    } else {
                        uint32_t t = hadd2(y[c], a0);
                        uint32_t pre = hmax2(hmax2(t, x[c]), b0);
                        y[c] = x[c];
                        x[c] = hmax2(hadd2(x[5 + (c - 4) % 3], b0), pre);
                    }
                }
                else if (MIX == MIX_SW_CELL) {
                    // the three-op linear-gap cell: t = nw + s ; pre = max(w + g, t) ; h = max(n + g, pre)
                    uint32_t t = add2(y[c], a0);
                    uint32_t pre = max2(add2(x[c], b0), t);
                    y[c] = x[c];
                    x[c] = max2(add2(x[(c + 1) % NCHAIN], b0), pre);
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) acc ^= x[c] ^ y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - c0;
}

// dependent-issue latency: ONE warp, ONE chain of dependent instructions
enum LatOp { LAT_VIADDMNMX16 = 0, LAT_VIMNMX3_16, LAT_VIADDMNMX32, LAT_VIADDMNMX32_RELU, LAT_IMAD, LAT_IADD, LAT_SHFL, LAT_LDS, NLAT };
const char *kLatName[NLAT] = {"viaddmnmx_s16x2", "vimnmx3_s16x2", "viaddmnmx_s32", "viaddmnmx_s32_relu", "imad", "iadd3", "shfl_up", "lds"};

template <int OP>
__global__ void lat_kernel(uint32_t *out, long long *cyc, int n, uint32_t a0, uint32_t b0, uint32_t one)
{
    __shared__ uint32_t sm[64];
    sm[threadIdx.x] = (threadIdx.x + 1) & 31;
    sm[threadIdx.x + 32] = a0;
    __syncthreads();
    asm volatile("" : "+r"(a0), "+r"(b0), "+r"(one));
    uint32_t x = a0 + threadIdx.x;
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (OP == LAT_VIADDMNMX16) x = max2(add2(x, b0), a0);
            else if (OP == LAT_VIMNMX3_16) x = max2(max2(x, b0), a0);
            else if (OP == LAT_VIADDMNMX32) x = (uint32_t)__viaddmax_s32((int)x, (int)b0, (int)a0);
            else if (OP == LAT_VIADDMNMX32_RELU) x = (uint32_t)__viaddmax_s32_relu((int)x, (int)b0, (int)a0);
            else if (OP == LAT_IMAD) x = x * one + b0;
            else if (OP == LAT_IADD) x = x + b0;
            else if (OP == LAT_SHFL) x = __shfl_up_sync(0xffffffffu, x, 1);
            else if (OP == LAT_LDS) x = sm[x & 31];
        }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = c1 - c0;
}

template <int OP>
double lat_one(uint32_t *dout, long long *dcyc)
{
    const int n = 2000;
    lat_kernel<OP><<<1, 32>>>(dout, dcyc, n, 3, 0xfffcfffcu, 1);
    lat_kernel<OP><<<1, 32>>>(dout, dcyc, n, 3, 0xfffcfffcu, 1);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
    return (double)c / ((double)n * 32.0);
}
template <int OP>
void lat_all(uint32_t *dout, long long *dcyc, double *r)
{
    if constexpr (OP < NLAT) { r[OP] = lat_one<OP>(dout, dcyc); lat_all<OP + 1>(dout, dcyc, r); }
}

struct MixInfo { const char *name; int alu_ops; int total_ops; };
// ops per (chain, unroll) slot: ALU-pipe candidates vs all issued
const MixInfo kInfo[NMIX] = {
    {"viaddmnmx_s16x2", 1, 1}, {"vimnmx3_s16x2", 1, 1}, {"viadd_16x2", 1, 1}, {"iadd3_s32", 2, 2},
    {"imad_s32", 1, 1}, {"viaddmnmx_s32", 1, 1}, {"lop3", 2, 2}, {"viaddmnmx_s16x2+imad", 1, 2},
    {"viaddmnmx_s16x2+lds/8", 1, 1}, {"sw_cell_3op_s16x2", 3, 3},
    {"vimnmx_s16x2(2in)", 2, 2}, {"vimnmx_s16x2_relu(2in)", 2, 2}, {"viaddmnmx_s16x2_relu", 1, 1},
    {"viaddmnmx_s16x2+vimnmx2", 2, 2}, {"viaddmnmx_s16x2+viadd16x2", 2, 2}, {"vimnmx2+imad", 2, 2},
    {"prmt", 2, 2}, {"shf", 2, 2}, {"viaddmnmx_s16x2+shfl/8", 1, 1}, {"imnmx_s32(+iadd)", 3, 3},
    {"viadd16x2+vimnmx2", 2, 2},
    {"hmnmx2_f16x2", 2, 2}, {"hadd2_f16x2", 2, 2}, {"viaddmnmx_s16x2+hmnmx2", 1, 2}, {"viaddmnmx_s16x2+hadd2", 1, 2},
    {"sw_cell_5op_f16x2", 5, 5}, {"5x(imad+2dpx)+3x(5 f16x2 ops) per 8 cells", 4, 4},
};

template <int MIX>
double run_one(int iters, int sms, int clock_khz, uint32_t *dout, long long *dcyc, double *ms_out, double *mhz_out)
{
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench_kernel<MIX>, 256, 0);
    const int ctas = sms * occ;    // exactly one co-resident wave, so clock64 spans the whole run
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<MIX><<<ctas, 256>>>(dout, dcyc, iters, 3, 0xfffcfffcu, 1);    // warm-up (also ramps the clock)
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench_kernel<MIX><<<ctas, 256>>>(dout, dcyc, iters, 3, 0xfffcfffcu, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out = best;
    const double warp_instr = (double)ctas * 8 /*warps*/ * iters * UNROLL * NCHAIN * kInfo[MIX].total_ops;
    // real SM cycles of the last run: the longest-lived CTA spans the whole kernel (all CTAs
    // are co-resident, but the warp arbiter lets low-priority CTAs finish last: use max, not mean)
    static long long hc[4096];
    cudaMemcpy(hc, dcyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int k = 0; k < ctas; ++k) if ((double)hc[k] > cyc) cyc = (double)hc[k];
    *mhz_out = cyc / (best * 1e-3) / 1e6;   // effective SM clock during the run
    (void)clock_khz;
    return warp_instr / cyc / sms;          // warp-instructions per REAL clock per SM
}

template <int M>
void run_all(int iters, int sms, int clock_khz, uint32_t *dout, long long *dcyc, double *r, double *ms, double *mhz)
{
    if constexpr (M < NMIX) {
        r[M] = run_one<M>(iters, sms, clock_khz, dout, dcyc, &ms[M], &mhz[M]);
        run_all<M + 1>(iters, sms, clock_khz, dout, dcyc, r, ms, mhz);
    }
}

}  // namespace

extern "C" int swb_microbench_json(int device, int iters, char *buf, int buflen)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return -1;
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    uint32_t *dout;
    if (cudaMalloc(&dout, (size_t)p.multiProcessorCount * 8 * 256 * 4) != cudaSuccess) return -1;
    long long *dcyc;
    if (cudaMalloc(&dcyc, sizeof(long long) * 4096) != cudaSuccess) return -1;
    double r[NMIX], ms[NMIX], mhz[NMIX];
    run_all<0>(iters, p.multiProcessorCount, clock_khz, dout, dcyc, r, ms, mhz);
    double lat[NLAT];
    lat_all<0>(dout, dcyc, lat);
    cudaFree(dout); cudaFree(dcyc);
    std::string s = "{";
    char tmp[768];
    snprintf(tmp, sizeof tmp, "\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %.1f, \"iters\": %d, \"mixes\": {",
             p.name, p.multiProcessorCount, clock_khz / 1000.0, iters);
    s += tmp;
    for (int m = 0; m < NMIX; ++m) {
        snprintf(tmp, sizeof tmp,
                 "%s\"%s\": {\"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f, \"ms\": %.3f, \"sm_mhz\": %.0f}",
                 m ? ", " : "", kInfo[m].name, r[m], r[m] * 32.0, ms[m], mhz[m]);
        s += tmp;
    }
    s += "}, \"dependent_issue_latency_cycles\": {";
    for (int m = 0; m < NLAT; ++m) {
        snprintf(tmp, sizeof tmp, "%s\"%s\": %.2f", m ? ", " : "", kLatName[m], lat[m]);
        s += tmp;
    }
    s += "}, \"note\": \"rates are per REAL SM clock (clock64 inside the kernel); sm_mhz is the effective clock during that run\"}";
    if ((int)s.size() + 1 > buflen) return -2;
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

#ifdef SWB_MICROBENCH_MAIN
int main(int argc, char **argv)
{
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    static char buf[16384];
    int rc = swb_microbench_json(0, iters, buf, sizeof buf);
    if (rc) { fprintf(stderr, "microbench failed: %d\n", rc); return 1; }
    puts(buf);
    return 0;
}
#endif
