// swb_fill.cu -- score-matrix fill of the short-read path (sm_100a, DPX s16x2).
//
// Computes, for every (read, reference) pair of a batch, the maximum of
//   H(i,j) = max(0, H(i,j-1)+gap, H(i-1,j)+gap, H(i-1,j-1)+match|mismatch)
// which is the score recurrence of GetCellScore.call / InsDelScore.call /
// AlignmentScore.call (reference SmithWaterman.java:217-252, :277-280, :309-318) as
// iterated by ScoreMatrix.call (:157-187).  Scores only: the tie rule of the ">="
// cascade affects alignment types, never H, so the fill needs no direction state.
// The traceback kernels (swb_trace.cu) restart from the checkpoints written here.
//
// Per cell pair (two reads packed in s16x2):  VIADDMNMX.RELU  x   = max(NW + s, 0)
//                                             VIADDMNMX       pre = max(W + gap, x)
//                                             VIADDMNMX       H   = max(N + gap, pre)
// plus half a VIMNMX3 for the running tile maximum.  The dependent chain down a lane's
// K rows is one VIADDMNMX per row; x and pre of all rows are independent.
#include "swb_internal.h"
#include "swb_device.cuh"

namespace swb {

template <int K>
__global__ void __launch_bounds__(256, 2) fill_kernel(const BatchParams P, int quads_per_cta)
{
    using G = Geo<K>;
    extern __shared__ __align__(16) uint32_t prof[];            // [4 codes][GL lanes][KS]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int t = lane & (GL - 1), g = lane >> 3;
    const int rp = blockIdx.x % P.n_rp;                           // longest references first
    const int chunk = blockIdx.x / P.n_rp;

    // ---- query profile of this CTA's read pair -------------------------------------
    const int ra = P.rp_reads[2 * rp], rb = P.rp_reads[2 * rp + 1];
    const int64_t offa = P.read_off[ra];
    const int ma = (int)(P.read_off[ra + 1] - offa);
    const int64_t offb = rb >= 0 ? P.read_off[rb] : 0;
    const int mb = rb >= 0 ? (int)(P.read_off[rb + 1] - offb) : 0;
    for (int idx = threadIdx.x; idx < 4 * GL * K; idx += blockDim.x) {
        const int c = idx / (GL * K), rem = idx - c * GL * K;
        const int tt = rem / K, r = rem - tt * K;
        const int row = tt * K + r;
        int lo = S_PAD, hi = S_PAD;
        if (row < ma) lo = (P.read_codes[offa + row] == c) ? P.match : P.mismatch;
        if (row < mb) hi = (P.read_codes[offb + row] == c) ? P.match : P.mismatch;
        prof[c * G::CSTRIDE + tt * G::KS + r] = pack2(lo, hi);
    }
    __syncthreads();

    const uint32_t g2 = pack2(P.gap, P.gap);
    uint32_t zero = 0;
    asm volatile("" : "+r"(zero));                                // a register zero: keeps PRMT out of the loop
    const uint32_t lmask = t ? 0xffffffffu : 0u;
    const int n_quads = (P.n_refs + 3) >> 2;
    const int q_begin = chunk * quads_per_cta;
    const int q_end = min(n_quads, q_begin + quads_per_cta);
    const int my_prof = t * G::KS;

    for (int q = q_begin + warp; q < q_end; q += nwarps) {
        const int ref = 4 * q + g;                                // sorted reference index of this group
        const bool has_ref = ref < P.n_refs;
        const int n_g = has_ref ? P.ref_len[ref] : 0;
        const int nmax = P.ref_len[4 * q];                        // sorted descending
        const int nmin = (4 * q + 3 < P.n_refs) ? P.ref_len[4 * q + 3] : 0;
        const uint32_t *wp = P.ref_words + (has_ref ? P.ref_word_off[ref] : 0);
        const int64_t blk0 = (int64_t)rp * P.blocks_per_rp + (has_ref ? P.ref_blk_off[ref] : 0);
        const int my_steps = has_ref ? n_g + GL - 1 : 0;          // steps this group needs

        uint32_t H[K];
#pragma unroll
        for (int r = 0; r < K; ++r) H[r] = 0;
        uint32_t diag = 0, tmax = 0, gmax = 0, wprev = 0;

        const int nsteps = nmax + GL - 1;
        for (int s0 = 0; s0 < nsteps; s0 += 16) {
            const uint32_t wnew = (s0 < n_g) ? __ldg(wp + (s0 >> 4)) : 0u;
            // window of this lane's 16 codes for steps s0..s0+15: columns (0-based) s0-t ..
            const uint32_t win = __funnelshift_rc(wprev, wnew, 32 - 2 * t);
            wprev = wnew;
            const bool fast = (s0 >= GL - 1) && (s0 + 16 <= nmin);

            if (fast) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t top = __shfl_up_sync(0xffffffffu, H[K - 1], 1) & lmask;
                    const uint32_t c = (win >> (2 * u)) & 3u;
                    uint32_t sv[G::KP];
                    load_profile<G::KP>(prof + c * G::CSTRIDE + my_prof, sv);
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
                    tmax = colmax<K>(tmax, H);
                }
            } else {
#pragma unroll 1
                for (int u = 0; u < 16; ++u) {
                    const int s = s0 + u;
                    const uint32_t top = __shfl_up_sync(0xffffffffu, H[K - 1], 1) & lmask;
                    const uint32_t c = (win >> (2 * u)) & 3u;
                    const bool valid = (s >= t) && (s < n_g + t);        // column j = s-t+1 in [1, n_g]
                    if (valid) {
                        uint32_t sv[G::KP];
                        load_profile<G::KP>(prof + c * G::CSTRIDE + my_prof, sv);
                        uint32_t nw = diag, nn = top;
#pragma unroll
                        for (int r = 0; r < K; ++r) {
                            const uint32_t x = viaddmax_relu(nw, sv[r], zero);
                            const uint32_t pre = viaddmax(H[r], g2, x);
                            nw = H[r];
                            H[r] = viaddmax(nn, g2, pre);
                            nn = H[r];
                        }
                        tmax = colmax<K>(tmax, H);
                    }
                    diag = top;
                }
            }

            // ---- end of a checkpoint block: state before step s0+16, tile max of the block
            const int s_next = s0 + 16;
            if ((s_next % CB) == 0) {
                const int b = s_next / CB;                       // block that starts at s_next
                if (s_next - CB < my_steps) {                     // block b-1 exists for this group
                    P.tmx[(blk0 + b - 1) * GL + t] = tmax;
                    gmax = vmax2(gmax, tmax);
                    tmax = 0;
                }
                if (s_next < my_steps) {                          // block b exists: checkpoint it
                    uint32_t *ck = P.ck + (blk0 + b) * (int64_t)(G::KW * GL) + t * 4;
                    store_checkpoint<K>(ck, H, diag);
                }
            }
        }
        // last (partial) block
        {
            const int s_end = ((nsteps + 15) >> 4) << 4;          // first step not executed
            if ((s_end % CB) != 0) {
                const int b_last = s_end / CB;
                if (b_last * CB < my_steps) {
                    P.tmx[(blk0 + b_last) * GL + t] = tmax;
                    gmax = vmax2(gmax, tmax);
                }
            }
        }
        // ---- pair scores: maximum over the group's 8 lanes, both halves ---------------
        gmax = vmax2(gmax, __shfl_xor_sync(0xffffffffu, gmax, 1));
        gmax = vmax2(gmax, __shfl_xor_sync(0xffffffffu, gmax, 2));
        gmax = vmax2(gmax, __shfl_xor_sync(0xffffffffu, gmax, 4));
        if (t == 0 && has_ref) {
            const int64_t ro = P.ref_orig[ref];
            P.scores[ro * P.n_reads + ra] = (int)(int16_t)(gmax & 0xffffu);
            if (rb >= 0) P.scores[ro * P.n_reads + rb] = (int)(int16_t)(gmax >> 16);
        }
    }
}

template <int K>
static cudaError_t launch_fill_k(const BatchParams &P, int sm_count, cudaStream_t st)
{
    using G = Geo<K>;
    const int n_quads = (P.n_refs + 3) / 4;
    // enough CTAs to fill the machine several times over, at least one quad per warp
    const int warps = 8;
    int quads_per_cta = warps * 4;
    while (quads_per_cta > warps && (int64_t)P.n_rp * ((n_quads + quads_per_cta - 1) / quads_per_cta) < 4LL * sm_count)
        quads_per_cta -= warps;
    const int chunks = (n_quads + quads_per_cta - 1) / quads_per_cta;
    const size_t smem = G::PROF_WORDS * sizeof(uint32_t);
    fill_kernel<K><<<dim3((unsigned)(chunks * P.n_rp)), dim3(warps * 32), smem, st>>>(P, quads_per_cta);
    return cudaGetLastError();
}

cudaError_t launch_fill(int K, const BatchParams &P, int sm_count, cudaStream_t st)
{
    switch (K) {
        case 4:  return launch_fill_k<4>(P, sm_count, st);
        case 8:  return launch_fill_k<8>(P, sm_count, st);
        case 13: return launch_fill_k<13>(P, sm_count, st);
        case 16: return launch_fill_k<16>(P, sm_count, st);
        case 19: return launch_fill_k<19>(P, sm_count, st);
        case 25: return launch_fill_k<25>(P, sm_count, st);
        case 32: return launch_fill_k<32>(P, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace swb
