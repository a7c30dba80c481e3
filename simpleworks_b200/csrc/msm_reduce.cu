// MSM stage 5: partial sums -> buckets, then sum_k k * B_k per window.  Cold relative to the
// accumulation kernel, so the field product is called out of line here (smaller code, faster build).
#define SWB_FP_NOINLINE_MUL
#include "msm_common.cuh"

namespace swb {

// pstart[b] = first partial sum whose bucket id is >= b (partial keys are sorted), b in [0, nb]
__global__ void k_msm_partial_bounds(uint32_t* __restrict__ pstart, const uint32_t* __restrict__ pkey,
                                     const uint32_t* __restrict__ np_ptr, uint32_t nb, uint32_t* __restrict__ heavy) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) heavy[0] = 0;
    if (b > nb) return;
    uint32_t lo = 0, hi = *np_ptr;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pkey[mid] < b) lo = mid + 1; else hi = mid;
    }
    pstart[b] = lo;
}

// one thread per bucket: add its partial sums; buckets with many of them are queued for the
// block-wide kernel instead
__global__ void __launch_bounds__(128) k_msm_gather(G1Xyzz* __restrict__ buckets, const G1Xyzz* __restrict__ partial,
                                                     const uint32_t* __restrict__ pstart, uint32_t nb,
                                                     uint32_t* __restrict__ heavy) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t p0 = pstart[b], p1 = pstart[b + 1];
    if (p1 - p0 > (uint32_t)MSM_GATHER_INLINE) {
        const uint32_t slot = atomicAdd(&heavy[0], 1u);
        heavy[1 + slot] = b;
        return;
    }
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t p = p0; p < p1; p++) acc.add(partial[p]);
    buckets[b] = acc;
}

// heavy buckets: one block per bucket, strided sums then a shared-memory tree
__global__ void __launch_bounds__(MSM_RED_THREADS) k_msm_gather_heavy(G1Xyzz* __restrict__ buckets,
                                                                       const G1Xyzz* __restrict__ partial,
                                                                       const uint32_t* __restrict__ pstart,
                                                                       const uint32_t* __restrict__ heavy) {
    extern __shared__ unsigned char smem_raw[];
    G1Xyzz* buf = reinterpret_cast<G1Xyzz*>(smem_raw);
    const uint32_t nheavy = heavy[0], t = threadIdx.x;
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const uint32_t b = heavy[1 + h];
        const uint32_t p0 = pstart[b], p1 = pstart[b + 1];
        G1Xyzz acc = G1Xyzz::identity();
        for (uint32_t p = p0 + t; p < p1; p += MSM_RED_THREADS) acc.add(partial[p]);
        buf[t] = acc;
        __syncthreads();
        for (uint32_t d = MSM_RED_THREADS >> 1; d > 0; d >>= 1) {
            if (t < d) {
                G1Xyzz a = buf[t];
                a.add(buf[t + d]);
                buf[t] = a;
            }
            __syncthreads();
        }
        if (t == 0) buckets[b] = buf[0];
        __syncthreads();
    }
}

// ---- 5a. per-segment running sums: S = sum B_j, Wt = sum (j_local+1) B_j -------------------------
__global__ void __launch_bounds__(128) k_msm_segments(G1Xyzz* __restrict__ seg_s, G1Xyzz* __restrict__ seg_w,
                                                       const G1Xyzz* __restrict__ buckets, uint32_t B, uint32_t L,
                                                       uint32_t nseg_total) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;   // global segment id (window-major)
    if (g >= nseg_total) return;
    const G1Xyzz* base = buckets + (size_t)g * L;          // B is a multiple of L
    (void)B;
    G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity();
    for (uint32_t j = L; j-- > 0;) {
        running.add(base[j]);
        acc.add(running);
    }
    seg_s[g] = running;
    seg_w[g] = acc;
}

// ---- 5b. one block per window: total = sum_s Wt_s + L * sum_{j>=1} suffix_j(S) --------------
__global__ void __launch_bounds__(MSM_RED_THREADS) k_msm_window_reduce(G1Xyzz* __restrict__ win_sums,
                                                                        const G1Xyzz* __restrict__ seg_s,
                                                                        const G1Xyzz* __restrict__ seg_w, uint32_t nseg,
                                                                        uint32_t log_L) {
    extern __shared__ unsigned char smem_raw[];
    G1Xyzz* bufA = reinterpret_cast<G1Xyzz*>(smem_raw);
    G1Xyzz* bufB = bufA + MSM_RED_THREADS;
    const uint32_t t = threadIdx.x, w = blockIdx.x;
    G1Xyzz mine = t < nseg ? seg_s[(size_t)w * nseg + t] : G1Xyzz::identity();
    // inclusive suffix scan (Hillis-Steele): suffix_t = sum_{s >= t} S_s
    bufA[t] = mine;
    __syncthreads();
    G1Xyzz* src = bufA;
    G1Xyzz* dst = bufB;
    for (uint32_t d = 1; d < MSM_RED_THREADS; d <<= 1) {
        G1Xyzz v = src[t];
        if (t + d < MSM_RED_THREADS) v.add(src[t + d]);
        dst[t] = v;
        __syncthreads();
        G1Xyzz* tmp = src; src = dst; dst = tmp;
    }
    G1Xyzz v = src[t];
    __syncthreads();
    if (t == 0) v = G1Xyzz::identity();                     // j >= 1 only
    for (uint32_t i = 0; i < log_L; i++) v = v.dbl();        // times L
    if (t < nseg) v.add(seg_w[(size_t)w * nseg + t]);
    // tree reduction
    src[t] = v;
    __syncthreads();
    for (uint32_t d = MSM_RED_THREADS >> 1; d > 0; d >>= 1) {
        if (t < d) {
            G1Xyzz a = src[t];
            a.add(src[t + d]);
            src[t] = a;
        }
        __syncthreads();
    }
    if (t == 0) win_sums[w] = src[0];
}



int msm_launch_gather(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf) {
    k_msm_partial_bounds<<<(pl.nb + 1 + 255) / 256, 256, 0, c->stream>>>(bf.pstart, bf.pkey, bf.range_off + pl.nranges, pl.nb, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_partial_bounds");
    k_msm_gather<<<(pl.nb + 127) / 128, 128, 0, c->stream>>>(bf.buckets, bf.partial, bf.pstart, pl.nb, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_gather");
    const size_t smem = MSM_RED_THREADS * sizeof(G1Xyzz);
    SWB_CUDA(c, cudaFuncSetAttribute(k_msm_gather_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_msm_gather_heavy<<<c->sm_count, MSM_RED_THREADS, smem, c->stream>>>(bf.buckets, bf.partial, bf.pstart, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_gather_heavy");
    return SWB_OK;
}

int msm_launch_reduce(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf) {
    const uint32_t B = pl.B;
    uint32_t nseg = B < (uint32_t)MSM_RED_THREADS ? B : (uint32_t)MSM_RED_THREADS;
    uint32_t L = B / nseg, log_L = 0;
    while ((1u << log_L) < L) log_L++;
    const uint32_t nseg_total = nseg * (uint32_t)pl.nwin;
    G1Xyzz* seg = bf.seg;
    k_msm_segments<<<(nseg_total + 127) / 128, 128, 0, c->stream>>>(seg, seg + nseg_total, bf.buckets, B, L, nseg_total);
    SWB_LAUNCH_CHECK(c, "k_msm_segments");
    const size_t red_smem = 2 * MSM_RED_THREADS * sizeof(G1Xyzz);
    SWB_CUDA(c, cudaFuncSetAttribute(k_msm_window_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
    k_msm_window_reduce<<<pl.nwin, MSM_RED_THREADS, red_smem, c->stream>>>(bf.wins, seg, seg + nseg_total, nseg, log_L);
    SWB_LAUNCH_CHECK(c, "k_msm_window_reduce");
    return SWB_OK;
}

}  // namespace swb
