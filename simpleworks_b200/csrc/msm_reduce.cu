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

// ---- 5. window sums: R_w = sum_{k=1..B} k * B_k, as a multi-level segmented reduction -------------
// Level 0 cuts the B buckets of a window into segments of L: S_s = sum B, C_s = sum (j_local+1) B.
// Then R = sum_s C_s + L * sum_s s * S_s, which has the same shape one level up: grouping G
// consecutive segments, S'_g = sum_j S_{gG+j},  C'_g = sum_j C_{gG+j} + M * sum_j j * S_{gG+j}  with
// M the product of the segment lengths below (a power of two: M * x is log2(M) doublings), and
// R = sum_g C'_g + (M G) * sum_g g * S'_g.  Every level is one thread per group; the last level
// leaves R in C_0.  Bucket counts up to 2^22 per window stay parallel this way.
__global__ void __launch_bounds__(128) k_msm_segments(G1Xyzz* __restrict__ seg_c, G1Xyzz* __restrict__ seg_s,
                                                       const G1Xyzz* __restrict__ buckets, uint32_t L, uint32_t nseg_total) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;   // global segment id (window-major); B is a multiple of L
    if (g >= nseg_total) return;
    const G1Xyzz* base = buckets + (size_t)g * L;
    G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity();
    for (uint32_t j = L; j-- > 0;) {
        running.add(base[j]);
        acc.add(running);
    }
    seg_s[g] = running;
    seg_c[g] = acc;
}

__global__ void __launch_bounds__(128) k_msm_reduce_level(G1Xyzz* __restrict__ out_c, G1Xyzz* __restrict__ out_s,
                                                           const G1Xyzz* __restrict__ in_c, const G1Xyzz* __restrict__ in_s,
                                                           uint32_t m, uint32_t m_out, uint32_t G, uint32_t log_M, uint32_t nwin) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m_out * nwin) return;
    const uint32_t w = t / m_out, g = t % m_out;
    const size_t base = (size_t)w * m + (size_t)g * G;
    const uint32_t cnt = (g + 1) * G <= m ? G : m - g * G;
    G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity(), csum = G1Xyzz::identity();
    for (uint32_t j = cnt; j-- > 1;) {                     // sum_j j * S_j as a running sum
        running.add(in_s[base + j]);
        acc.add(running);
    }
    running.add(in_s[base]);
    for (uint32_t j = 0; j < cnt; j++) csum.add(in_c[base + j]);
    for (uint32_t i = 0; i < log_M; i++) acc = acc.dbl();
    csum.add(acc);
    out_s[(size_t)w * m_out + g] = running;
    out_c[(size_t)w * m_out + g] = csum;
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
    const uint32_t B = pl.B, nwin = (uint32_t)pl.nwin;
    const uint32_t L = msm_reduce_seg_len(nwin, B);
    uint32_t log_M = 0;
    while ((1u << log_M) < L) log_M++;
    uint32_t m = B / L;                                   // segments per window
    G1Xyzz *cur_c = bf.seg, *cur_s = bf.seg + (size_t)nwin * m;
    G1Xyzz *alt_c = bf.seg2, *alt_s = nullptr;
    k_msm_segments<<<(nwin * m + 127) / 128, 128, 0, c->stream>>>(cur_c, cur_s, bf.buckets, L, nwin * m);
    SWB_LAUNCH_CHECK(c, "k_msm_segments");
    while (m > 1) {
        // a group costs its thread ~3 G serial additions: short groups while the level has few threads
        uint32_t G = MSM_SEG_LEN;
        while (G > (uint32_t)MSM_GROUP_MIN && (size_t)nwin * ((m + G - 1) / G) < 8192) G >>= 1;
        uint32_t log_G = 0;
        while ((1u << log_G) < G) log_G++;
        const uint32_t m_out = (m + G - 1) / G;
        alt_s = alt_c + (size_t)nwin * m_out;
        k_msm_reduce_level<<<(nwin * m_out + 127) / 128, 128, 0, c->stream>>>(alt_c, alt_s, cur_c, cur_s, m, m_out, G, log_M, nwin);
        SWB_LAUNCH_CHECK(c, "k_msm_reduce_level");
        G1Xyzz* old = cur_c;                              // ping-pong: the old input region is free again
        cur_c = alt_c;
        cur_s = alt_s;
        alt_c = old;
        m = m_out;
        log_M += log_G;
    }
    // m == 1: window w's result is cur_c[w]
    SWB_CUDA(c, cudaMemcpyAsync(bf.wins, cur_c, sizeof(G1Xyzz) * nwin, cudaMemcpyDeviceToDevice, c->stream));
    return SWB_OK;
}

}  // namespace swb
