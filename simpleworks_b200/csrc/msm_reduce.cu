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
__global__ void __launch_bounds__(MSM_TAIL_THREADS) k_msm_gather(G1Xyzz* __restrict__ buckets, const G1Xyzz* __restrict__ partial,
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

// heavy buckets (more than MSM_GATHER_INLINE partial sums; a bucket that holds a large share of all points
// has one per range, i.e. up to ~10^5): two stages.  Stage 1 cuts every heavy bucket's partial sums into
// chunks of MSM_HEAVY_CHUNK, one block per chunk (strided sums, then a shared-memory tree) -> chunk sums;
// stage 2 adds the chunk sums of a bucket.  Every block walks the (short) heavy list itself to find where a
// bucket's chunks start, so no scan is needed.
__device__ __forceinline__ G1Xyzz block_sum(G1Xyzz* buf, const G1Xyzz* __restrict__ src, uint32_t lo, uint32_t hi) {
    const uint32_t t = threadIdx.x;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t p = lo + t; p < hi; p += MSM_HEAVY_THREADS) acc.add(src[p]);
    buf[t] = acc;
    __syncthreads();
    for (uint32_t d = MSM_HEAVY_THREADS >> 1; d > 0; d >>= 1) {
        if (t < d) {
            G1Xyzz a = buf[t];
            a.add(buf[t + d]);
            buf[t] = a;
        }
        __syncthreads();
    }
    const G1Xyzz r = buf[0];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) k_msm_heavy_chunks(G1Xyzz* __restrict__ chunk_sum,
                                                                       const G1Xyzz* __restrict__ partial,
                                                                       const uint32_t* __restrict__ pstart,
                                                                       const uint32_t* __restrict__ heavy) {
    extern __shared__ unsigned char smem_raw[];
    G1Xyzz* buf = reinterpret_cast<G1Xyzz*>(smem_raw);
    const uint32_t nheavy = heavy[0];
    uint32_t first = 0;                                   // index of the bucket's first chunk
    for (uint32_t h = 0; h < nheavy; h++) {
        const uint32_t b = heavy[1 + h];
        const uint32_t p0 = pstart[b], p1 = pstart[b + 1];
        const uint32_t nch = (p1 - p0 + MSM_HEAVY_CHUNK - 1) / MSM_HEAVY_CHUNK;
        // chunks are dealt round-robin over the grid, continuing across buckets
        for (uint32_t j = (blockIdx.x + gridDim.x - first % gridDim.x) % gridDim.x; j < nch; j += gridDim.x) {
            const uint32_t lo = p0 + j * MSM_HEAVY_CHUNK, hi = lo + MSM_HEAVY_CHUNK < p1 ? lo + MSM_HEAVY_CHUNK : p1;
            const G1Xyzz r = block_sum(buf, partial, lo, hi);
            if (threadIdx.x == 0) chunk_sum[first + j] = r;
        }
        first += nch;
    }
}
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) k_msm_heavy_finish(G1Xyzz* __restrict__ buckets,
                                                                       const G1Xyzz* __restrict__ chunk_sum,
                                                                       const uint32_t* __restrict__ pstart,
                                                                       const uint32_t* __restrict__ heavy) {
    extern __shared__ unsigned char smem_raw[];
    G1Xyzz* buf = reinterpret_cast<G1Xyzz*>(smem_raw);
    const uint32_t nheavy = heavy[0];
    uint32_t first = 0;
    for (uint32_t h = 0; h < nheavy; h++) {
        const uint32_t b = heavy[1 + h];
        const uint32_t nch = (pstart[b + 1] - pstart[b] + MSM_HEAVY_CHUNK - 1) / MSM_HEAVY_CHUNK;
        if (h % gridDim.x == blockIdx.x) {
            const G1Xyzz r = block_sum(buf, chunk_sum, first, first + nch);
            if (threadIdx.x == 0) buckets[b] = r;
        }
        first += nch;
    }
}

// ---- 5. bucket-set sums: R_w = sum_{k=1..B} k * B_k ------------------------------------------------
// Level 0 cuts the B buckets of a set into m = B / L segments: P_s = sum of the segment's buckets,
// W_s = sum (j_local + 1) B (running sums, one thread per segment), so that
//     R = sum_s W_s + L * sum_s s * P_s.
// The second term is taken bit by bit of the segment index: sum_s s P_s = sum_b 2^b T_b with
// T_b = sum of P_s over the segments whose index has bit b set -- plain sums, which reduce as parallel
// trees instead of serial running sums (a lone thread needs ~10 us per XYZZ addition, and a chain of
// levels of running sums cost 2-3 ms whatever the size).  k_msm_bit_sums: job 0 = sum W_s, job 1 + b =
// T_b, a few blocks per job; k_msm_bit_tree: the blocks' partial results of one job, then 2^b L by
// doublings; k_msm_bit_final: the jobs of one set.
__global__ void __launch_bounds__(MSM_TAIL_THREADS) k_msm_segments(G1Xyzz* __restrict__ seg_w, G1Xyzz* __restrict__ seg_p,
                                                       const G1Xyzz* __restrict__ buckets, uint32_t L, uint32_t nseg_total) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;   // global segment id (set-major); B is a multiple of L
    if (g >= nseg_total) return;
    const G1Xyzz* base = buckets + (size_t)g * L;
    G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity();
    for (uint32_t j = L; j-- > 0;) {
        running.add(base[j]);
        acc.add(running);
    }
    seg_p[g] = running;
    seg_w[g] = acc;
}

// (bucket shards also need the plain sum of all buckets: job `plain_job` = sum of every P_s, unweighted)
__global__ void __launch_bounds__(MSM_TAIL_THREADS) k_msm_bit_sums(G1Xyzz* __restrict__ out, const G1Xyzz* __restrict__ seg_w,
                                                                   const G1Xyzz* __restrict__ seg_p, uint32_t m, uint32_t per,
                                                                   uint32_t plain_job) {
    extern __shared__ unsigned char smem_raw[];
    G1Xyzz* buf = reinterpret_cast<G1Xyzz*>(smem_raw);
    const uint32_t t = threadIdx.x, blk = blockIdx.x, job = blockIdx.y, w = blockIdx.z;
    const uint32_t lo = blk * per, hi = lo + per < m ? lo + per : m;
    const G1Xyzz* src = (job == 0 ? seg_w : seg_p) + (size_t)w * m;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t s = lo + t; s < hi; s += MSM_TAIL_THREADS)
        if (job == 0 || job == plain_job || ((s >> (job - 1)) & 1u)) acc.add(src[s]);
    buf[t] = acc;
    __syncthreads();
    for (uint32_t d = MSM_TAIL_THREADS >> 1; d > 0; d >>= 1) {
        if (t < d) {
            G1Xyzz a = buf[t];
            a.add(buf[t + d]);
            buf[t] = a;
        }
        __syncthreads();
    }
    if (t == 0) out[((size_t)w * gridDim.y + job) * gridDim.x + blk] = buf[0];
}

// one warp per (job, set): sum the bpj partial results, then weight job 1 + b by 2^b * L
__global__ void __launch_bounds__(32) k_msm_bit_tree(G1Xyzz* __restrict__ val, const G1Xyzz* __restrict__ part, uint32_t bpj,
                                                      uint32_t log_L, uint32_t plain_job) {
    __shared__ G1Xyzz buf[32];
    const uint32_t i = threadIdx.x, job = blockIdx.x, w = blockIdx.y, njobs = gridDim.x;
    buf[i] = i < bpj ? part[((size_t)w * njobs + job) * bpj + i] : G1Xyzz::identity();
    __syncwarp();
    for (uint32_t d = 16; d > 0; d >>= 1) {
        if (i < d) {
            G1Xyzz a = buf[i];
            a.add(buf[i + d]);
            buf[i] = a;
        }
        __syncwarp();
    }
    if (i == 0) {
        G1Xyzz x = buf[0];
        if (job > 0 && job != plain_job)
            for (uint32_t k = 0; k < job - 1 + log_L; k++) x = x.dbl();
        val[(size_t)w * njobs + job] = x;
    }
}
// one warp per set: sum its (at most 32) weighted job values
// (the plain job, when present, is the last one: it goes to wins[MSM_MAX_WINDOWS + w] instead of into the sum)
__global__ void __launch_bounds__(32) k_msm_bit_final(G1Xyzz* __restrict__ wins, const G1Xyzz* __restrict__ val, uint32_t njobs,
                                                       uint32_t plain_job) {
    __shared__ G1Xyzz buf[32];
    const uint32_t i = threadIdx.x, w = blockIdx.x;
    buf[i] = (i < njobs && i != plain_job) ? val[(size_t)w * njobs + i] : G1Xyzz::identity();
    if (i == plain_job && i < njobs) wins[MSM_MAX_WINDOWS + w] = val[(size_t)w * njobs + i];
    __syncwarp();
    for (uint32_t d = 16; d > 0; d >>= 1) {
        if (i < d) {
            G1Xyzz a = buf[i];
            a.add(buf[i + d]);
            buf[i] = a;
        }
        __syncwarp();
    }
    if (i == 0) wins[w] = buf[0];
}

int msm_launch_gather(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf) {
    k_msm_partial_bounds<<<(pl.nb + 1 + 255) / 256, 256, 0, c->stream>>>(bf.pstart, bf.pkey, bf.range_off + pl.nranges, pl.nb, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_partial_bounds");
    k_msm_gather<<<(pl.nb + MSM_TAIL_THREADS - 1) / MSM_TAIL_THREADS, MSM_TAIL_THREADS, 0, c->stream>>>(bf.buckets, bf.partial, bf.pstart, pl.nb, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_gather");
    // chunk sums: at most pcap / MSM_HEAVY_CHUNK full chunks plus one partly filled chunk per heavy bucket
    // (a heavy bucket has > MSM_GATHER_INLINE partial sums, so there are fewer than pcap / MSM_GATHER_INLINE)
    const size_t max_chunks = (size_t)pl.pcap / MSM_HEAVY_CHUNK + (size_t)pl.pcap / MSM_GATHER_INLINE + 2;
    G1Xyzz* chunk_sum = (G1Xyzz*)get_scratch(c, "msm_heavy_chunks", max_chunks * sizeof(G1Xyzz));
    if (!chunk_sum) return SWB_ENOMEM;
    const size_t smem = MSM_HEAVY_THREADS * sizeof(G1Xyzz);
    SWB_CUDA(c, cudaFuncSetAttribute(k_msm_heavy_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SWB_CUDA(c, cudaFuncSetAttribute(k_msm_heavy_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_msm_heavy_chunks<<<c->sm_count * 2, MSM_HEAVY_THREADS, smem, c->stream>>>(chunk_sum, bf.partial, bf.pstart, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_heavy_chunks");
    k_msm_heavy_finish<<<c->sm_count, MSM_HEAVY_THREADS, smem, c->stream>>>(bf.buckets, chunk_sum, bf.pstart, bf.heavy);
    SWB_LAUNCH_CHECK(c, "k_msm_heavy_finish");
    return SWB_OK;
}

int msm_launch_reduce(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf) {
    const uint32_t B = pl.B, nwin = (uint32_t)pl.nwin;
    const uint32_t L = msm_reduce_seg_len(nwin, B);
    uint32_t log_L = 0;
    while ((1u << log_L) < L) log_L++;
    const uint32_t m = B / L;                             // segments per set (a power of two)
    uint32_t nbits = 0;
    while ((1u << nbits) < m) nbits++;
    const bool sharded = pl.shard_shift != 0;             // then the host also needs the plain sum of our buckets
    const uint32_t plain_job = sharded ? 1 + nbits : 0xffffffffu;
    const uint32_t njobs = 1 + nbits + (sharded ? 1 : 0); // <= 1 + 22 + 1
    G1Xyzz *seg_w = bf.seg, *seg_p = bf.seg + (size_t)nwin * m;
    k_msm_segments<<<(nwin * m + MSM_TAIL_THREADS - 1) / MSM_TAIL_THREADS, MSM_TAIL_THREADS, 0, c->stream>>>(seg_w, seg_p, bf.buckets, L, nwin * m);
    SWB_LAUNCH_CHECK(c, "k_msm_segments");
    // a few blocks per job so that no thread adds more than ~8 segments serially
    uint32_t bpj = 1;
    while (bpj < 32 && (size_t)bpj * MSM_TAIL_THREADS * 8 < m) bpj <<= 1;
    const uint32_t per = (m + bpj - 1) / bpj;
    G1Xyzz* part = bf.seg2;                               // [nwin][njobs][bpj], then [nwin][njobs] values
    G1Xyzz* val = part + (size_t)nwin * njobs * bpj;
    const size_t smem = MSM_TAIL_THREADS * sizeof(G1Xyzz);
    k_msm_bit_sums<<<dim3(bpj, njobs, nwin), MSM_TAIL_THREADS, smem, c->stream>>>(part, seg_w, seg_p, m, per, plain_job);
    SWB_LAUNCH_CHECK(c, "k_msm_bit_sums");
    k_msm_bit_tree<<<dim3(njobs, nwin), 32, 0, c->stream>>>(val, part, bpj, log_L, plain_job);
    SWB_LAUNCH_CHECK(c, "k_msm_bit_tree");
    k_msm_bit_final<<<nwin, 32, 0, c->stream>>>(bf.wins, val, njobs, plain_job);
    SWB_LAUNCH_CHECK(c, "k_msm_bit_final");
    return SWB_OK;
}

}  // namespace swb
