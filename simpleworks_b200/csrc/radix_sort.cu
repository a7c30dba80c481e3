// Hand-written device primitives for the MSM's "sort scalars into buckets" stage: a stable
// segmented LSD radix sort of (bucket id, point index) pairs and an exclusive prefix sum.
// (arkworks has no such stage: VariableBaseMSM scatters into per-window bucket arrays on the CPU --
// ark-ec 0.3 msm/variable_base.rs; on the GPU the pairs are grouped by bucket first so that
// accumulation reads each bucket's points as one contiguous run.)
//
// The pairs arrive window-major (one segment of n pairs per window), so only the c-bit bucket id
// inside each segment is sorted: passes of up to 10 key bits -- 1024 bins, affordable because a
// B200 CTA can hold the per-warp cursors (32 KB) and a staging tile (64 KB) in shared memory --
// i.e. two passes for every window width this library uses.
// A block owns a tile of 8192 consecutive pairs of one segment, a warp a 1024-pair slice of it.
//   k_rs_hist     per-block digit histogram (shared-memory atomics) -> hist[segment][bin][tile]
//   exclusive scan of hist = where each tile's run of every digit starts in the output
//   k_rs_scatter  per-warp digit counts -> cursors; each warp walks its slice 32 pairs at a time and
//                 ranks equal digits with __match_any_sync (lane order = input order, so the pass is
//                 stable); pairs are first placed in tile-sorted order in shared memory and then
//                 written out by consecutive threads, so stores to a digit's run are contiguous
// All traffic is 4- and 8-byte streams: the stage is HBM-bound (20 B moved per pair and pass).
#include "ctx.hpp"
#include "radix_sort.hpp"

namespace swb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 32;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // 8192 pairs per block
constexpr int RS_SLICE = 32 * RS_ROUNDS;            // 1024 pairs per warp
constexpr int RS_MAX_BITS = 10;
constexpr int RS_MAX_BINS = 1 << RS_MAX_BITS;

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive scan ------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block(uint32_t* __restrict__ data, uint32_t* __restrict__ block_sums, size_t n) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? data[base + i] : 0u;
        sum += v[i];
    }
    // inclusive scan of the per-thread sums inside the warp
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const uint32_t t = warp_tot[w];
        if ((uint32_t)w < wid) warp_off += t;
        total += t;
    }
    uint32_t run = warp_off + incl - sum;           // exclusive prefix of this thread's first item
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (block_sums && threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_scan_add(uint32_t* __restrict__ data, const uint32_t* __restrict__ block_off, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_off[i / SCAN_TILE];
}

static int scan_rec(swb_ctx* c, uint32_t* data, size_t n, uint32_t* scratch) {
    const size_t nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nblk <= 1) {
        k_scan_block<<<1, SCAN_THREADS, 0, c->stream>>>(data, nullptr, n);
        SWB_LAUNCH_CHECK(c, "k_scan_block");
        return SWB_OK;
    }
    k_scan_block<<<(unsigned)nblk, SCAN_THREADS, 0, c->stream>>>(data, scratch, n);
    SWB_LAUNCH_CHECK(c, "k_scan_block");
    int rc = scan_rec(c, scratch, nblk, scratch + nblk);
    if (rc != SWB_OK) return rc;
    k_scan_add<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(data, scratch, n);
    SWB_LAUNCH_CHECK(c, "k_scan_add");
    return SWB_OK;
}

int exclusive_scan_u32(swb_ctx* c, uint32_t* data, size_t n) {
    if (n == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    size_t need = 0, m = n;
    while (m > SCAN_TILE) { m = (m + SCAN_TILE - 1) / SCAN_TILE; need += m; }
    uint32_t* scratch = (uint32_t*)get_scratch(c, "scan_sums", (need + 64) * sizeof(uint32_t));
    if (!scratch) return SWB_ENOMEM;
    return scan_rec(c, data, n, scratch);
}

// ---- radix sort passes ---------------------------------------------------------------------------
struct RsGeom {
    size_t seg_len;        // pairs per segment (window)
    uint32_t tiles_per_seg;
    uint32_t shift, bits;
};
__device__ __forceinline__ void rs_tile_range(const RsGeom& g, size_t* lo, size_t* hi, uint32_t* seg, uint32_t* tile) {
    *seg = blockIdx.x / g.tiles_per_seg;
    *tile = blockIdx.x % g.tiles_per_seg;
    *lo = (size_t)*seg * g.seg_len + (size_t)*tile * RS_TILE;
    const size_t end = (size_t)(*seg + 1) * g.seg_len;
    *hi = *lo + RS_TILE < end ? *lo + RS_TILE : end;
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(uint32_t* __restrict__ hist, const uint32_t* __restrict__ keys, RsGeom g) {
    __shared__ uint32_t h[RS_MAX_BINS];
    const uint32_t nbins = 1u << g.bits, mask = nbins - 1u;
    for (uint32_t i = threadIdx.x; i < nbins; i += RS_THREADS) h[i] = 0;
    __syncthreads();
    size_t lo, hi;
    uint32_t seg, tile;
    rs_tile_range(g, &lo, &hi, &seg, &tile);
    for (size_t i = lo + threadIdx.x; i < hi; i += RS_THREADS) atomicAdd(&h[(keys[i] >> g.shift) & mask], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += RS_THREADS)
        hist[((size_t)seg * nbins + b) * g.tiles_per_seg + tile] = h[b];
}

__global__ void __launch_bounds__(RS_THREADS, 2) k_rs_scatter(uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                            const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                            const uint32_t* __restrict__ hist, RsGeom g) {
    extern __shared__ uint32_t rs_smem[];
    const uint32_t nbins = 1u << g.bits, mask = nbins - 1u;
    uint32_t* cur = rs_smem;                          // [RS_WARPS][nbins] per-warp digit counts, then cursors
    uint32_t* lstart = cur + RS_WARPS * nbins;        // [nbins] tile-local start of each digit
    uint32_t* delta = lstart + nbins;                 // [nbins] output position minus tile-sorted position
    uint32_t* stage_k = delta + nbins;                // [RS_TILE]
    uint32_t* stage_v = stage_k + RS_TILE;            // [RS_TILE]
    __shared__ uint32_t warp_tot[RS_WARPS];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    size_t lo, hi;
    uint32_t seg, tile;
    rs_tile_range(g, &lo, &hi, &seg, &tile);
    const size_t slice = lo + (size_t)wid * RS_SLICE + lane;
    const int vr = slice < hi ? (int)((hi - slice + 31) / 32) : 0;   // this lane's pairs are rounds 0 .. vr-1
    const uint32_t* kp = keys_in + slice;
    const uint32_t* vp = vals_in + slice;
    // the slice's keys stay in registers for the whole kernel: all loads are in flight at once
    uint32_t k[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) k[r] = r < vr ? kp[r * 32] : 0u;
    for (uint32_t i = threadIdx.x; i < RS_WARPS * nbins; i += RS_THREADS) cur[i] = 0;
    __syncthreads();
    // a. rank every pair among the equal digits of its warp's slice (input order) and count digits:
    //    the lowest lane of each group of equal digits bumps the warp's counter once for the group
    uint32_t* wcur = cur + wid * nbins;
    uint32_t rk[RS_ROUNDS / 2];                       // two 16-bit ranks per register
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const bool valid = r < vr;
        const uint32_t d = valid ? ((k[r] >> g.shift) & mask) : (RS_MAX_BINS + lane);   // padding lanes match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = __popc(peers & lt_mask);
        uint32_t old = 0;
        if (valid && below == 0) old = atomicAdd(&wcur[d], (uint32_t)__popc(peers));
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        const uint32_t rank = old + below;
        if (r & 1) rk[r / 2] |= rank << 16;
        else rk[r / 2] = rank;
    }
    __syncthreads();
    // b. per digit: prefix over warps, then a block-wide exclusive scan of the digit totals
    const uint32_t bpt = nbins / RS_THREADS ? nbins / RS_THREADS : 1;   // bins per thread (contiguous)
    uint32_t my_tot = 0;
    for (uint32_t j = 0; j < bpt; j++) {
        const uint32_t b = threadIdx.x * bpt + j;
        if (b >= nbins) break;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t t = cur[w * nbins + b];
            cur[w * nbins + b] = run;
            run += t;
        }
        lstart[b] = run;                              // digit total for now
        my_tot += run;
    }
    uint32_t incl = my_tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    // the slice's values: in flight while the scan finishes
    uint32_t v[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) v[r] = r < vr ? vp[r * 32] : 0u;
    __syncthreads();
    uint32_t off = incl - my_tot;
    for (uint32_t w = 0; w < wid; w++) off += warp_tot[w];
    for (uint32_t j = 0; j < bpt; j++) {
        const uint32_t b = threadIdx.x * bpt + j;
        if (b >= nbins) break;
        const uint32_t t = lstart[b];
        lstart[b] = off;
        delta[b] = hist[((size_t)seg * nbins + b) * g.tiles_per_seg + tile] - off;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) cur[w * nbins + b] += off;
        off += t;
    }
    __syncthreads();
    // c. place the pairs at their tile-sorted position
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        if (r < vr) {
            const uint32_t d = (k[r] >> g.shift) & mask;
            const uint32_t pos = wcur[d] + ((rk[r / 2] >> (16 * (r & 1))) & 0xffffu);
            stage_k[pos] = k[r];
            stage_v[pos] = v[r];
        }
    }
    __syncthreads();
    // d. consecutive threads write consecutive tile-sorted pairs: contiguous inside every digit run
    const uint32_t count = (uint32_t)(hi - lo);
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < count; idx += RS_THREADS) {
        const uint32_t kk = stage_k[idx];
        const size_t pos = (size_t)(delta[(kk >> g.shift) & mask] + idx);
        keys_out[pos] = kk;
        vals_out[pos] = stage_v[idx];
    }
}

int radix_sort_segmented(swb_ctx* c, uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t seg_len,
                         uint32_t nseg, int key_bits, uint32_t** sorted_keys, uint32_t** sorted_vals) {
    *sorted_keys = keys;
    *sorted_vals = vals;
    if (seg_len == 0 || nseg == 0 || key_bits <= 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    const int passes = (key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;
    const int bits = (key_bits + passes - 1) / passes;
    RsGeom g;
    g.seg_len = seg_len;
    g.tiles_per_seg = (uint32_t)((seg_len + RS_TILE - 1) / RS_TILE);
    g.bits = (uint32_t)bits;
    const size_t nblocks = (size_t)g.tiles_per_seg * nseg;
    SWB_REQUIRE(c, nblocks < ((size_t)1 << 31), "radix sort: too many tiles");
    SWB_REQUIRE(c, seg_len * nseg < ((size_t)1 << 32), "radix sort: positions must fit 32 bits");
    const size_t hist_n = ((size_t)1 << bits) * nblocks;
    uint32_t* hist = (uint32_t*)get_scratch(c, "rs_hist", (hist_n + 16) * sizeof(uint32_t));
    if (!hist) return SWB_ENOMEM;
    const size_t smem = ((size_t)(RS_WARPS + 2) * ((size_t)1 << bits) + 2 * RS_TILE) * sizeof(uint32_t);
    SWB_CUDA(c, cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t *kin = keys, *kout = keys_alt, *vin = vals, *vout = vals_alt;
    for (int p = 0; p < passes; p++) {
        g.shift = (uint32_t)(p * bits);
        k_rs_hist<<<(unsigned)nblocks, RS_THREADS, 0, c->stream>>>(hist, kin, g);
        SWB_LAUNCH_CHECK(c, "k_rs_hist");
        int rc = exclusive_scan_u32(c, hist, hist_n);
        if (rc != SWB_OK) return rc;
        k_rs_scatter<<<(unsigned)nblocks, RS_THREADS, smem, c->stream>>>(kout, vout, kin, vin, hist, g);
        SWB_LAUNCH_CHECK(c, "k_rs_scatter");
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *sorted_keys = kin;
    *sorted_vals = vin;
    return SWB_OK;
}

}  // namespace swb
