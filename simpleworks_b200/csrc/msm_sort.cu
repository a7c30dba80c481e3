// MSM stages 1-3: signed-digit decomposition, radix sort of (bucket, point) pairs, and the
// per-range run counts that tell the accumulation kernel where to write.
#include <cub/cub.cuh>

#include "msm_common.cuh"

namespace swb {

// ---- 1. signed-digit decomposition --------------------------------------------------------
// keys[w*n + i] = w*B + |d| - 1  (or `invalid` when d == 0), vals = i | sign << 31
__global__ void __launch_bounds__(256) k_msm_digits(uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                     const uint32_t* __restrict__ scalars, size_t n, int c, int nwin,
                                                     int montgomery, uint32_t invalid) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t B = 1u << (c - 1);
    for (; i < n; i += stride) {
        Fr s;
        const uint4* q = reinterpret_cast<const uint4*>(scalars + 8 * i);
        uint4 a = q[0], b = q[1];
        s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
        s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
        if (montgomery) s = s.to_canonical();
        uint32_t carry = 0;
        for (int w = 0; w < nwin; w++) {
            const int bit = w * c, limb = bit >> 5, off = bit & 31;
            uint32_t v = 0;
            if (limb < 8) {
                v = s.l[limb] >> off;
                if (off + c > 32 && limb + 1 < 8) v |= s.l[limb + 1] << (32 - off);
            }
            v = (v & ((1u << c) - 1u)) + carry;
            uint32_t neg = 0;
            if (v > B) { v = (1u << c) - v; neg = 1; carry = 1; } else carry = 0;
            keys[(size_t)w * n + i] = v ? (uint32_t)w * B + v - 1 : invalid;
            vals[(size_t)w * n + i] = (uint32_t)i | (neg << 31);
        }
    }
}

// ---- 3. number of runs of equal (valid) keys inside each range of `len` sorted positions -----
// Thread r owns sorted positions [r*len, (r+1)*len).  A run is a maximal stretch of one bucket id
// inside the range; k_msm_accumulate emits exactly one partial sum per run, so the exclusive scan
// of these counts is where each range writes.
__global__ void __launch_bounds__(256) k_msm_range_count(uint32_t* __restrict__ cnt, const uint32_t* __restrict__ keys,
                                                          size_t total, uint32_t nb, uint32_t len, uint32_t nranges) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nranges) return;
    if (r == nranges) { cnt[r] = 0; return; }
    const size_t p0 = (size_t)r * len;
    const size_t p1 = p0 + len < total ? p0 + len : total;
    uint32_t runs = 0, prev = 0xffffffffu;
    for (size_t p = p0; p < p1; p++) {
        const uint32_t k = keys[p];
        if (k >= nb) break;                 // digit-0 entries are sorted to the end
        runs += (k != prev);
        prev = k;
    }
    cnt[r] = runs;
}

int msm_launch_digits_sort(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf, const void* scalars_dev, int montgomery,
                           const uint32_t** sorted_keys, const uint32_t** sorted_vals) {
    const size_t total = pl.total;
    uint32_t* keys2 = bf.keys + total;
    uint32_t* vals2 = bf.vals + total;
    {
        size_t blocks = (pl.n + 255) / 256;
        size_t cap = (size_t)c->sm_count * 8;
        if (blocks > cap) blocks = cap;
        k_msm_digits<<<(unsigned)blocks, 256, 0, c->stream>>>(bf.keys, bf.vals, (const uint32_t*)scalars_dev, pl.n, pl.cb, pl.nwin,
                                                              montgomery, pl.nb);
        SWB_LAUNCH_CHECK(c, "k_msm_digits");
    }
    {
        int end_bit = 1;
        while ((1ull << end_bit) <= pl.nb) end_bit++;
        size_t tmp_bytes = 0;
        SWB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, bf.keys, keys2, bf.vals, vals2, total, 0, end_bit, c->stream));
        void* tmp = get_scratch(c, "msm_sort_tmp", tmp_bytes);
        if (!tmp) return SWB_ENOMEM;
        SWB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, bf.keys, keys2, bf.vals, vals2, total, 0, end_bit, c->stream));
        c->launches += 4;
    }
    k_msm_range_count<<<(pl.nranges + 1 + 255) / 256, 256, 0, c->stream>>>(bf.range_cnt, keys2, total, pl.nb, pl.range_len, pl.nranges);
    SWB_LAUNCH_CHECK(c, "k_msm_range_count");
    {
        size_t tmp_bytes = 0;
        SWB_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bf.range_cnt, bf.range_off, pl.nranges + 1, c->stream));
        void* tmp = get_scratch(c, "msm_scan_tmp", tmp_bytes);
        if (!tmp) return SWB_ENOMEM;
        SWB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, bf.range_cnt, bf.range_off, pl.nranges + 1, c->stream));
        c->launches += 2;
    }
    *sorted_keys = keys2;
    *sorted_vals = vals2;
    return SWB_OK;
}

}  // namespace swb
