// MSM stage 4: bucket accumulation -- the dominant kernel (n * W mixed additions).
//
// Work is split by POSITION in the sorted pair list, not by bucket: thread r adds the points of
// sorted positions [r*len, (r+1)*len) and emits one XYZZ partial sum per run of equal bucket ids.
// Every lane therefore performs the same number of mixed additions whatever the digit
// distribution (a bucket holding half of all points -- top window, or scalars equal to 1 -- is
// simply spread over many threads), and a bucket that straddles ranges gets several partial sums
// which msm_reduce.cu adds up.
#include "msm_common.cuh"

namespace swb {

__device__ __forceinline__ G1Aff ld_aff(const Fq* __restrict__ xy, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(xy + 2 * i);
    G1Aff p;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint4 a = q[k];
        p.x.l[4 * k] = a.x; p.x.l[4 * k + 1] = a.y; p.x.l[4 * k + 2] = a.z; p.x.l[4 * k + 3] = a.w;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint4 a = q[3 + k];
        p.y.l[4 * k] = a.x; p.y.l[4 * k + 1] = a.y; p.y.l[4 * k + 2] = a.z; p.y.l[4 * k + 3] = a.w;
    }
    return p;
}


__global__ void __launch_bounds__(128) k_msm_accumulate(G1Xyzz* __restrict__ partial, uint32_t* __restrict__ pkey,
                                                         const uint32_t* __restrict__ range_off,
                                                         const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                         const Fq* __restrict__ bases, size_t total, size_t n, uint32_t B,
                                                         uint32_t len, uint32_t nranges) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nranges) return;
    const size_t p0 = (size_t)r * len;
    const size_t p1 = p0 + len < total ? p0 + len : total;
    uint32_t out = range_off[r];
    uint32_t cur = 0xffffffffu;
    G1Xyzz acc = G1Xyzz::identity();
    // software pipeline: the next point is in flight while the current one is added
    uint32_t k_next = keys[p0], v_next = vals[p0];
    G1Aff pt_next;
    if (k_next < B) pt_next = ld_aff(bases, v_next & 0x7fffffffu);
    for (size_t p = p0; p < p1; p++) {
        const uint32_t kl = k_next, v = v_next;
        G1Aff pt = pt_next;
        if (p + 1 < p1) {
            k_next = keys[p + 1];
            v_next = vals[p + 1];
            if (k_next < B) pt_next = ld_aff(bases, v_next & 0x7fffffffu);
        }
        if (kl >= B) continue;                              // zero digit (tail of a window's segment)
        const uint32_t k = (uint32_t)(p / n) * B + kl;      // global bucket id
        if (k != cur) {
            if (cur != 0xffffffffu) {
                partial[out] = acc;
                pkey[out] = cur;
                out++;
            }
            cur = k;
            acc = G1Xyzz::identity();
        }
        if (pt.is_identity()) continue;
        if (v >> 31) pt.y = pt.y.neg();
        acc.add_affine(pt.x, pt.y);
    }
    if (cur != 0xffffffffu) {
        partial[out] = acc;
        pkey[out] = cur;
    }
}

int msm_launch_accumulate(swb_ctx* c, const MsmPlan& pl, const MsmBuffers& bf, const uint32_t* sorted_keys,
                          const uint32_t* sorted_vals, const Fq* bases) {
    k_msm_accumulate<<<(pl.nranges + 127) / 128, 128, 0, c->stream>>>(bf.partial, bf.pkey, bf.range_off, sorted_keys, sorted_vals,
                                                                     bases, pl.total, pl.seg_len, pl.B, pl.range_len, pl.nranges);
    SWB_LAUNCH_CHECK(c, "k_msm_accumulate");
    return SWB_OK;
}

}  // namespace swb
