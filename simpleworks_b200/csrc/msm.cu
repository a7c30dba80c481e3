// Variable-base multi-scalar multiplication over BLS12-377 G1 (Pippenger, bucket method).
//
// Replaces ark_ec::msm::VariableBaseMSM::multi_scalar_mul (ark-ec 0.3 msm/variable_base.rs),
// the hot loop of kzg10::commit / open under Marlin::index and Marlin::prove (reference
// src/marlin/mod.rs:75,92; src/merkle_tree/simple_merkle_tree.rs:83,119).  arkworks uses unsigned
// c-bit windows with 2^c - 1 Jacobian buckets per window and one rayon task per window; here:
//
//   1. k_msm_digits     scalars -> signed c-bit digits (halves the bucket count); one
//                        (bucket id, point index | sign) pair per scalar and window
//   2. radix sort        pairs grouped by bucket id over all windows at once
//   3. k_msm_bounds     bucket boundaries by binary search in the sorted keys
//   4. k_msm_accumulate one thread per bucket: XYZZ += +-base (8M + 2S mixed additions)
//   5. k_msm_segments / k_msm_window_reduce
//                        sum_k k * B_k per window: per-segment running sums, then a
//                        block-cooperative suffix scan + tree reduction in shared memory
//   6. host              Horner over the W window sums (c doublings each) and normalisation
//
// The result is the affine value (as a Z = 1 Jacobian), which is unique, hence bit-identical to
// what arkworks' callers see after into_affine().
#include <cub/cub.cuh>

#include "ctx.hpp"
#include "g1.cuh"

namespace swb {

constexpr int MSM_MAX_WINDOWS = 64;
constexpr int MSM_RED_THREADS = 256;   // segments per window in the bucket reduction

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

// 104-byte ABI records -> 96-byte device records; infinity -> (0,0)
__global__ void k_bases_convert(Fq* __restrict__ out, const uint8_t* __restrict__ in, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + i * 104);
    const bool inf = (src[24] & 0xffu) != 0;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + 2 * i);
#pragma unroll
    for (int k = 0; k < 24; k++) dst[k] = inf ? 0u : src[k];
}

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

// ---- 3. start[b] = first sorted position with key >= b, b in [0, nb] ---------------------------
__global__ void k_msm_bounds(uint32_t* __restrict__ start, const uint32_t* __restrict__ keys, size_t total, uint32_t nb) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    size_t lo = 0, hi = total;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (keys[mid] < b) lo = mid + 1; else hi = mid;
    }
    start[b] = (uint32_t)lo;
}

// ---- 4. bucket accumulation ---------------------------------------------------------------
__global__ void __launch_bounds__(128) k_msm_accumulate(G1Xyzz* __restrict__ buckets, const uint32_t* __restrict__ start,
                                                         const uint32_t* __restrict__ vals, const Fq* __restrict__ bases,
                                                         uint32_t nb) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t p0 = start[b], p1 = start[b + 1];
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t v = vals[p];
        G1Aff pt = ld_aff(bases, v & 0x7fffffffu);
        if (pt.is_identity()) continue;
        if (v >> 31) pt.y = pt.y.neg();
        acc.add_affine(pt.x, pt.y);
    }
    buckets[b] = acc;
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

// ---- host-side tail ------------------------------------------------------------------------
static void xyzz_to_out(const G1Xyzz& p, swb_g1_jacobian* out) {
    memset(out, 0, sizeof *out);
    Fq one = Fq::one();
    if (p.is_identity()) {
        memcpy(out->y.l, one.l, 48);    // (0, 1, 0) = GroupProjective::zero()
        return;
    }
    Fq inv = (p.zz * p.zzz).inverse();          // 1 / (zz * zzz)
    Fq x = p.x * (p.zzz * inv);                 // X / ZZ
    Fq y = p.y * (p.zz * inv);                  // Y / ZZZ
    memcpy(out->x.l, x.l, 48);
    memcpy(out->y.l, y.l, 48);
    memcpy(out->z.l, one.l, 48);
}

static int pick_window(swb_ctx* c, size_t n) {
    if (c->msm_window_override) return c->msm_window_override;
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    int cb = lg - 6;
    if (cb < 4) cb = 4;
    if (cb > 16) cb = 16;
    return cb;
}

static int msm_run(swb_ctx* c, const swb_bases* bases, size_t offset, const void* scalars_dev, size_t n, int montgomery,
                   swb_g1_jacobian* out) {
    SWB_REQUIRE(c, bases && out, "msm: NULL argument");
    SWB_REQUIRE(c, bases->ctx == c, "msm: bases belong to another context");
    SWB_REQUIRE(c, offset <= bases->n && n <= bases->n - offset, "msm: offset + n exceeds the loaded bases");
    SWB_REQUIRE(c, n < ((size_t)1 << 31), "msm: n must be < 2^31");
    if (n == 0) {
        xyzz_to_out(G1Xyzz::identity(), out);
        return SWB_OK;
    }
    SWB_REQUIRE(c, scalars_dev != nullptr, "msm: NULL scalars");
    SWB_CUDA(c, cudaSetDevice(c->device));
    const int cb = pick_window(c, n);
    const int nwin = (254 + cb - 1) / cb;
    SWB_REQUIRE(c, nwin <= MSM_MAX_WINDOWS, "msm: too many windows");
    const uint32_t B = 1u << (cb - 1);
    const uint32_t nb = (uint32_t)nwin * B;
    const size_t total = n * (size_t)nwin;
    SWB_REQUIRE(c, total < ((size_t)1 << 32), "msm: n * windows must be < 2^32");

    uint32_t* keys = (uint32_t*)get_scratch(c, "msm_keys", total * 4 * 2);
    uint32_t* vals = (uint32_t*)get_scratch(c, "msm_vals", total * 4 * 2);
    uint32_t* start = (uint32_t*)get_scratch(c, "msm_start", ((size_t)nb + 2) * 4);
    G1Xyzz* buckets = (G1Xyzz*)get_scratch(c, "msm_buckets", (size_t)nb * sizeof(G1Xyzz));
    if (!keys || !vals || !start || !buckets) return SWB_ENOMEM;
    uint32_t* keys2 = keys + total;
    uint32_t* vals2 = vals + total;

    {
        size_t blocks = (n + 255) / 256;
        size_t cap = (size_t)c->sm_count * 8;
        if (blocks > cap) blocks = cap;
        k_msm_digits<<<(unsigned)blocks, 256, 0, c->stream>>>(keys, vals, (const uint32_t*)scalars_dev, n, cb, nwin, montgomery, nb);
        SWB_LAUNCH_CHECK(c, "k_msm_digits");
    }
    {
        int end_bit = 1;
        while ((1ull << end_bit) <= nb) end_bit++;
        size_t tmp_bytes = 0;
        SWB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, total, 0, end_bit, c->stream));
        void* tmp = get_scratch(c, "msm_sort_tmp", tmp_bytes);
        if (!tmp) return SWB_ENOMEM;
        SWB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, total, 0, end_bit, c->stream));
        c->launches += 4;
    }
    k_msm_bounds<<<(nb + 1 + 255) / 256, 256, 0, c->stream>>>(start, keys2, total, nb);
    SWB_LAUNCH_CHECK(c, "k_msm_bounds");
    k_msm_accumulate<<<(nb + 127) / 128, 128, 0, c->stream>>>(buckets, start, vals2, bases->xy + 2 * offset, nb);
    SWB_LAUNCH_CHECK(c, "k_msm_accumulate");

    // bucket reduction
    uint32_t nseg = B < (uint32_t)MSM_RED_THREADS ? B : (uint32_t)MSM_RED_THREADS;
    uint32_t L = B / nseg, log_L = 0;
    while ((1u << log_L) < L) log_L++;
    const uint32_t nseg_total = nseg * (uint32_t)nwin;
    G1Xyzz* seg = (G1Xyzz*)get_scratch(c, "msm_seg", (size_t)nseg_total * 2 * sizeof(G1Xyzz));
    G1Xyzz* wins = (G1Xyzz*)get_scratch(c, "msm_wins", (size_t)MSM_MAX_WINDOWS * sizeof(G1Xyzz));
    if (!seg || !wins) return SWB_ENOMEM;
    k_msm_segments<<<(nseg_total + 127) / 128, 128, 0, c->stream>>>(seg, seg + nseg_total, buckets, B, L, nseg_total);
    SWB_LAUNCH_CHECK(c, "k_msm_segments");
    const size_t red_smem = 2 * MSM_RED_THREADS * sizeof(G1Xyzz);
    SWB_CUDA(c, cudaFuncSetAttribute(k_msm_window_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
    k_msm_window_reduce<<<nwin, MSM_RED_THREADS, red_smem, c->stream>>>(wins, seg, seg + nseg_total, nseg, log_L);
    SWB_LAUNCH_CHECK(c, "k_msm_window_reduce");

    G1Xyzz hw[MSM_MAX_WINDOWS];
    SWB_CUDA(c, cudaMemcpyAsync(hw, wins, sizeof(G1Xyzz) * nwin, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    // Horner over windows, most significant first
    G1Xyzz acc = hw[nwin - 1];
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < cb; k++) acc = acc.dbl();
        acc.add(hw[w]);
    }
    xyzz_to_out(acc, out);
    return SWB_OK;
}

}  // namespace swb

using namespace swb;

extern "C" {

int swb_bases_load_dev(swb_ctx* c, const swb_g1_affine* dev, size_t n, swb_bases** out) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, out != nullptr, "bases_load: out is NULL");
    SWB_REQUIRE(c, n == 0 || dev != nullptr, "bases_load: NULL points");
    SWB_CUDA(c, cudaSetDevice(c->device));
    swb_bases* b = new swb_bases();
    b->ctx = c;
    b->n = n;
    cudaError_t e = cudaMalloc(&b->xy, (n ? n : 1) * 96);
    if (e != cudaSuccess) {
        delete b;
        return set_err(c, SWB_ENOMEM, "bases_load: cudaMalloc(%zu) failed: %s", n * 96, cudaGetErrorString(e));
    }
    if (n) {
        k_bases_convert<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(b->xy, (const uint8_t*)dev, n);
        c->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            cudaFree(b->xy);
            delete b;
            return cuda_fail(c, e, "k_bases_convert");
        }
    }
    *out = b;
    return SWB_OK;
}

int swb_bases_load(swb_ctx* c, const swb_g1_affine* host, size_t n, swb_bases** out) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, out != nullptr, "bases_load: out is NULL");
    SWB_REQUIRE(c, n == 0 || host != nullptr, "bases_load: NULL points");
    SWB_CUDA(c, cudaSetDevice(c->device));
    void* staging = nullptr;
    if (n) {
        cudaError_t e = cudaMalloc(&staging, n * sizeof(swb_g1_affine));
        if (e != cudaSuccess) return set_err(c, SWB_ENOMEM, "bases_load: staging cudaMalloc failed: %s", cudaGetErrorString(e));
        e = cudaMemcpyAsync(staging, host, n * sizeof(swb_g1_affine), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            cudaFree(staging);
            return cuda_fail(c, e, "bases_load H2D");
        }
    }
    int rc = swb_bases_load_dev(c, (const swb_g1_affine*)staging, n, out);
    if (staging) cudaFree(staging);
    return rc;
}

size_t swb_bases_len(const swb_bases* b) { return b ? b->n : 0; }

void swb_bases_free(swb_bases* b) {
    if (!b) return;
    if (b->ctx) {
        cudaSetDevice(b->ctx->device);
        cudaStreamSynchronize(b->ctx->stream);
    }
    if (b->xy) cudaFree(b->xy);
    delete b;
}

int swb_msm_set_window_bits(swb_ctx* c, int cb) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, cb == 0 || (cb >= 2 && cb <= 20), "msm_set_window_bits: c must be 0 or in [2,20]");
    c->msm_window_override = cb;
    return SWB_OK;
}

int swb_msm_g1_dev(swb_ctx* c, const swb_bases* b, size_t offset, const swb_bigint256* scalars_dev, size_t n,
                   swb_g1_jacobian* out) {
    if (!c) return SWB_EARG;
    return msm_run(c, b, offset, scalars_dev, n, 0, out);
}

int swb_msm_g1_fr_dev(swb_ctx* c, const swb_bases* b, size_t offset, const swb_fr* scalars_dev, size_t n,
                      swb_g1_jacobian* out) {
    if (!c) return SWB_EARG;
    return msm_run(c, b, offset, scalars_dev, n, 1, out);
}

int swb_msm_g1(swb_ctx* c, const swb_bases* b, size_t offset, const swb_bigint256* scalars_host, size_t n,
               swb_g1_jacobian* out) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, n == 0 || scalars_host != nullptr, "msm: NULL scalars");
    void* d = nullptr;
    if (n) {
        SWB_CUDA(c, cudaSetDevice(c->device));
        d = get_scratch(c, "msm_scalars", n * 32);
        if (!d) return SWB_ENOMEM;
        SWB_CUDA(c, cudaMemcpyAsync(d, scalars_host, n * 32, cudaMemcpyHostToDevice, c->stream));
    }
    return msm_run(c, b, offset, d, n, 0, out);
}

int swb_g1_sum_jacobian(swb_ctx* c, const swb_g1_jacobian* pts, size_t n, swb_g1_jacobian* out) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, out && (n == 0 || pts), "g1_sum: NULL argument");
    G1Xyzz acc = G1Xyzz::identity();
    for (size_t i = 0; i < n; i++) {
        G1Xyzz p;
        Fq z;
        memcpy(p.x.l, pts[i].x.l, 48);
        memcpy(p.y.l, pts[i].y.l, 48);
        memcpy(z.l, pts[i].z.l, 48);
        if (z.is_zero()) continue;
        p.zz = z.sqr();                 // Jacobian (X, Y, Z) == XYZZ (X, Y, Z^2, Z^3)
        p.zzz = p.zz * z;
        acc.add(p);
    }
    xyzz_to_out(acc, out);
    return SWB_OK;
}

}  // extern "C"
