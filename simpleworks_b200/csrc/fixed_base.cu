// Fixed-base powers: out[i] = beta^i * g, i < n, as affine points.
//
// Replaces what KZG10::setup does for powers_of_g / powers_of_gamma_g (ark-poly-commit 0.3
// kzg10/mod.rs): FixedBaseMSM::get_window_table + FixedBaseMSM::multi_scalar_mul over the
// powers of beta, then ProjectiveCurve::batch_normalization_into_affine.  Reached from
// Marlin::universal_setup <- reference src/marlin/mod.rs:52, simple_merkle_tree.rs:39 (1 572 862
// points per table for the (100000, 25000, 300000) bound every non-toy example uses).
//
// Device schedule: a 32 x 256 table of d * 2^(8o) * g (8-bit windows, affine), then one thread
// per power: beta^i by square-and-multiply, 32 mixed additions, one Fermat inversion to affine.
// Affine results are unique, so they equal arkworks' bit for bit.
#define SWB_FP_NOINLINE_MUL
#include "ctx.hpp"
#include "g1.cuh"

namespace swb {

constexpr int FB_WINDOW = 8;
constexpr int FB_OUTER = 32;

__device__ __forceinline__ G1Aff xyzz_to_affine(const G1Xyzz& p) {
    G1Aff a;
    if (p.is_identity()) {
        a.x = Fq::zero();
        a.y = Fq::zero();
        return a;
    }
    Fq inv = (p.zz * p.zzz).inverse();
    a.x = p.x * (p.zzz * inv);
    a.y = p.y * (p.zz * inv);
    return a;
}

// table[o*256 + d] = d * 2^(8o) * g from pow2[k] = 2^k * g
__global__ void __launch_bounds__(128) k_fb_table(G1Aff* __restrict__ table, const G1Aff* __restrict__ pow2) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= FB_OUTER * 256) return;
    const uint32_t o = e >> 8, d = e & 255u;
    G1Xyzz acc = G1Xyzz::identity();
    for (int b = 0; b < FB_WINDOW; b++) {
        if ((d >> b) & 1u) {
            G1Aff p = pow2[o * FB_WINDOW + b];
            if (!p.is_identity()) acc.add_affine(p.x, p.y);
        }
    }
    table[e] = xyzz_to_affine(acc);
}

// RECORD = 104: ABI GroupAffine records (x | y | infinity flag); RECORD = 96: resident base layout
template <int RECORD>
__global__ void __launch_bounds__(128) k_fb_powers(uint8_t* __restrict__ out, const G1Aff* __restrict__ table, Fr beta, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr s = beta.pow_u64((uint64_t)i).to_canonical();
    G1Xyzz acc = G1Xyzz::identity();
    for (int o = 0; o < FB_OUTER; o++) {
        const uint32_t d = (s.l[o >> 2] >> ((o & 3) * 8)) & 255u;
        if (d) {
            G1Aff p = table[o * 256 + d];
            if (!p.is_identity()) acc.add_affine(p.x, p.y);
        }
    }
    const G1Aff a = xyzz_to_affine(acc);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + i * RECORD);
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k] = a.x.l[k];
#pragma unroll
    for (int k = 0; k < 12; k++) dst[12 + k] = a.y.l[k];
    if (RECORD == 104) {
        dst[24] = a.is_identity() ? 1u : 0u;
        dst[25] = 0u;
    }
}

// resident 96-byte records -> 104-byte ABI records
__global__ void k_bases_export(uint8_t* __restrict__ out, const Fq* __restrict__ xy, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(xy + 2 * i);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + i * 104);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 24; k++) { dst[k] = src[k]; any |= src[k]; }
    dst[24] = any ? 0u : 1u;
    dst[25] = 0u;
}

}  // namespace swb

using namespace swb;

// builds the window table for g on the device; returns it in *table
static int fb_prepare(swb_ctx* c, const swb_g1_jacobian* g_host, const swb_fr* beta_host, G1Aff** table, Fr* beta) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    // host: 2^k * g for k < 256, affine
    G1Xyzz p;
    Fq z;
    memcpy(p.x.l, g_host->x.l, 48);
    memcpy(p.y.l, g_host->y.l, 48);
    memcpy(z.l, g_host->z.l, 48);
    if (z.is_zero()) p = G1Xyzz::identity();
    else {
        p.zz = z.sqr();
        p.zzz = p.zz * z;
    }
    std::vector<G1Aff> pow2(FB_OUTER * FB_WINDOW);
    for (size_t k = 0; k < pow2.size(); k++) {
        if (p.is_identity()) {
            pow2[k].x = Fq::zero();
            pow2[k].y = Fq::zero();
        } else {
            Fq inv = (p.zz * p.zzz).inverse();
            pow2[k].x = p.x * (p.zzz * inv);
            pow2[k].y = p.y * (p.zz * inv);
        }
        p = p.dbl();
    }
    memcpy(beta->l, beta_host->l, 32);
    G1Aff* d_pow2 = (G1Aff*)get_scratch(c, "fb_pow2", sizeof(G1Aff) * pow2.size());
    G1Aff* d_table = (G1Aff*)get_scratch(c, "fb_table", sizeof(G1Aff) * FB_OUTER * 256);
    if (!d_pow2 || !d_table) return SWB_ENOMEM;
    SWB_CUDA(c, cudaMemcpyAsync(d_pow2, pow2.data(), sizeof(G1Aff) * pow2.size(), cudaMemcpyHostToDevice, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));   // pow2 is a local
    k_fb_table<<<FB_OUTER * 256 / 128, 128, 0, c->stream>>>(d_table, d_pow2);
    SWB_LAUNCH_CHECK(c, "k_fb_table");
    *table = d_table;
    return SWB_OK;
}

extern "C" int swb_fixed_base_powers(swb_ctx* c, const swb_g1_jacobian* g_host, const swb_fr* beta_host, size_t n,
                                     swb_g1_affine* out_host) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, g_host && beta_host && (n == 0 || out_host), "fixed_base_powers: NULL argument");
    if (n == 0) return SWB_OK;
    G1Aff* d_table = nullptr;
    Fr beta;
    int rc = fb_prepare(c, g_host, beta_host, &d_table, &beta);
    if (rc != SWB_OK) return rc;
    uint8_t* d_out = (uint8_t*)get_scratch(c, "fb_out", n * 104);
    if (!d_out) return SWB_ENOMEM;
    k_fb_powers<104><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_out, d_table, beta, n);
    SWB_LAUNCH_CHECK(c, "k_fb_powers");
    SWB_CUDA(c, cudaMemcpyAsync(out_host, d_out, n * 104, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SWB_OK;
}

extern "C" int swb_bases_from_powers(swb_ctx* c, const swb_g1_jacobian* g_host, const swb_fr* beta_host, size_t n,
                                     swb_bases** out) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, g_host && beta_host && out, "bases_from_powers: NULL argument");
    G1Aff* d_table = nullptr;
    Fr beta;
    int rc = fb_prepare(c, g_host, beta_host, &d_table, &beta);
    if (rc != SWB_OK) return rc;
    swb_bases* b = new swb_bases();
    b->ctx = c;
    b->n = n;
    cudaError_t e = cudaMalloc(&b->xy, (n ? n : 1) * 96);
    if (e != cudaSuccess) {
        delete b;
        return set_err(c, SWB_ENOMEM, "bases_from_powers: cudaMalloc(%zu) failed: %s", n * 96, cudaGetErrorString(e));
    }
    if (n) {
        k_fb_powers<96><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((uint8_t*)b->xy, d_table, beta, n);
        c->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            cudaFree(b->xy);
            delete b;
            return cuda_fail(c, e, "k_fb_powers");
        }
    }
    *out = b;
    return SWB_OK;
}

extern "C" int swb_bases_export(swb_ctx* c, const swb_bases* b, size_t offset, size_t n, swb_g1_affine* out_host) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, b && (n == 0 || out_host), "bases_export: NULL argument");
    SWB_REQUIRE(c, offset <= b->n && n <= b->n - offset, "bases_export: range exceeds the loaded bases");
    if (n == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    uint8_t* d_out = (uint8_t*)get_scratch(c, "fb_out", n * 104);
    if (!d_out) return SWB_ENOMEM;
    k_bases_export<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_out, b->xy + 2 * offset, n);
    SWB_LAUNCH_CHECK(c, "k_bases_export");
    SWB_CUDA(c, cudaMemcpyAsync(out_host, d_out, n * 104, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SWB_OK;
}
