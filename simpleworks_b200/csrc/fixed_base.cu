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
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + i * 104);
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k] = a.x.l[k];
#pragma unroll
    for (int k = 0; k < 12; k++) dst[12 + k] = a.y.l[k];
    dst[24] = a.is_identity() ? 1u : 0u;
    dst[25] = 0u;
}

}  // namespace swb

using namespace swb;

extern "C" int swb_fixed_base_powers(swb_ctx* c, const swb_g1_jacobian* g_host, const swb_fr* beta_host, size_t n,
                                     swb_g1_affine* out_host) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, g_host && beta_host && (n == 0 || out_host), "fixed_base_powers: NULL argument");
    if (n == 0) return SWB_OK;
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
    Fr beta;
    memcpy(beta.l, beta_host->l, 32);
    G1Aff* d_pow2 = (G1Aff*)get_scratch(c, "fb_pow2", sizeof(G1Aff) * pow2.size());
    G1Aff* d_table = (G1Aff*)get_scratch(c, "fb_table", sizeof(G1Aff) * FB_OUTER * 256);
    uint8_t* d_out = (uint8_t*)get_scratch(c, "fb_out", n * 104);
    if (!d_pow2 || !d_table || !d_out) return SWB_ENOMEM;
    SWB_CUDA(c, cudaMemcpyAsync(d_pow2, pow2.data(), sizeof(G1Aff) * pow2.size(), cudaMemcpyHostToDevice, c->stream));
    k_fb_table<<<FB_OUTER * 256 / 128, 128, 0, c->stream>>>(d_table, d_pow2);
    SWB_LAUNCH_CHECK(c, "k_fb_table");
    k_fb_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_out, d_table, beta, n);
    SWB_LAUNCH_CHECK(c, "k_fb_powers");
    SWB_CUDA(c, cudaMemcpyAsync(out_host, d_out, n * 104, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SWB_OK;
}
