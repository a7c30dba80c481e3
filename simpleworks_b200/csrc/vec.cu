// Element-wise field kernels (parity probes for the Montgomery arithmetic, batch inversion) and
// the limb-product peak measurement that gives the integer roofline its denominator.
#include "ctx.hpp"

namespace swb {

template <class F, int OP>
__global__ void __launch_bounds__(256) k_vec_op(F* __restrict__ r, const F* __restrict__ a, const F* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        F x = a[i], y = b[i];
        F z = OP == 0 ? x * y : (OP == 1 ? x + y : x - y);
        r[i] = z;
    }
}

template <class F, int OP>
static int launch_vec(swb_ctx* c, void* r, const void* a, const void* b, size_t n) {
    if (!c) return SWB_EARG;
    if (n == 0) return SWB_OK;
    SWB_REQUIRE(c, r && a && b, "vec op: NULL pointer");
    SWB_CUDA(c, cudaSetDevice(c->device));
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)c->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_vec_op<F, OP><<<(unsigned)blocks, 256, 0, c->stream>>>((F*)r, (const F*)a, (const F*)b, n);
    SWB_LAUNCH_CHECK(c, "k_vec_op");
    return SWB_OK;
}

// ark_ff::batch_inversion: each thread inverts a chunk of CH consecutive elements with
// Montgomery's trick (one Fermat inversion per chunk); zero entries are left untouched.
template <int CH>
__global__ void __launch_bounds__(128) k_fr_batch_inverse(Fr* __restrict__ v, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * CH;
    if (lo >= n) return;
    size_t hi = lo + CH < n ? lo + CH : n;
    Fr prefix[CH];
    Fr acc = Fr::one();
    for (size_t i = lo; i < hi; i++) {
        prefix[i - lo] = acc;
        Fr x = v[i];
        if (!x.is_zero()) acc = acc * x;
    }
    acc = acc.inverse();
    for (size_t i = hi; i-- > lo;) {
        Fr x = v[i];
        if (x.is_zero()) continue;
        v[i] = acc * prefix[i - lo];
        acc = acc * x;
    }
}

// Register-resident multiplication loop: ILP independent chains per thread, nothing but
// Montgomery products.  Used to measure the achievable limb-product rate.
template <class F, int ILP>
__global__ void __launch_bounds__(256) k_mul_peak(F* __restrict__ out, const F* __restrict__ seed, int iters) {
    F x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        x[k] = seed[(threadIdx.x + k) & 31];
        y[k] = seed[(threadIdx.x + 7 * k + 3) & 31];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = x[k] * y[k];
#pragma unroll
        for (int k = 0; k < ILP; k++) y[k] = y[k] * x[k];
    }
    F acc = x[0];
#pragma unroll
    for (int k = 1; k < ILP; k++) acc = acc + x[k];
#pragma unroll
    for (int k = 0; k < ILP; k++) acc = acc + y[k];
    // keep the result observable without meaningful memory traffic
    if (acc.l[0] == 0x12345678u && acc.l[1] == 0x1abcdef0u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// raw pipe probes: 16 independent accumulators per thread, no memory traffic
template <int KIND>
__global__ void __launch_bounds__(256) k_imad_peak(uint64_t* __restrict__ out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t lo[16], hi[16];
    uint32_t a = a0 + threadIdx.x, b = b0 ^ (blockIdx.x * 2654435761u);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        lo[k] = k * 0x9e3779b9u + threadIdx.x;
        hi[k] = k * 0x7f4a7c15u + blockIdx.x;
    }
    for (int it = 0; it < iters; it++) {
        if (KIND == 2) {
            // carry chains of four IMAD.WIDE.U32.X each, the shape the Montgomery rows have
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                    "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                    "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                    "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(lo[k]), "+r"(hi[k]), "+r"(lo[k + 1]), "+r"(hi[k + 1]), "+r"(lo[k + 2]), "+r"(hi[k + 2]),
                      "+r"(lo[k + 3]), "+r"(hi[k + 3])
                    : "r"(a), "r"(b));
            }
        } else if (KIND == 1) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                // multiplicand = the accumulator's own low word, so the product cannot be hoisted
                asm volatile("{.reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t;}"
                             : "+r"(lo[k]), "+r"(hi[k]) : "r"(b));
        } else if (KIND == 3) {
            // add-with-carry chains of four (IADD3.X), the ALU-pipe side of the arithmetic
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
                asm volatile(
                    "add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, %9;"
                    "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %8; addc.u32 %7, %7, %9;"
                    : "+r"(lo[k]), "+r"(hi[k]), "+r"(lo[k + 1]), "+r"(hi[k + 1]), "+r"(lo[k + 2]), "+r"(hi[k + 2]),
                      "+r"(lo[k + 3]), "+r"(hi[k + 3])
                    : "r"(a), "r"(b));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[k]) : "r"(b), "r"(a));
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= lo[k] ^ hi[k];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int KIND>
static int measure_imad(swb_ctx* c, int iters, double* ops) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    const int threads = 256, blocks = c->sm_count * 8;
    uint64_t* buf = (uint64_t*)get_scratch(c, "peak", sizeof(uint64_t) * (size_t)threads * blocks + 4096);
    if (!buf) return SWB_ENOMEM;
    cudaEvent_t e0, e1;
    SWB_CUDA(c, cudaEventCreate(&e0));
    SWB_CUDA(c, cudaEventCreate(&e1));
    k_imad_peak<KIND><<<blocks, threads, 0, c->stream>>>(buf, 12345u, 678910u, iters / 4 + 1);
    c->launches++;
    SWB_CUDA(c, cudaEventRecord(e0, c->stream));
    k_imad_peak<KIND><<<blocks, threads, 0, c->stream>>>(buf, 12345u, 678910u, iters);
    SWB_LAUNCH_CHECK(c, "k_imad_peak");
    SWB_CUDA(c, cudaEventRecord(e1, c->stream));
    SWB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    SWB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops = (double)threads * blocks * (double)iters * (KIND == 3 ? 32.0 : 16.0) / (ms * 1e-3);
    return SWB_OK;
}

template <class F, int ILP, int LIMBS32>
static int measure_peak(swb_ctx* c, int iters, double* lps, double* mps) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    const int threads = 256;
    const int blocks = c->sm_count * 4;
    F* buf = (F*)get_scratch(c, "peak", sizeof(F) * ((size_t)threads * blocks + 32));
    if (!buf) return SWB_ENOMEM;
    // seeds: small non-trivial field elements
    F h[32];
    for (int i = 0; i < 32; i++) {
        h[i] = F::one();
        for (int j = 0; j <= i; j++) h[i] = h[i] + h[i] + F::one();
        h[i] = h[i] * h[i] * h[i];
    }
    F* seed = buf + (size_t)threads * blocks;
    SWB_CUDA(c, cudaMemcpyAsync(seed, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    SWB_CUDA(c, cudaEventCreate(&e0));
    SWB_CUDA(c, cudaEventCreate(&e1));
    k_mul_peak<F, ILP><<<blocks, threads, 0, c->stream>>>(buf, seed, iters / 4 + 1);   // warm-up
    c->launches++;
    SWB_CUDA(c, cudaEventRecord(e0, c->stream));
    k_mul_peak<F, ILP><<<blocks, threads, 0, c->stream>>>(buf, seed, iters);
    SWB_LAUNCH_CHECK(c, "k_mul_peak");
    SWB_CUDA(c, cudaEventRecord(e1, c->stream));
    SWB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    SWB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double muls = (double)threads * blocks * (double)iters * ILP * 2.0;
    double per_mul = 2.0 * LIMBS32 * LIMBS32;    // counted in 32-bit-limb products whatever the representation
    if (mps) *mps = muls / (ms * 1e-3);
    if (lps) *lps = muls * per_mul / (ms * 1e-3);
    return SWB_OK;
}

}  // namespace swb

using namespace swb;

extern "C" {

int swb_fr_mul_vec_dev(swb_ctx* c, swb_fr* r, const swb_fr* a, const swb_fr* b, size_t n) { return launch_vec<Fr, 0>(c, r, a, b, n); }
int swb_fr_add_vec_dev(swb_ctx* c, swb_fr* r, const swb_fr* a, const swb_fr* b, size_t n) { return launch_vec<Fr, 1>(c, r, a, b, n); }
int swb_fr_sub_vec_dev(swb_ctx* c, swb_fr* r, const swb_fr* a, const swb_fr* b, size_t n) { return launch_vec<Fr, 2>(c, r, a, b, n); }
int swb_fq_mul_vec_dev(swb_ctx* c, swb_fq* r, const swb_fq* a, const swb_fq* b, size_t n) { return launch_vec<Fq, 0>(c, r, a, b, n); }
int swb_fq_add_vec_dev(swb_ctx* c, swb_fq* r, const swb_fq* a, const swb_fq* b, size_t n) { return launch_vec<Fq, 1>(c, r, a, b, n); }
int swb_fq_sub_vec_dev(swb_ctx* c, swb_fq* r, const swb_fq* a, const swb_fq* b, size_t n) { return launch_vec<Fq, 2>(c, r, a, b, n); }

int swb_fr_batch_inverse_dev(swb_ctx* c, swb_fr* v, size_t n) {
    if (!c) return SWB_EARG;
    if (n == 0) return SWB_OK;
    SWB_REQUIRE(c, v, "batch_inverse: NULL pointer");
    SWB_CUDA(c, cudaSetDevice(c->device));
    constexpr int CH = 16;
    size_t threads = (n + CH - 1) / CH;
    k_fr_batch_inverse<CH><<<(unsigned)((threads + 127) / 128), 128, 0, c->stream>>>((Fr*)v, n);
    SWB_LAUNCH_CHECK(c, "k_fr_batch_inverse");
    return SWB_OK;
}

int swb_measure_imad_peak(swb_ctx* c, int kind, int iters, double* ops) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, iters > 0 && ops && kind >= 0 && kind <= 3, "measure_imad_peak: bad arguments");
    switch (kind) {
        case 0: return measure_imad<0>(c, iters, ops);
        case 1: return measure_imad<1>(c, iters, ops);
        case 2: return measure_imad<2>(c, iters, ops);
        default: return measure_imad<3>(c, iters, ops);
    }
}

int swb_measure_mul_peak(swb_ctx* c, int field, int iters, double* lps, double* mps) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, iters > 0 && (field == 0 || field == 1), "measure_mul_peak: bad arguments");
    return field == 0 ? measure_peak<Fr, 4, 8>(c, iters, lps, mps) : measure_peak<Fq, 2, 12>(c, iters, lps, mps);
}

}  // extern "C"
