// Radix-2 evaluation-domain NTT over BLS12-377 Fr.
//
// Replaces ark_poly::Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place (ark-poly
// 0.3 domain/radix2/{mod,fft}.rs), which Marlin's indexer and the three prover rounds call ~20
// times per proof (reference src/marlin/mod.rs:75,92 -> ark-marlin ahp/{indexer,prover}.rs).
// Same contract: natural order in and out, w_n = ROOT^(2^(47-log n)), inverse scales by n^-1,
// coset shift g = 22.  DFT values are unique, so results are bit-identical to arkworks'.
//
// Schedule: n = R_1 * R_2 * ... * R_m with every R_s <= 2^8 (multi-radix Cooley-Tukey).  Pass s<m
// transforms along digit s for a tile of T adjacent columns held in shared memory (coalesced
// T*32-byte runs), multiplies by the inter-digit twiddles and writes back in place; the last pass
// transforms contiguous rows and writes them transposed (digit-reversed) so the output is in
// natural order.  Inside a tile the R-point transform is a decimation-in-frequency network with
// the bit-reversal folded into the store indices.  Twiddles come from three 1024-entry tables of
// powers of the 2^30-th root of unity (w^e = T2[e>>20] * T1[(e>>10)&1023] * T0[e&1023]).
#include "ctx.hpp"

namespace swb {

#ifndef NTT_THREADS_DEF
#define NTT_THREADS_DEF 256
#endif
constexpr int NTT_THREADS = NTT_THREADS_DEF;
#ifndef NTT_TILE_LOG_DEF
#define NTT_TILE_LOG_DEF 11
#endif
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 3
#endif
constexpr int NTT_TILE_LOG = NTT_TILE_LOG_DEF;   // R*T = 2048 elements = 64 KB of shared memory
constexpr uint32_t NTT_MAX_LOG = 30;
#ifndef NTT_RADIX4
#define NTT_RADIX4 1                        // stage pairs in registers (0: one stage per shared-memory round trip)
#endif
#ifndef NTT_BF_ILP
#define NTT_BF_ILP 1                        // butterflies in flight per thread
#endif
constexpr int NTT_MAX_DIGIT = NTT_TILE_LOG;           // one column of the tile
constexpr int NTT_MAX_SMEM = 3 * (1 << NTT_TILE_LOG) * 16;   // final pass with T = 2: [R][T + 1] elements = 96 KB

struct NttPass {
    const Fr* in;
    Fr* out;
    uint32_t log_n, log_r, log_t;
    uint32_t mode;       // 0: column pass (in place layout), 1: final row pass (transposed store)
    uint32_t log_m;      // column mode: log2 of the trailing block size M_s
    uint32_t tw_shift;   // column mode: 30 - (log_r + log_m)
    uint32_t ndig;       // final mode: number of leading digits
    uint32_t dig_log[4]; // final mode: log2 R_1 .. R_{m-1}
    uint32_t inverse, pre_coset, post_coset, post_scale;
    Fr scale;            // n^-1 (Montgomery) when post_scale
    size_t batch_stride;
    const Fr* tw_full;   // column mode: powers of the 2^full_log-th root of unity, or NULL (running products)
    uint32_t full_log;
};

__device__ __forceinline__ Fr tw_lookup(const Fr* __restrict__ tab, uint32_t e) {
    const uint32_t d0 = e & 1023u, d1 = (e >> 10) & 1023u, d2 = e >> 20;
    Fr r = tab[2048 + d2];
    if (d1) r = r * tab[1024 + d1];
    if (d0) r = r * tab[d0];
    return r;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t v, uint32_t bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

// shared tile: 16-byte halves of each element kept in two planes so that consecutive elements are
// consecutive 16-byte words (conflict-free 128-bit accesses).
struct Tile {
    uint4* lo;
    uint4* hi;
    __device__ __forceinline__ Fr get(uint32_t e) const {
        Fr r;
        uint4 a = lo[e], b = hi[e];
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void put(uint32_t e, const Fr& v) const {
        lo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) k_ntt_pass(NttPass p, const Fr* __restrict__ tw_root,
                                                           const Fr* __restrict__ tw_coset) {
    extern __shared__ uint4 smem[];
    const uint32_t R = 1u << p.log_r, T = 1u << p.log_t, tile = R * T;
    // the tile is [j][t] in both modes (t fastest: butterflies and stores touch consecutive words).
    // Final mode loads rows of consecutive j, i.e. it writes the tile with stride Tp between lanes:
    // Tp = T + 1 is odd in 16-byte words, which spreads those stores over the banks.
    const uint32_t Tp = (p.mode == 0 || T == 1u) ? T : T + 1u;
    const uint32_t planes = R * Tp;
    Tile tl{smem, smem + planes};
    const Fr* in = p.in + (size_t)blockIdx.y * p.batch_stride;
    Fr* out = p.out + (size_t)blockIdx.y * p.batch_stride;
    const uint32_t tid = threadIdx.x;

    // ---- where this block's tile lives ------------------------------------------------------
    size_t in_base, out_base, in_j_stride, in_t_stride, out_k_stride, out_t_stride;
    uint32_t col0 = 0;   // first column index (column mode), for the inter-digit twiddle
    if (p.mode == 0) {
        const uint32_t chunks = 1u << (p.log_m - p.log_t);
        const uint32_t outer = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
        col0 = chunk << p.log_t;
        in_base = ((size_t)outer << (p.log_r + p.log_m)) + col0;
        out_base = in_base;
        in_j_stride = out_k_stride = (size_t)1 << p.log_m;
        in_t_stride = out_t_stride = 1;
    } else {
        // rows rho' = g*T + t in output (natural) row order; their position in the in-place
        // layout is the digit reversal of rho' over (R_1 .. R_{m-1})
        const uint32_t rp0 = blockIdx.x << p.log_t;
        uint32_t rest = rp0, rho = 0;
        for (uint32_t d = 0; d < p.ndig; d++) {
            const uint32_t dig = rest & ((1u << p.dig_log[d]) - 1u);
            rest >>= p.dig_log[d];
            rho = (rho << p.dig_log[d]) | dig;   // k_1 ends up most significant
        }
        in_base = (size_t)rho << p.log_r;
        in_j_stride = 1;
        in_t_stride = p.ndig ? ((size_t)1 << (p.log_n - p.dig_log[0])) : 0;
        out_base = rp0;
        out_t_stride = 1;
        out_k_stride = (size_t)1 << (p.log_n - p.log_r);
    }

    // ---- load (with optional coset pre-scaling by g^i) ---------------------------------------
    for (uint32_t idx = tid; idx < tile; idx += NTT_THREADS) {
        uint32_t j, t, e;
        if (p.mode == 0) { t = idx & (T - 1); j = idx >> p.log_t; e = idx; }
        else { j = idx & (R - 1); t = idx >> p.log_r; e = j * Tp + t; }   // global runs along j, tile along t
        const size_t gi = in_base + (size_t)j * in_j_stride + (size_t)t * in_t_stride;
        Fr v = ld_fr(in + gi);
        if (p.pre_coset) v = v * tw_lookup(tw_coset, (uint32_t)gi);
        tl.put(e, v);
    }
    __syncthreads();

    // ---- R-point DIF network along j ----------------------------------------------------------
    // element (j,t) sits at j*Tp + t
    const uint32_t js = Tp, ts = 1u;
    const uint32_t nbf = tile >> 1;
    uint32_t s = 0;
#if NTT_RADIX4
    // two stages per trip through shared memory: a thread holds the four elements {j0, j0 + q, j0 + 2q, j0 + 3q}
    // (q = quarter of the current block) in registers, runs the stage-s butterflies (j0, j0 + 2q), (j0 + q, j0 + 3q)
    // and then the stage-(s+1) butterflies (j0, j0 + q), (j0 + 2q, j0 + 3q): same products, half the shared-memory
    // traffic and half the barriers.  An odd number of stages leaves the last one to the radix-2 loop below.
    for (; s + 2 <= p.log_r; s += 2) {
        const uint32_t log_q = p.log_r - 2 - s, q = 1u << log_q;
        const uint32_t sh = NTT_MAX_LOG - p.log_r, emask = (1u << NTT_MAX_LOG) - 1u;
        for (uint32_t id = tid; id < (tile >> 2); id += NTT_THREADS) {
            const uint32_t t = id & (T - 1), b = id >> p.log_t;
            const uint32_t pos = b & (q - 1), grp = b >> log_q;
            const uint32_t j0 = (grp << (log_q + 2)) + pos;
            const uint32_t i0 = j0 * js + t * ts, i1 = i0 + q * js, i2 = i1 + q * js, i3 = i2 + q * js;
            uint32_t ex0 = ((pos << s) << sh), ex1 = (((pos + q) << s) << sh), ex2 = ((pos << (s + 1)) << sh);
            if (p.inverse) {
                ex0 = ((1u << NTT_MAX_LOG) - ex0) & emask;
                ex1 = ((1u << NTT_MAX_LOG) - ex1) & emask;
                ex2 = ((1u << NTT_MAX_LOG) - ex2) & emask;
            }
            Fr a0 = tl.get(i0), a2 = tl.get(i2);
            Fr b0 = Fr::add_lazy(a0, a2);
            Fr b2 = Fr::sub_lazy(a0, a2);
            if (ex0) b2 = Fr::mul_lazy(b2, tw_lookup(tw_root, ex0));
            else Fr::final_sub2(b2);
            Fr a1 = tl.get(i1), a3 = tl.get(i3);
            Fr b1 = Fr::add_lazy(a1, a3);
            Fr b3 = Fr::mul_lazy(Fr::sub_lazy(a1, a3), tw_lookup(tw_root, ex1));     // w^(R/4) * ...: never trivial
            Fr c1 = Fr::sub_lazy(b0, b1), c3 = Fr::sub_lazy(b2, b3);
            if (ex2) {
                const Fr w2 = tw_lookup(tw_root, ex2);
                c1 = Fr::mul_lazy(c1, w2);
                c3 = Fr::mul_lazy(c3, w2);
            } else {
                Fr::final_sub2(c1);
                Fr::final_sub2(c3);
            }
            tl.put(i0, Fr::add_lazy(b0, b1));
            tl.put(i1, c1);
            tl.put(i2, Fr::add_lazy(b2, b3));
            tl.put(i3, c3);
        }
        __syncthreads();
    }
#endif
    for (; s < p.log_r; s++) {
        const uint32_t log_half = p.log_r - 1 - s, half = 1u << log_half;
        // two butterflies per iteration, loaded before either is computed: their Montgomery products are
        // independent dependency chains that the scheduler interleaves (one chain alone leaves the multiplier idle
        // between dependent instructions)
        for (uint32_t idx = tid; idx < nbf; idx += NTT_BF_ILP * NTT_THREADS) {
            uint32_t e0[NTT_BF_ILP], e1[NTT_BF_ILP], ex[NTT_BF_ILP];
            Fr u[NTT_BF_ILP], v[NTT_BF_ILP];
            bool on[NTT_BF_ILP];
#pragma unroll
            for (int q = 0; q < NTT_BF_ILP; q++) {
                const uint32_t id = idx + q * NTT_THREADS;
                on[q] = id < nbf;
                const uint32_t t = id & (T - 1), b = id >> p.log_t;
                const uint32_t pos = b & (half - 1), grp = b >> log_half;
                const uint32_t j0 = (grp << (log_half + 1)) + pos;
                e0[q] = j0 * js + t * ts;
                e1[q] = e0[q] + half * js;
                ex[q] = (pos << s) << (NTT_MAX_LOG - p.log_r);
                if (p.inverse) ex[q] = ((1u << NTT_MAX_LOG) - ex[q]) & ((1u << NTT_MAX_LOG) - 1u);
                if (on[q]) { u[q] = tl.get(e0[q]); v[q] = tl.get(e1[q]); }
            }
            Fr tw[NTT_BF_ILP];
#pragma unroll
            for (int q = 0; q < NTT_BF_ILP; q++)
                if (on[q] && ex[q]) tw[q] = tw_lookup(tw_root, ex[q]);
#pragma unroll
            for (int q = 0; q < NTT_BF_ILP; q++) {
                if (!on[q]) continue;
                // lazy butterflies (Harvey): tile values live in [0, 2r); the difference needs no comparison and the
                // product no final subtraction -- its result is below 2r for an operand below 4r
                const Fr sum = Fr::add_lazy(u[q], v[q]);
                Fr dif = Fr::sub_lazy(u[q], v[q]);
                if (ex[q]) dif = Fr::mul_lazy(dif, tw[q]);
                else Fr::final_sub2(dif);
                tl.put(e0[q], sum);
                tl.put(e1[q], dif);
            }
        }
        __syncthreads();
    }

    // ---- store: X[k] is at bit-reversed position; apply inter-digit twiddle / scaling --------
    if (p.mode == 0) {
        // thread keeps its column t fixed and walks k
        const uint32_t t = tid & (T - 1);
        const uint32_t col = col0 + t;
        const uint32_t k0 = tid >> p.log_t, dk = NTT_THREADS >> p.log_t;
        if (p.tw_full) {
            // inter-digit twiddle w^(k*col), w of order R*M, read from the table of all powers: one product per
            // element instead of two (apply + advance a running power), and no serial chain between iterations.
            // The table loads are scattered 32-byte reads; the pass is issue-bound with DRAM nearly idle.
            const uint32_t up = p.full_log - (p.log_r + p.log_m), fmask = (1u << p.full_log) - 1u;
            constexpr uint32_t U = 4;
            for (uint32_t k = k0; k < R; k += U * dk) {
                Fr w[U];
#pragma unroll
                for (uint32_t q = 0; q < U; q++) {
                    const uint32_t kk = k + q * dk;
                    uint32_t e = (kk * col) << up;                       // kk * col < R * M <= 2^full_log
                    if (p.inverse) e = ((1u << p.full_log) - e) & fmask;
                    if (kk < R && e) w[q] = ld_fr(p.tw_full + e);
                    else w[q] = Fr::one();
                }
#pragma unroll
                for (uint32_t q = 0; q < U; q++) {
                    const uint32_t kk = k + q * dk;
                    if (kk >= R) break;
                    Fr v = tl.get(bitrev(kk, p.log_r) * T + t);
                    if (col) v = Fr::mul_lazy(v, w[q]);                  // intermediate passes hand on values in [0, 2r)
                    st_fr(out + out_base + (size_t)kk * out_k_stride + t, v);
                }
            }
        } else {
            // twiddle w^(k*col) advances by w^(dk*col)
            const uint32_t mask = (1u << NTT_MAX_LOG) - 1u;
            uint32_t e_w = (uint32_t)((((uint64_t)k0 * col) << p.tw_shift) & mask);
            uint32_t e_s = (uint32_t)((((uint64_t)dk * col) << p.tw_shift) & mask);
            if (p.inverse) { e_w = ((1u << NTT_MAX_LOG) - e_w) & mask; e_s = ((1u << NTT_MAX_LOG) - e_s) & mask; }
            Fr w = tw_lookup(tw_root, e_w);
            const Fr step = tw_lookup(tw_root, e_s);
            for (uint32_t k = k0; k < R; k += dk) {
                Fr v = tl.get(bitrev(k, p.log_r) * T + t);
                if (col) v = Fr::mul_lazy(v, w);
                st_fr(out + out_base + (size_t)k * out_k_stride + t, v);
                w = w * step;
            }
        }
    } else {
        for (uint32_t idx = tid; idx < tile; idx += NTT_THREADS) {
            const uint32_t t = idx & (T - 1), k = idx >> p.log_t;
            Fr v = tl.get(bitrev(k, p.log_r) * Tp + t);
            const size_t go = out_base + (size_t)k * out_k_stride + (size_t)t * out_t_stride;
            // v is in [0, 2r): a full product brings it back to the canonical range, otherwise one subtraction does
            if (p.post_coset) v = v * tw_lookup(tw_coset, (uint32_t)go);
            if (p.post_scale) v = v * p.scale;
            if (!p.post_coset && !p.post_scale) v = Fr::reduce_lazy(v);
            st_fr(out + go, v);
        }
    }
}

// table[l][j] = base_l ^ j
__global__ void k_build_pow_table(Fr* __restrict__ tab, Fr b0, Fr b1, Fr b2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * 1024) return;
    const uint32_t l = i >> 10, j = i & 1023u;
    const Fr base = l == 0 ? b0 : (l == 1 ? b1 : b2);
    tab[i] = base.pow_u64(j);
}

// tab[i] = w^i for the 2^log_l-th root of unity w, from the three-level tables
__global__ void k_build_full_table(Fr* __restrict__ tab, const Fr* __restrict__ tw_root, uint32_t log_l) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_l)) return;
    st_fr(tab + i, tw_lookup(tw_root, (uint32_t)(i << (NTT_MAX_LOG - log_l))));
}

// Direct twiddle table for transforms of up to 2^log_n points (32 B << log_n of device memory: 2 GiB at 2^26), grown on
// demand; SWB_NTT_TABLE_MAX_LOG caps it (0 switches it off).  Returns the table's log size or 0 (running products).
static uint32_t ntt_full_table(swb_ctx* c, uint32_t log_n) {
    // read at every call (a getenv is nothing beside a transform), so that tests can compare both twiddle paths
    const char* env = getenv("SWB_NTT_TABLE_MAX_LOG");
    const long capv = env ? atol(env) : 26;
    const uint32_t cap = (uint32_t)(capv < 0 ? 0 : (capv > 27 ? 27 : capv));
    if (log_n > cap || c->tw_full_failed) return 0;
    if (c->tw_full && c->tw_full_log >= log_n) return c->tw_full_log;
    if (c->tw_full) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->tw_full);
        c->tw_full = nullptr;
        c->tw_full_log = 0;
    }
    if (cudaMalloc(&c->tw_full, sizeof(Fr) << log_n) != cudaSuccess) {
        cudaGetLastError();
        c->tw_full = nullptr;
        c->tw_full_failed = true;
        return 0;
    }
    k_build_full_table<<<(unsigned)((((size_t)1 << log_n) + 255) / 256), 256, 0, c->stream>>>(c->tw_full, c->tw_root, log_n);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) {
        cudaFree(c->tw_full);
        c->tw_full = nullptr;
        c->tw_full_failed = true;
        return 0;
    }
    c->tw_full_log = log_n;
    return log_n;
}

// widest digit of the automatic plan (SWB_NTT_MAX_DIGIT overrides: tuning aid)
static uint32_t ntt_max_digit() {
    static const uint32_t v = [] {
        const char* e = getenv("SWB_NTT_MAX_DIGIT");
        const long d = e ? atol(e) : 10;
        return (uint32_t)(d >= 1 && d <= NTT_MAX_DIGIT ? d : 10);
    }();
    return v;
}

static Fr host_fr_const(const uint32_t (&v)[8]) {
    Fr r;
    for (int i = 0; i < 8; i++) r.l[i] = v[i];
    return r;
}

int ntt_build_tables(swb_ctx* c) {
    SWB_CUDA(c, cudaMalloc(&c->tw_root, sizeof(Fr) * 3 * 1024));
    SWB_CUDA(c, cudaMalloc(&c->tw_gen, sizeof(Fr) * 3 * 1024));
    SWB_CUDA(c, cudaMalloc(&c->tw_geninv, sizeof(Fr) * 3 * 1024));
    const uint32_t root_init[8] = SWB_FR_ROOT_OF_UNITY_INIT, gen_init[8] = SWB_FR_GENERATOR_INIT,
                   geninv_init[8] = SWB_FR_GENERATOR_INV_INIT;
    Fr w = host_fr_const(root_init);
    for (uint32_t i = NTT_MAX_LOG; i < SWB_FR_TWO_ADICITY; i++) w = w.sqr();   // 2^30-th root of unity
    Fr bases[3][3];
    bases[0][0] = w;
    bases[1][0] = host_fr_const(gen_init);
    bases[2][0] = host_fr_const(geninv_init);
    for (int k = 0; k < 3; k++)
        for (int l = 1; l < 3; l++) {
            Fr b = bases[k][l - 1];
            for (int i = 0; i < 10; i++) b = b.sqr();
            bases[k][l] = b;
        }
    Fr* tabs[3] = {c->tw_root, c->tw_gen, c->tw_geninv};
    for (int k = 0; k < 3; k++) {
        k_build_pow_table<<<12, 256, 0, c->stream>>>(tabs[k], bases[k][0], bases[k][1], bases[k][2]);
        SWB_LAUNCH_CHECK(c, "k_build_pow_table");
    }
    SWB_CUDA(c, cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_MAX_SMEM));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SWB_OK;
}

// plan: n = R_1 * ... * R_m.  Every pass costs a trip through HBM and (all but the last) one inter-digit
// twiddle multiplication per element on top of its butterflies, so fewer, wider passes win as long as the
// tile keeps a few adjacent columns (T = 2^(11 - digit)) for coalescing: measured on B200 (profiles/
// r2_ntt_plan_sweep.json).  SWB_NTT_PLAN_<log_n>=a,b,c overrides (tuning aid).
static int plan_digits(uint32_t log_n, uint32_t* dig) {
    if (log_n == 0) { dig[0] = 0; return 1; }
    char name[32];
    snprintf(name, sizeof name, "SWB_NTT_PLAN_%u", log_n);
    if (const char* e = getenv(name)) {
        int m = 0;
        uint32_t sum = 0;
        while (*e && m < 8) {
            char* end = nullptr;
            const unsigned long v = strtoul(e, &end, 10);
            if (end == e || v == 0 || v > (unsigned long)NTT_MAX_DIGIT) { m = 0; break; }
            dig[m++] = (uint32_t)v;
            sum += (uint32_t)v;
            e = *end == ',' ? end + 1 : end;
        }
        if (m > 0 && sum == log_n) return m;
    }
    const uint32_t maxd = ntt_max_digit();
    int m = (int)((log_n + maxd - 1) / maxd);
    for (int s = 0; s < m; s++) dig[s] = log_n / m + ((uint32_t)s < log_n % m ? 1 : 0);
    // the butterfly network runs its stages in pairs (NTT_RADIX4), an odd digit leaves a single stage with a barrier of
    // its own: trade one stage between two odd digits (2^26: 9,9,8 -> 10,8,8 16.84 -> 16.46 ms; 2^22: 8,7,7 -> 8,8,6)
    for (;;) {
        int a = -1, b = -1;
        for (int s = 0; s < m; s++)
            if (dig[s] & 1u) { if (a < 0) a = s; else if (b < 0) b = s; }
        if (b < 0 || dig[a] + 1 > maxd || dig[b] < 2) break;
        dig[a]++;
        dig[b]--;
    }
    return m;
}

static int ntt_run(swb_ctx* c, Fr* data, uint32_t log_n, size_t batch, int inverse, int coset) {
    SWB_REQUIRE(c, log_n <= NTT_MAX_LOG, "ntt: log_n > 30 not supported");
    SWB_REQUIRE(c, data != nullptr, "ntt: NULL data");
    if (batch == 0) return SWB_OK;
    SWB_REQUIRE(c, batch < 65536, "ntt: batch too large");
    SWB_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)1 << log_n;
    uint32_t dig[8];
    const int m = plan_digits(log_n, dig);
    SWB_REQUIRE(c, m <= 5, "ntt: internal plan error");
    Fr* tmp = nullptr;
    if (m > 1) {
        tmp = (Fr*)get_scratch(c, "ntt_tmp", sizeof(Fr) * n * batch);
        if (!tmp) return SWB_ENOMEM;
    }
    Fr scale = Fr::one();
    if (inverse) {
        // n^-1 = (2^-1)^log_n
        const uint32_t two_inv[8] = SWB_FR_TWO_INV_INIT;
        Fr ti = host_fr_const(two_inv);
        for (uint32_t i = 0; i < log_n; i++) scale = scale * ti;
    }
    const uint32_t full_log = m > 1 ? ntt_full_table(c, log_n) : 0;
    uint32_t log_m = log_n;   // trailing block size before pass s
    static const char* const pass_names[8] = {"pass0", "pass1", "pass2", "pass3", "pass4", "pass5", "pass6", "pass7"};
    StageTimer tm(c, "ntt");
    for (int s = 0; s < m; s++) {
        NttPass p{};
        p.log_n = log_n;
        p.log_r = dig[s];
        p.inverse = inverse ? 1 : 0;
        p.batch_stride = n;
        p.scale = scale;
        log_m -= dig[s];
        const bool last = (s == m - 1);
        p.in = (s == 0) ? data : tmp;
        p.out = (m == 1) ? data : (last ? data : tmp);
        p.pre_coset = (s == 0 && coset && !inverse) ? 1 : 0;
        p.post_coset = (last && coset && inverse) ? 1 : 0;
        p.post_scale = (last && inverse) ? 1 : 0;
        uint32_t log_t;
        size_t blocks;
        if (!last) {
            p.mode = 0;
            p.log_m = log_m;
            p.tw_shift = NTT_MAX_LOG - (dig[s] + log_m);
            p.tw_full = full_log ? c->tw_full : nullptr;
            p.full_log = full_log;
            log_t = NTT_TILE_LOG - dig[s];
            if (log_t > log_m) log_t = log_m;
            blocks = n >> (dig[s] + log_t);
        } else {
            p.mode = 1;
            p.ndig = (uint32_t)(m - 1);
            for (int d = 0; d < m - 1; d++) p.dig_log[d] = dig[d];
            log_t = NTT_TILE_LOG - dig[s];
            const uint32_t lim = m > 1 ? dig[0] : 0;
            if (log_t > lim) log_t = lim;
            blocks = n >> (dig[s] + log_t);
        }
        p.log_t = log_t;
        const size_t smem = (last && log_t > 0) ? ((size_t)32 << dig[s]) * (((size_t)1 << log_t) + 1)          // [R][T + 1] elements
                                                : ((size_t)32) << (dig[s] + log_t);
        SWB_REQUIRE(c, smem <= (size_t)NTT_MAX_SMEM, "ntt: plan needs more shared memory than the kernel may use");
        const Fr* coset_tab = inverse ? c->tw_geninv : c->tw_gen;
        dim3 grid((unsigned)blocks, (unsigned)batch);
        k_ntt_pass<<<grid, NTT_THREADS, smem, c->stream>>>(p, c->tw_root, coset_tab);
        SWB_LAUNCH_CHECK(c, "k_ntt_pass");
        tm.mark(pass_names[s]);
    }
    return SWB_OK;
}

}  // namespace swb

using namespace swb;

extern "C" {

int swb_ntt_fr_dev(swb_ctx* c, swb_fr* inout, uint32_t log_n, int inverse, int coset) {
    if (!c) return SWB_EARG;
    return ntt_run(c, (Fr*)inout, log_n, 1, inverse, coset);
}

int swb_ntt_fr_batch_dev(swb_ctx* c, swb_fr* inout, uint32_t log_n, size_t batch, int inverse, int coset) {
    if (!c) return SWB_EARG;
    return ntt_run(c, (Fr*)inout, log_n, batch, inverse, coset);
}

int swb_ntt_fr(swb_ctx* c, swb_fr* inout_host, uint32_t log_n, int inverse, int coset) {
    if (!c) return SWB_EARG;
    SWB_REQUIRE(c, log_n <= NTT_MAX_LOG, "ntt: log_n > 30 not supported");
    SWB_REQUIRE(c, inout_host != nullptr, "ntt: NULL data");
    SWB_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = sizeof(Fr) << log_n;
    Fr* d = (Fr*)get_scratch(c, "ntt_io", bytes);
    if (!d) return SWB_ENOMEM;
    SWB_CUDA(c, cudaMemcpyAsync(d, inout_host, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = ntt_run(c, d, log_n, 1, inverse, coset);
    if (rc != SWB_OK) return rc;
    SWB_CUDA(c, cudaMemcpyAsync(inout_host, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SWB_OK;
}

}  // extern "C"
