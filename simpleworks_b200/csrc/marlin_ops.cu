// Prover-side device kernels beside the polynomial arithmetic of polyops.cu, so that the large
// vectors of a Marlin proof are born in HBM instead of being built on the host and copied:
//   k_rand_*       the 3|H| coefficients of the outer-sumcheck mask polynomial (DensePolynomial::rand in
//                  ark-marlin's prover_first_round) -- the zk RNG is a ChaCha stream, so candidate k is
//                  a pure function of (key, position) and the accept/reject walk is a prefix sum
//   k_csr_spmv     z_A = A z, z_B = B z, z_C = C z and t = sum_M eta_M M^T r(alpha, .) (ark-marlin
//                  ahp/prover.rs: prover_init, prover_second_round) on the index's CSR / CSC arrays
//   k_witness_evals  the layout of the witness on H \ X (prover_first_round)
// All memory-bound, one thread per output element.
#include "ctx.hpp"
#include "marlin_ops.hpp"
#include "radix_sort.hpp"

namespace swb {

struct ChaChaKey { uint32_t k[8]; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

__device__ void chacha_block(const ChaChaKey& key, uint64_t counter, int rounds, uint32_t out[16]) {
    uint32_t st[16];
    st[0] = 0x61707865u; st[1] = 0x3320646eu; st[2] = 0x79622d32u; st[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) st[4 + i] = key.k[i];
    st[12] = (uint32_t)counter;
    st[13] = (uint32_t)(counter >> 32);
    st[14] = 0;
    st[15] = 0;
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = st[i];
#define SWB_QR(a, b, c, d)                               \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);        \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);        \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);         \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < rounds; r += 2) {
        SWB_QR(0, 4, 8, 12) SWB_QR(1, 5, 9, 13) SWB_QR(2, 6, 10, 14) SWB_QR(3, 7, 11, 15)
        SWB_QR(0, 5, 10, 15) SWB_QR(1, 6, 11, 12) SWB_QR(2, 7, 8, 13) SWB_QR(3, 4, 9, 14)
    }
#undef SWB_QR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

// candidate k: 8 stream words from `pos + 8k`; returns whether it is a valid field element
__device__ bool rand_candidate(const ChaChaKey& key, int rounds, uint64_t pos, size_t k, Fr* out) {
    const uint64_t w0 = pos + 8 * (uint64_t)k;
    const uint32_t off = (uint32_t)(w0 & 15);
    uint32_t buf[32], w[8];
    chacha_block(key, w0 >> 4, rounds, buf);
    if (off > 8) chacha_block(key, (w0 >> 4) + 1, rounds, buf + 16);   // straddles into the next block
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = buf[off + i];
    w[7] &= 0xFFFFFFFFu >> 3;
    bool lt = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        const uint32_t m = FrParams::mod(i);
        if (w[i] != m) { lt = w[i] < m; break; }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out->l[i] = w[i];
    return lt;
}

__global__ void __launch_bounds__(256) k_rand_flags(uint32_t* __restrict__ flags, ChaChaKey key, int rounds, uint64_t pos, size_t cand) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > cand) return;
    Fr v;
    flags[k] = k < cand && rand_candidate(key, rounds, pos, k, &v) ? 1u : 0u;   // flags[cand] = 0: the scan leaves the total there
}
// idx = exclusive scan of the flags; the first `want` accepted candidates go to out[idx]
__global__ void __launch_bounds__(256) k_rand_emit(Fr* __restrict__ out, uint32_t* __restrict__ used, const uint32_t* __restrict__ idx,
                                                    ChaChaKey key, int rounds, uint64_t pos, size_t cand, uint32_t want) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cand) return;
    const uint32_t at = idx[k];
    if (idx[k + 1] == at || at >= want) return;
    Fr v;
    rand_candidate(key, rounds, pos, k, &v);
    uint4* q = reinterpret_cast<uint4*>(out + at);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    if (at == want - 1) *used = (uint32_t)(k + 1);
}

int rand_fr_dev(swb_ctx* c, Fr* out, size_t n, const uint32_t key[8], int rounds, uint64_t pos, uint64_t* words_used) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    ChaChaKey kk;
    for (int i = 0; i < 8; i++) kk.k[i] = key[i];
    const uint64_t pos0 = pos;
    size_t done = 0;
    while (done < n) {
        const size_t want = n - done;
        SWB_REQUIRE(c, want < ((size_t)1 << 31), "rand_fr: too many elements");
        size_t cand = want + want - want / 4 + 1024;          // acceptance is ~0.583: 1.75x covers it with margin
        if (cand > ((size_t)1 << 31)) cand = (size_t)1 << 31;
        uint32_t* idx = (uint32_t*)get_scratch(c, "rand_idx", (cand + 2) * sizeof(uint32_t));
        if (!idx) return SWB_ENOMEM;
        uint32_t* used = idx + cand + 1;
        k_rand_flags<<<(unsigned)((cand + 1 + 255) / 256), 256, 0, c->stream>>>(idx, kk, rounds, pos, cand);
        SWB_LAUNCH_CHECK(c, "k_rand_flags");
        int rc = exclusive_scan_u32(c, idx, cand + 1);
        if (rc != SWB_OK) return rc;
        SWB_CUDA(c, cudaMemsetAsync(used, 0, sizeof(uint32_t), c->stream));
        k_rand_emit<<<(unsigned)((cand + 255) / 256), 256, 0, c->stream>>>(out + done, used, idx, kk, rounds, pos, cand, (uint32_t)want);
        SWB_LAUNCH_CHECK(c, "k_rand_emit");
        uint32_t h[2];                                        // [accepted among the candidates, candidates consumed]
        SWB_CUDA(c, cudaMemcpyAsync(h, idx + cand, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        SWB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h[0] >= want) {
            pos += 8 * (uint64_t)h[1];
            done = n;
        } else {
            pos += 8 * (uint64_t)cand;
            done += h[0];
        }
    }
    *words_used = pos - pos0;
    return SWB_OK;
}

// ---- sparse matrix-vector product --------------------------------------------------------------
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
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

struct SpmvWeights { Fr w[3]; };

__global__ void __launch_bounds__(256) k_csr_spmv(Fr* __restrict__ out, size_t nout, size_t nrows, const uint32_t* __restrict__ start,
                                                   const uint32_t* __restrict__ col, const Fr* __restrict__ coef,
                                                   const uint8_t* __restrict__ tag, const Fr* __restrict__ x, SpmvWeights wt) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nout) return;
    Fr acc = Fr::zero();
    if (r < nrows) {
        for (uint32_t k = start[r]; k < start[r + 1]; k++) {
            Fr term = ld_fr(coef + k) * ld_fr(x + col[k]);
            if (tag) {
                const uint8_t t = tag[k];
                term = term * (t == 0 ? wt.w[0] : t == 1 ? wt.w[1] : wt.w[2]);
            }
            acc = acc + term;
        }
    }
    st_fr(out + r, acc);
}

int csr_spmv_dev(swb_ctx* c, Fr* out, size_t nout, size_t nrows, const uint32_t* start, const uint32_t* col, const Fr* coef,
                 const uint8_t* tag, const Fr* x, const Fr weights[3]) {
    if (nout == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    SpmvWeights wt;
    for (int i = 0; i < 3; i++) wt.w[i] = weights ? weights[i] : Fr::one();
    k_csr_spmv<<<(unsigned)((nout + 255) / 256), 256, 0, c->stream>>>(out, nout, nrows, start, col, coef, tag, x, wt);
    SWB_LAUNCH_CHECK(c, "k_csr_spmv");
    return SWB_OK;
}

// ---- witness layout on H ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_witness_evals(Fr* __restrict__ out, size_t nh, size_t ratio, const Fr* __restrict__ z,
                                                        size_t ninst, size_t nvars, const Fr* __restrict__ xh) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nh) return;
    Fr v = Fr::zero();
    if (k % ratio != 0) {
        const size_t wi = ninst + (k - k / ratio - 1);
        if (wi < nvars) v = ld_fr(z + wi);
        v = v - ld_fr(xh + k);
    }
    st_fr(out + k, v);
}

int witness_evals_dev(swb_ctx* c, Fr* out, size_t nh, size_t ratio, const Fr* z, size_t ninst, size_t nvars, const Fr* xh) {
    if (nh == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    k_witness_evals<<<(unsigned)((nh + 255) / 256), 256, 0, c->stream>>>(out, nh, ratio, z, ninst, nvars, xh);
    SWB_LAUNCH_CHECK(c, "k_witness_evals");
    return SWB_OK;
}

}  // namespace swb
