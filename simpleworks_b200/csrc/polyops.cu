// Device-resident polynomial / vector kernels over Fr: everything the Marlin rounds do to a
// polynomial between an NTT and an MSM, so that round data never bounces through PCIe (SURVEY
// 8f-1).  Restates on the device: DensePolynomial +=, scalar mul, evaluate (Horner),
// divide_by_vanishing_poly, the KZG witness quotient (p - p(z)) / (X - z) (ark-poly 0.3
// polynomial/univariate/dense.rs, ark-poly-commit kzg10::compute_witness_polynomial), and the
// element-wise products of evaluation vectors.  All bandwidth-bound (32 B in / 32 B out per element,
// one Fr product): coalesced 2 x 128-bit accesses, grid-stride loops sized to the SM count.
#include "ctx.hpp"
#include "polyops.hpp"

namespace swb {

__device__ __forceinline__ Fr ldf(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stf(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// op: 0 a*=b, 1 a+=b, 2 a-=b, 3 a+=c*b, 4 a*=c, 5 a=c0+c1*a
template <int OP>
__global__ void __launch_bounds__(256) k_poly_ew(Fr* __restrict__ a, const Fr* __restrict__ b, size_t n, Fr c0, Fr c1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        Fr x = ldf(a + i), y;
        if (OP <= 3) y = ldf(b + i);
        Fr r;
        if (OP == 0) r = x * y;
        else if (OP == 1) r = x + y;
        else if (OP == 2) r = x - y;
        else if (OP == 3) r = x + c0 * y;
        else if (OP == 4) r = x * c0;
        else r = c0 + c1 * x;
        stf(a + i, r);
    }
}

// out[c] = sum_{i in chunk c} p[i] x^(i - lo)   (chunks of L consecutive coefficients)
__global__ void __launch_bounds__(128) k_poly_chunk_horner(Fr* __restrict__ out, const Fr* __restrict__ p, size_t n, uint32_t L, Fr x) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = c * L;
    if (lo >= n) return;
    const size_t hi = lo + L < n ? lo + L : n;
    Fr acc = Fr::zero();
    for (size_t i = hi; i-- > lo;) acc = acc * x + ldf(p + i);
    stf(out + c, acc);
}

// q[i] = sum_{k>=1} p[i + k n] (i < len - n),  r[i] = sum_{k>=0} p[i + k n] (i < n)
__global__ void __launch_bounds__(256) k_poly_div_vanishing(Fr* __restrict__ q, Fr* __restrict__ r, const Fr* __restrict__ p, size_t len,
                                                             size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t qlen = len > n ? len - n : 0;
    if (i < qlen) {
        Fr acc = Fr::zero();
        for (size_t j = i + n; j < len; j += n) acc = acc + ldf(p + j);
        stf(q + i, acc);
    }
    if (i < n) {
        Fr acc = Fr::zero();
        for (size_t j = i; j < len; j += n) acc = acc + ldf(p + j);
        stf(r + i, acc);
    }
}

// ---- division by X^n - 1 when the quotient has many "rows" (n much smaller than the length) -----
// Index i = j*n + r.  For every residue r the quotient is a suffix sum over j:
//   H_j = p[j n + r] + H_{j+1},   q[(j-1) n + r] = H_j (j >= 1),   rem[r] = H_0.
// Three passes over blocks of C rows: block totals, their suffix (one thread per residue), apply.
__global__ void __launch_bounds__(128) k_divv_totals(Fr* __restrict__ tot, const Fr* __restrict__ p, size_t len, size_t n, size_t rows,
                                                      uint32_t C, size_t nblk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblk * n) return;
    const size_t b = t / n, r = t % n;
    const size_t j1 = (b + 1) * C < rows ? (b + 1) * C : rows;
    Fr acc = Fr::zero();
    for (size_t j = b * C; j < j1; j++) {
        const size_t idx = j * n + r;
        if (idx < len) acc = acc + ldf(p + idx);
    }
    stf(tot + t, acc);
}
__global__ void __launch_bounds__(128) k_divv_carry(Fr* __restrict__ carry, const Fr* __restrict__ tot, size_t n, size_t nblk) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    Fr run = Fr::zero();
    for (size_t b = nblk; b-- > 0;) {
        stf(carry + b * n + r, run);
        run = run + ldf(tot + b * n + r);
    }
}
__global__ void __launch_bounds__(128) k_divv_apply(Fr* __restrict__ q, Fr* __restrict__ rem, const Fr* __restrict__ p,
                                                     const Fr* __restrict__ carry, size_t len, size_t n, size_t rows, uint32_t C,
                                                     size_t nblk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblk * n) return;
    const size_t b = t / n, r = t % n;
    const size_t j1 = (b + 1) * C < rows ? (b + 1) * C : rows;
    Fr h = ldf(carry + t);
    for (size_t j = j1; j-- > b * C;) {
        const size_t idx = j * n + r;
        if (idx >= len) continue;            // ragged last row: nothing there, and no quotient slot either
        h = h + ldf(p + idx);
        if (j >= 1) stf(q + (j - 1) * n + r, h);
        else stf(rem + r, h);
    }
}

// Given the suffix-Horner value at each chunk's upper boundary (carry[c] = H_{hi_c}), walk the chunk
// backwards: H_j = p_j + z H_{j+1};  quotient q_{j-1} = H_j  for j >= 1.
__global__ void __launch_bounds__(128) k_poly_div_linear_apply(Fr* __restrict__ q, const Fr* __restrict__ p, const Fr* __restrict__ carry,
                                                                size_t n, uint32_t L, Fr z) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = c * L;
    if (lo >= n) return;
    const size_t hi = lo + L < n ? lo + L : n;
    Fr h = carry ? ldf(carry + c) : Fr::zero();
    for (size_t j = hi; j-- > lo;) {
        h = ldf(p + j) + z * h;
        if (j >= 1) stf(q + (j - 1), h);
    }
}
// suffix-Horner of a short array in place, one thread: s[c] <- H at the UPPER boundary of chunk c,
// i.e. carry[c] = sum_{c' > c} s[c'] y^(c' - c - 1)
__global__ void k_poly_suffix_small(Fr* __restrict__ carry, const Fr* __restrict__ s, size_t m, Fr y) {
    if (blockIdx.x || threadIdx.x) return;
    Fr h = Fr::zero();
    for (size_t c = m; c-- > 0;) {
        stf(carry + c, h);
        h = ldf(s + c) + y * h;
    }
}

// highest index with a non-zero element, plus one (0 for the zero vector)
__global__ void __launch_bounds__(256) k_poly_len(unsigned long long* __restrict__ out, const Fr* __restrict__ p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long best = 0;
    for (; i < n; i += stride) {
        Fr x = ldf(p + i);
        if (!x.is_zero()) best = i + 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

// out[i] = g^i
__global__ void __launch_bounds__(128) k_poly_powers(Fr* __restrict__ out, size_t n, uint32_t L, Fr g) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = c * L;
    if (lo >= n) return;
    const size_t hi = lo + L < n ? lo + L : n;
    Fr v = g.pow_u64(lo);
    for (size_t i = lo; i < hi; i++) {
        stf(out + i, v);
        v = v * g;
    }
}

static unsigned grid_for(swb_ctx* c, size_t n, int threads) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)c->sm_count * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

template <int OP>
static int launch_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n, const Fr& c0, const Fr& c1) {
    if (n == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    k_poly_ew<OP><<<grid_for(c, n, 256), 256, 0, c->stream>>>(a, b, n, c0, c1);
    SWB_LAUNCH_CHECK(c, "k_poly_ew");
    return SWB_OK;
}

int poly_mul_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n) { return launch_ew<0>(c, a, b, n, Fr::zero(), Fr::zero()); }
int poly_add_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n) { return launch_ew<1>(c, a, b, n, Fr::zero(), Fr::zero()); }
int poly_sub_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n) { return launch_ew<2>(c, a, b, n, Fr::zero(), Fr::zero()); }
int poly_add_scaled_ew(swb_ctx* c, Fr* a, const Fr& s, const Fr* b, size_t n) { return launch_ew<3>(c, a, b, n, s, Fr::zero()); }
int poly_scale_ew(swb_ctx* c, Fr* a, const Fr& s, size_t n) { return launch_ew<4>(c, a, nullptr, n, s, Fr::zero()); }
int poly_lin_ew(swb_ctx* c, Fr* a, const Fr& c0, const Fr& c1, size_t n) { return launch_ew<5>(c, a, nullptr, n, c0, c1); }

int poly_eval_dev(swb_ctx* c, const Fr* p, size_t n, const Fr& x, Fr* out) {
    if (n == 0) { *out = Fr::zero(); return SWB_OK; }
    SWB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t L = 64;
    Fr* tmp = (Fr*)get_scratch(c, "poly_eval", sizeof(Fr) * (2 * ((n + L - 1) / L) + 64));
    if (!tmp) return SWB_ENOMEM;
    const Fr* cur = p;
    size_t m = n;
    Fr y = x;
    Fr* bufs[2] = {tmp, tmp + (n + L - 1) / L + 32};
    int which = 0;
    while (m > L) {
        const size_t chunks = (m + L - 1) / L;
        k_poly_chunk_horner<<<(unsigned)((chunks + 127) / 128), 128, 0, c->stream>>>(bufs[which], cur, m, L, y);
        SWB_LAUNCH_CHECK(c, "k_poly_chunk_horner");
        cur = bufs[which];
        which ^= 1;
        m = chunks;
        y = y.pow_u64(L);
    }
    Fr host[64];
    SWB_CUDA(c, cudaMemcpyAsync(host, cur, sizeof(Fr) * m, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    Fr acc = Fr::zero();
    for (size_t i = m; i-- > 0;) acc = acc * y + host[i];
    *out = acc;
    return SWB_OK;
}

int poly_div_vanishing_dev(swb_ctx* c, Fr* q, Fr* r, const Fr* p, size_t len, size_t n) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    const size_t rows = (len + n - 1) / n;
    if (rows > 8) {
        // long division (e.g. by v_X with |X| = 2): blocked suffix sums, O(len) work
        const uint32_t C = 128;
        const size_t nblk = (rows + C - 1) / C;
        Fr* buf = (Fr*)get_scratch(c, "poly_divv", sizeof(Fr) * (2 * nblk * n + 16));
        if (!buf) return SWB_ENOMEM;
        Fr* tot = buf;
        Fr* carry = buf + nblk * n;
        const unsigned g1 = (unsigned)((nblk * n + 127) / 128);
        k_divv_totals<<<g1, 128, 0, c->stream>>>(tot, p, len, n, rows, C, nblk);
        SWB_LAUNCH_CHECK(c, "k_divv_totals");
        k_divv_carry<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(carry, tot, n, nblk);
        SWB_LAUNCH_CHECK(c, "k_divv_carry");
        k_divv_apply<<<g1, 128, 0, c->stream>>>(q, r, p, carry, len, n, rows, C, nblk);
        SWB_LAUNCH_CHECK(c, "k_divv_apply");
        return SWB_OK;
    }
    const size_t work = len > 2 * n ? len - n : n;
    k_poly_div_vanishing<<<(unsigned)((work + 255) / 256), 256, 0, c->stream>>>(q, r, p, len, n);
    SWB_LAUNCH_CHECK(c, "k_poly_div_vanishing");
    return SWB_OK;
}

// q (n-1 elements) = (p - p(z)) / (X - z)
int poly_div_linear_dev(swb_ctx* c, Fr* q, const Fr* p, size_t n, const Fr& z) {
    if (n <= 1) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t L = 256;
    // level sizes: n -> ceil(n/L) -> ... until <= L
    std::vector<size_t> sizes{n};
    while (sizes.back() > L) sizes.push_back((sizes.back() + L - 1) / L);
    size_t total = 0;
    for (size_t i = 1; i < sizes.size(); i++) total += 2 * sizes[i] + 8;
    Fr* buf = (Fr*)get_scratch(c, "poly_divlin", sizeof(Fr) * (total + 16));
    if (!buf) return SWB_ENOMEM;
    std::vector<Fr*> sums(sizes.size(), nullptr), carries(sizes.size(), nullptr);
    std::vector<Fr> ys(sizes.size());
    Fr* cursor = buf;
    ys[0] = z;
    for (size_t i = 1; i < sizes.size(); i++) {
        sums[i] = cursor; cursor += sizes[i] + 4;
        carries[i] = cursor; cursor += sizes[i] + 4;
        ys[i] = ys[i - 1].pow_u64(L);
    }
    // upward: chunk totals of each level (level i array = totals of level i-1 chunks)
    const Fr* cur = p;
    for (size_t i = 1; i < sizes.size(); i++) {
        k_poly_chunk_horner<<<(unsigned)((sizes[i] + 127) / 128), 128, 0, c->stream>>>(sums[i], cur, sizes[i - 1], L, ys[i - 1]);
        SWB_LAUNCH_CHECK(c, "k_poly_chunk_horner");
        cur = sums[i];
    }
    // top level: carries of the last (short) array, one thread
    const size_t top = sizes.size() - 1;
    if (top >= 1) {
        k_poly_suffix_small<<<1, 1, 0, c->stream>>>(carries[top], sums[top], sizes[top], ys[top]);
        SWB_LAUNCH_CHECK(c, "k_poly_suffix_small");
        // downward: level i carries -> carries of level i-1 chunks
        for (size_t i = top; i-- > 1;) {
            // carries[i][c] = H at the upper boundary of chunk c of level i-1 ... computed from level i+1:
            // walk level-i array backwards inside each chunk of size L starting from carries[i+1]
            // (reuse the apply kernel: it writes q[j-1] = H_j, i.e. the carry of element j-1)
            k_poly_div_linear_apply<<<(unsigned)((sizes[i + 1] + 127) / 128), 128, 0, c->stream>>>(carries[i], sums[i], carries[i + 1],
                                                                                                  sizes[i], L, ys[i]);
            SWB_LAUNCH_CHECK(c, "k_poly_div_linear_apply");
            // the last element of level i has no successor: its carry is zero
            SWB_CUDA(c, cudaMemsetAsync(carries[i] + (sizes[i] - 1), 0, sizeof(Fr), c->stream));
        }
    }
    k_poly_div_linear_apply<<<(unsigned)(((n + L - 1) / L + 127) / 128), 128, 0, c->stream>>>(q, p, top >= 1 ? carries[1] : nullptr, n, L, z);
    SWB_LAUNCH_CHECK(c, "k_poly_div_linear_apply");
    return SWB_OK;
}

int poly_len_dev(swb_ctx* c, const Fr* p, size_t n, size_t* len) {
    *len = 0;
    if (n == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    unsigned long long* d = (unsigned long long*)get_scratch(c, "poly_len", 64);
    if (!d) return SWB_ENOMEM;
    SWB_CUDA(c, cudaMemsetAsync(d, 0, 8, c->stream));
    k_poly_len<<<grid_for(c, n, 256), 256, 0, c->stream>>>(d, p, n);
    SWB_LAUNCH_CHECK(c, "k_poly_len");
    unsigned long long h = 0;
    SWB_CUDA(c, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    *len = (size_t)h;
    return SWB_OK;
}

int poly_powers_dev(swb_ctx* c, Fr* out, size_t n, const Fr& g) {
    if (n == 0) return SWB_OK;
    SWB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t L = 64;
    const size_t chunks = (n + L - 1) / L;
    k_poly_powers<<<(unsigned)((chunks + 127) / 128), 128, 0, c->stream>>>(out, n, L, g);
    SWB_LAUNCH_CHECK(c, "k_poly_powers");
    return SWB_OK;
}

}  // namespace swb
