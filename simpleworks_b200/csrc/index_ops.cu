// Marlin's indexer on the device: the joint arithmetisation of A, B, C (ark-marlin ahp/constraint_systems.rs:
// sum_matrices + arithmetize_matrix, reached from simpleworks' generate_proving_and_verifying_keys, reference
// src/marlin/mod.rs:88-94) and the column-grouped copy of the matrices that the prover's t polynomial needs.
// The host only flattens the constraint system into three CSR matrices and uploads them; everything else is
// sorts, scans and element-wise kernels the library already owns:
//   1. every matrix entry -> (row, column, which matrix, coefficient)
//   2. stable LSD sort by column, then by row (radix_sort.cu): entries ordered by (row, column)
//   3. heads of equal (row, column) runs + exclusive scan -> position k of the entry in the JOINT sparsity
//      pattern (the per-row sorted union of the column sets); their number is num_non_zero
//   4. one thread per joint entry adds the coefficients of its run into val_A / val_B / val_C
//   5. evaluations on K: row_k = H[pos(column)], col_k = H[row], val_M(k) = M[row][column] * row_k / |H|
//      (u_H(x, x) = |H| / x on H), row_col = row * col; padding entries row = col = H[0], val = 0
//   6. all entries sorted by pos(column) -> the CSR-by-position copy (constraint row, matrix tag, coefficient)
// pos(c) = Radix2EvaluationDomain::reindex_by_subdomain(H, X, c): instance columns sit on the X-subdomain.
#include "index_ops.hpp"

#include "radix_sort.hpp"

namespace swb {

__device__ __forceinline__ uint32_t reindex_pos(uint32_t c, uint32_t nx, uint32_t period) {
    if (c < nx) return c * period;
    const uint32_t i = c - nx, x = period - 1;
    return i + i / x + 1;
}

__device__ __forceinline__ Fr ld_fr_g(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr_g(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

struct Csr3 {
    const uint32_t* start[3];
    const uint32_t* col[3];
    const Fr* coef[3];
    uint32_t first[4];          // global entry ids: matrix m owns [first[m], first[m + 1])
    uint32_t nrows;
};

// global entry g -> (matrix, index inside it)
__device__ __forceinline__ void entry_of(const Csr3& m, uint32_t g, uint32_t* which, uint32_t* e) {
    const uint32_t w = g >= m.first[2] ? 2u : (g >= m.first[1] ? 1u : 0u);
    *which = w;
    *e = g - m.first[w];
}

// step 1: row of every entry (binary search in the row starts) and its column as the first sort key
__global__ void __launch_bounds__(256) k_index_entries(uint32_t* __restrict__ ent_row, uint32_t* __restrict__ key_col,
                                                        uint32_t* __restrict__ val_id, Csr3 m, uint32_t total) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t w, e;
    entry_of(m, g, &w, &e);
    uint32_t lo = 0, hi = m.nrows;                // largest r with start[r] <= e
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (m.start[w][mid] <= e) lo = mid; else hi = mid;
    }
    ent_row[g] = lo;
    key_col[g] = m.col[w][e];
    val_id[g] = g;
}
// second sort key: the row of the entry each sorted slot holds
__global__ void __launch_bounds__(256) k_index_key_rows(uint32_t* __restrict__ key, const uint32_t* __restrict__ ids,
                                                         const uint32_t* __restrict__ ent_row, uint32_t total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) key[i] = ent_row[ids[i]];
}
// step 3: head[i] = 1 when sorted slot i starts a new (row, column); head[total] = 0 so that the scan leaves the count there
__global__ void __launch_bounds__(256) k_index_heads(uint32_t* __restrict__ head, const uint32_t* __restrict__ ids,
                                                      const uint32_t* __restrict__ ent_row, Csr3 m, uint32_t total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > total) return;
    if (i == total) { head[i] = 0; return; }
    if (i == 0) { head[i] = 1; return; }
    const uint32_t g = ids[i], gp = ids[i - 1];
    uint32_t w, e, wp, ep;
    entry_of(m, g, &w, &e);
    entry_of(m, gp, &wp, &ep);
    head[i] = (ent_row[g] != ent_row[gp] || m.col[w][e] != m.col[wp][ep]) ? 1u : 0u;
}
// step 4: the head of every run writes (row, pos(column)) of joint entry k and the three coefficient sums of the run
__global__ void __launch_bounds__(256) k_index_joint(uint32_t* __restrict__ er, uint32_t* __restrict__ epos, Fr* __restrict__ va,
                                                      Fr* __restrict__ vb, Fr* __restrict__ vc, const uint32_t* __restrict__ scan,
                                                      const uint32_t* __restrict__ ids, const uint32_t* __restrict__ ent_row, Csr3 m,
                                                      uint32_t total, uint32_t nx, uint32_t period) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t k = scan[i];
    if (scan[i + 1] == k) return;                 // exclusive scan of the head flags: slot i is a head iff the scan steps after it
    Fr s[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
    uint32_t w, e;
    entry_of(m, ids[i], &w, &e);
    const uint32_t row = ent_row[ids[i]], col = m.col[w][e];
    for (uint32_t j = i; j < total; j++) {
        if (j > i && scan[j + 1] != scan[j]) break;          // the next head
        uint32_t wj, ej;
        entry_of(m, ids[j], &wj, &ej);
        s[wj] = s[wj] + ld_fr_g(m.coef[wj] + ej);
    }
    er[k] = row;
    epos[k] = reindex_pos(col, nx, period);
    st_fr_g(va + k, s[0]);
    st_fr_g(vb + k, s[1]);
    st_fr_g(vc + k, s[2]);
}
// step 5 (in place on va / vb / vc; hel = the elements of H)
__global__ void __launch_bounds__(256) k_index_evals(Fr* __restrict__ row, Fr* __restrict__ col, Fr* __restrict__ va, Fr* __restrict__ vb,
                                                      Fr* __restrict__ vc, Fr* __restrict__ rowcol, const uint32_t* __restrict__ er,
                                                      const uint32_t* __restrict__ epos, const Fr* __restrict__ hel, Fr size_inv,
                                                      uint32_t nnz, uint32_t kn) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kn) return;
    if (k >= nnz) {
        const Fr one = Fr::one();                 // H[0]
        st_fr_g(row + k, one);
        st_fr_g(col + k, one);
        st_fr_g(rowcol + k, one);
        st_fr_g(va + k, Fr::zero());
        st_fr_g(vb + k, Fr::zero());
        st_fr_g(vc + k, Fr::zero());
        return;
    }
    const Fr r = ld_fr_g(hel + epos[k]), c = ld_fr_g(hel + er[k]);
    const Fr scale = r * size_inv;
    st_fr_g(row + k, r);
    st_fr_g(col + k, c);
    st_fr_g(rowcol + k, r * c);
    st_fr_g(va + k, ld_fr_g(va + k) * scale);
    st_fr_g(vb + k, ld_fr_g(vb + k) * scale);
    st_fr_g(vc + k, ld_fr_g(vc + k) * scale);
}
// step 6
__global__ void __launch_bounds__(256) k_index_key_pos(uint32_t* __restrict__ key, uint32_t* __restrict__ val_id, Csr3 m, uint32_t total,
                                                        uint32_t nx, uint32_t period) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t w, e;
    entry_of(m, g, &w, &e);
    key[g] = reindex_pos(m.col[w][e], nx, period);
    val_id[g] = g;
}
__global__ void __launch_bounds__(256) k_index_transposed(uint32_t* __restrict__ t_row, uint8_t* __restrict__ t_tag, Fr* __restrict__ t_coef,
                                                           const uint32_t* __restrict__ ids, const uint32_t* __restrict__ ent_row, Csr3 m,
                                                           uint32_t total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t w, e;
    entry_of(m, ids[i], &w, &e);
    t_row[i] = ent_row[ids[i]];
    t_tag[i] = (uint8_t)w;
    st_fr_g(t_coef + i, ld_fr_g(m.coef[w] + e));
}
// t_start[p] = first sorted slot whose position is >= p, p in [0, nh]
__global__ void __launch_bounds__(256) k_index_bounds(uint32_t* __restrict__ t_start, const uint32_t* __restrict__ sorted_pos, uint32_t total,
                                                       uint32_t nh) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nh) return;
    uint32_t lo = 0, hi = total;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_pos[mid] < p) lo = mid + 1; else hi = mid;
    }
    t_start[p] = lo;
}

static int bits_for(uint32_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) b++;
    return b;
}

int index_joint_dev(swb_ctx* c, const IndexCsr in[3], uint32_t nrows, uint32_t nvar, uint32_t nx, uint32_t nh, IndexJoint* out) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    Csr3 m{};
    uint64_t total64 = 0;
    for (int w = 0; w < 3; w++) {
        m.start[w] = in[w].start; m.col[w] = in[w].col; m.coef[w] = in[w].coef;
        m.first[w] = (uint32_t)total64;
        total64 += in[w].nnz;
    }
    SWB_REQUIRE(c, total64 < ((uint64_t)1 << 31), "index: too many matrix entries");
    const uint32_t total = (uint32_t)total64;
    m.first[3] = total;
    m.nrows = nrows;
    const uint32_t period = nh / nx;
    SWB_REQUIRE(c, period >= 2 || nvar <= nx, "index: no room for witness columns beside the instance subdomain");
    out->total = total;
    out->nnz = 0;
    if (total == 0) return SWB_OK;
    const size_t alt = ((size_t)total + 63) & ~(size_t)63;
    uint32_t* ent_row = (uint32_t*)get_scratch(c, "idx_ent_row", (size_t)total * 4);
    uint32_t* keys = (uint32_t*)get_scratch(c, "idx_keys", alt * 4 * 2);
    uint32_t* vals = (uint32_t*)get_scratch(c, "idx_vals", alt * 4 * 2);
    uint32_t* scan = (uint32_t*)get_scratch(c, "idx_scan", ((size_t)total + 2) * 4);
    if (!ent_row || !keys || !vals || !scan) return SWB_ENOMEM;
    const unsigned grid = (total + 255) / 256;
    k_index_entries<<<grid, 256, 0, c->stream>>>(ent_row, keys, vals, m, total);
    SWB_LAUNCH_CHECK(c, "k_index_entries");
    uint32_t *sk = nullptr, *sv = nullptr;
    int rc = radix_sort_segmented(c, keys, keys + alt, vals, vals + alt, total, 1, bits_for(nvar), &sk, &sv);
    if (rc != SWB_OK) return rc;
    // second LSD digit: the row.  The sorted ids move to the front half if they ended in the back one.
    uint32_t* ids_in = sv;
    uint32_t* key_in = sk;                          // the sorted column keys are no longer needed: rows go in their place
    uint32_t* key_out = (sk == keys) ? keys + alt : keys;
    uint32_t* ids_out = (sv == vals) ? vals + alt : vals;
    k_index_key_rows<<<grid, 256, 0, c->stream>>>(key_in, ids_in, ent_row, total);
    SWB_LAUNCH_CHECK(c, "k_index_key_rows");
    rc = radix_sort_segmented(c, key_in, key_out, ids_in, ids_out, total, 1, bits_for(nrows), &sk, &sv);
    if (rc != SWB_OK) return rc;
    const uint32_t* ids = sv;
    k_index_heads<<<(total + 1 + 255) / 256, 256, 0, c->stream>>>(scan, ids, ent_row, m, total);
    SWB_LAUNCH_CHECK(c, "k_index_heads");
    rc = exclusive_scan_u32(c, scan, (size_t)total + 1);
    if (rc != SWB_OK) return rc;
    uint32_t nnz = 0;
    SWB_CUDA(c, cudaMemcpyAsync(&nnz, scan + total, 4, cudaMemcpyDeviceToHost, c->stream));
    SWB_CUDA(c, cudaStreamSynchronize(c->stream));
    out->nnz = nnz;
    // joint entries; the caller's evaluation vectors are K.n long, K.n = next power of two of nnz
    size_t kn = 1;
    while (kn < nnz) kn <<= 1;
    out->kn = kn;
    uint32_t* er = (uint32_t*)get_scratch(c, "idx_er", (size_t)nnz * 4 + 4);
    uint32_t* epos = (uint32_t*)get_scratch(c, "idx_epos", (size_t)nnz * 4 + 4);
    if (!er || !epos) return SWB_ENOMEM;
    out->er = er;
    out->epos = epos;
    out->ids = ids;
    out->scan = scan;
    out->ent_row = ent_row;
    return SWB_OK;
}

int index_fill_dev(swb_ctx* c, const IndexCsr in[3], uint32_t nrows, uint32_t nx, uint32_t nh, const IndexJoint& j, const Fr* hel,
                   const Fr& size_inv, Fr* row, Fr* col, Fr* va, Fr* vb, Fr* vc, Fr* rowcol, uint32_t* t_start, uint32_t* t_row,
                   uint8_t* t_tag, Fr* t_coef) {
    SWB_CUDA(c, cudaSetDevice(c->device));
    Csr3 m{};
    uint32_t total = 0;
    for (int w = 0; w < 3; w++) {
        m.start[w] = in[w].start; m.col[w] = in[w].col; m.coef[w] = in[w].coef;
        m.first[w] = total;
        total += (uint32_t)in[w].nnz;
    }
    m.first[3] = total;
    m.nrows = nrows;
    const uint32_t period = nh / nx;
    if (total) {
        k_index_joint<<<(total + 255) / 256, 256, 0, c->stream>>>(j.er, j.epos, va, vb, vc, j.scan, j.ids, j.ent_row, m, total, nx, period);
        SWB_LAUNCH_CHECK(c, "k_index_joint");
    }
    k_index_evals<<<(unsigned)((j.kn + 255) / 256), 256, 0, c->stream>>>(row, col, va, vb, vc, rowcol, j.er, j.epos, hel, size_inv, j.nnz,
                                                                          (uint32_t)j.kn);
    SWB_LAUNCH_CHECK(c, "k_index_evals");
    // the column-grouped copy: all entries sorted by the H position of their column
    if (total) {
        const size_t alt = ((size_t)total + 63) & ~(size_t)63;
        uint32_t* keys = (uint32_t*)get_scratch(c, "idx_keys", alt * 4 * 2);
        uint32_t* vals = (uint32_t*)get_scratch(c, "idx_vals", alt * 4 * 2);
        if (!keys || !vals) return SWB_ENOMEM;
        k_index_key_pos<<<(total + 255) / 256, 256, 0, c->stream>>>(keys, vals, m, total, nx, period);
        SWB_LAUNCH_CHECK(c, "k_index_key_pos");
        uint32_t *sk = nullptr, *sv = nullptr;
        int rc = radix_sort_segmented(c, keys, keys + alt, vals, vals + alt, total, 1, bits_for(nh), &sk, &sv);
        if (rc != SWB_OK) return rc;
        k_index_transposed<<<(total + 255) / 256, 256, 0, c->stream>>>(t_row, t_tag, t_coef, sv, j.ent_row, m, total);
        SWB_LAUNCH_CHECK(c, "k_index_transposed");
        k_index_bounds<<<(nh + 1 + 255) / 256, 256, 0, c->stream>>>(t_start, sk, total, nh);
        SWB_LAUNCH_CHECK(c, "k_index_bounds");
    } else {
        SWB_CUDA(c, cudaMemsetAsync(t_start, 0, ((size_t)nh + 1) * 4, c->stream));
    }
    return SWB_OK;
}

}  // namespace swb
