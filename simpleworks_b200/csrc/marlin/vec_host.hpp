// Host implementation of the Engine vector interface used by the protocol templates
// (marlin.hpp): Vec = std::vector<Fr> in RAM, every operation a parallel loop.  The CPU arm
// (oracle build) derives its engine from this; the CUDA engine (marlin_abi.cu) implements the same
// interface on device memory so that polynomials stay in HBM between NTTs and MSMs.
#pragma once
#include "poly.hpp"
#include "rng.hpp"
#include "curve_host.hpp"

namespace swb {
namespace marlin {

struct HostVecOps {
    using Vec = std::vector<Fr>;
    Vec vzeros(size_t n) { return Vec(n, Fr::zero()); }
    Vec vfrom(const std::vector<Fr>& h) { return h; }
    Vec vfrom_ptr(const Fr* p, size_t n) { return Vec(p, p + n); }
    void vwrite(Vec& v, size_t at, const Fr* p, size_t n) { std::copy(p, p + n, v.begin() + at); }   // v[at .. at+n) = p
    std::vector<Fr> vhost(const Vec& v) { return v; }
    Vec vclone(const Vec& v) { return v; }
    void vresize(Vec& v, size_t n) { v.resize(n, Fr::zero()); }
    size_t vlen(const Vec& v) {
        size_t n = v.size();
        while (n > 0 && v[n - 1].is_zero()) n--;
        return n;
    }
    Fr vget(const Vec& v, size_t i) { return v[i]; }
    void vset(Vec& v, size_t i, const Fr& x) { v[i] = x; }
    void vmul(Vec& a, const Vec& b) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < a.size(); i++) a[i] = a[i] * b[i];
    }
    void vadd(Vec& a, const Vec& b) { poly_add(a, b); }
    void vsub(Vec& a, const Vec& b) { poly_sub(a, b); }
    void vadd_scaled(Vec& a, const Fr& c, const Vec& b) { poly_add_scaled(a, c, b); }
    void vscale(Vec& a, const Fr& c) { poly_scale(a, c); }
    void vlin(Vec& a, const Fr& c0, const Fr& c1) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < a.size(); i++) a[i] = c0 + c1 * a[i];
    }
    // a[off + i] += sign * b[i]
    void vadd_offset(Vec& a, size_t off, const Vec& b, bool negate) {
        if (a.size() < off + b.size()) a.resize(off + b.size(), Fr::zero());
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < b.size(); i++) a[off + i] = negate ? a[off + i] - b[i] : a[off + i] + b[i];
    }
    Fr veval(const Vec& p, const Fr& x) { return poly_eval(p, x); }
    void vdiv_vanishing(const Vec& p, size_t n, Vec* q, Vec* r) { poly_divide_by_vanishing(p, n, q, r); }
    Vec vdiv_linear(const Vec& p, const Fr& z) { return poly_divide_by_linear(p, z); }
    void vbatch_inverse(Vec& v) { batch_inverse(v); }
    Vec vshift_down(const Vec& p, size_t k) { return k < p.size() ? Vec(p.begin() + k, p.end()) : Vec(); }
    Vec vdomain(uint32_t log_n) { return Domain((size_t)1 << log_n).elements(); }
    // submit / collect interface of the MSM (asynchronous on the CUDA engine): here the derived engine's
    // msm() runs at submission
    std::vector<G1Point> msm_results;     // results of ids msm_base .. ; old ones are dropped in blocks
    size_t msm_base = 0;
    template <class Derived>
    size_t msm_submit_sync(Derived& self, void* h, size_t offset, const Vec& scalars, size_t n) {
        if (msm_results.size() >= 128) {
            msm_results.erase(msm_results.begin(), msm_results.begin() + 64);
            msm_base += 64;
        }
        msm_results.push_back(n ? self.msm(h, offset, scalars, n) : G1Point::identity());
        return msm_base + msm_results.size() - 1;
    }
    G1Point msm_result(size_t id) { return msm_results.at(id - msm_base); }
    size_t msm_local_count(size_t n) { return n; }
    void msm_drain() {}
    // n consecutive Fr::rand draws (DensePolynomial::rand)
    Vec vrand(ChaChaRng& rng, size_t n) {
        Vec v(n);
        rand_fr_bulk(rng, v.data(), n);
        return v;
    }
    // sparse matrices in CSR form (optionally a 0/1/2 tag per entry selecting one of three weights)
    struct HostCsr {
        std::vector<uint32_t> start, col;
        std::vector<Fr> coef;
        std::vector<uint8_t> tag;
    };
    void* csr_upload(const std::vector<uint32_t>& start, const std::vector<uint32_t>& col, const std::vector<Fr>& coef,
                     const std::vector<uint8_t>* tag) {
        HostCsr* m = new HostCsr();
        m->start = start; m->col = col; m->coef = coef;
        if (tag) m->tag = *tag;
        return m;
    }
    void csr_free(void* h) { delete static_cast<HostCsr*>(h); }
    Vec vspmv(const void* h, const Vec& x, size_t nout, const Fr* weights) {
        const HostCsr& m = *static_cast<const HostCsr*>(h);
        const size_t nrows = m.start.size() - 1;
        Vec out(nout, Fr::zero());
#pragma omp parallel for schedule(static)
        for (size_t r = 0; r < nrows; r++) {
            Fr acc = Fr::zero();
            for (uint32_t k = m.start[r]; k < m.start[r + 1]; k++) {
                Fr term = m.coef[k] * x[m.col[k]];
                if (!m.tag.empty()) term = term * weights[m.tag[k]];
                acc = acc + term;
            }
            out[r] = acc;
        }
        return out;
    }
    // w on H before interpolation: 0 on the X-subdomain, witness - x^ elsewhere (z = instance | witness)
    Vec vwitness_evals(const Vec& z, size_t ninst, const Vec& xh_on_h, size_t ratio) {
        const size_t nh = xh_on_h.size();
        Vec out(nh, Fr::zero());
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < nh; k++) {
            if (k % ratio == 0) continue;
            const size_t wi = ninst + (k - k / ratio - 1);
            out[k] = (wi < z.size() ? z[wi] : Fr::zero()) - xh_on_h[k];
        }
        return out;
    }
};

}  // namespace marlin
}  // namespace swb
