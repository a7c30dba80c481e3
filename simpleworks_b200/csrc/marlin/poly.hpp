// Host-side dense polynomials and radix-2 evaluation domains for the protocol layer.
// Restates the pieces of ark-poly 0.3 that Marlin's AHP uses around the NTT kernel:
// DensePolynomial (add / scale / evaluate / divide_by_vanishing_poly / division by X - z),
// Radix2EvaluationDomain (element, vanishing polynomial, reindex_by_subdomain) and the marlin
// crate's EvaluationDomainExt bivariate helper.  The transforms themselves are
// delegated to the Engine (GPU in the product, CPU in the oracle build).
#pragma once
#include <stdexcept>
#include <algorithm>
#include <vector>

#include "../fp.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

namespace swb {
namespace marlin {

// SWB_TRACE=1: wall-clock breakdown of the host protocol code by phase (stderr)
struct HostProfile {
    std::map<std::string, double> acc;
    bool on = false;        // phases are being accumulated
    bool print = false;     // SWB_TRACE: each report goes to stderr
    std::string last;       // "<what>: name=ms name=ms ..." of the most recent report
    HostProfile() { const char* e = getenv("SWB_TRACE"); on = print = e && atoi(e) > 0; }
    void report(const char* what) {
        if (!on) return;
        last = what;
        last += ":";
        char buf[96];
        for (auto& kv : acc) {
            snprintf(buf, sizeof buf, " %s=%.3f", kv.first.c_str(), kv.second * 1e3);
            last += buf;
        }
        if (print) fprintf(stderr, "[swb trace] %s host phases (ms)%s\n", what, last.c_str() + strlen(what) + 1);
        acc.clear();
    }
};
inline HostProfile& host_profile() { static HostProfile p; return p; }
struct ScopedPhase {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    explicit ScopedPhase(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~ScopedPhase() {
        if (!host_profile().on) return;
        host_profile().acc[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};

using Poly = std::vector<Fr>;   // coefficients, low degree first; may carry leading zeros

inline Fr fr_from_u64(uint64_t v) {
    Fr c = Fr::zero();
    c.l[0] = (uint32_t)v;
    c.l[1] = (uint32_t)(v >> 32);
    return c.from_canonical();
}
inline Fr fr_pow(Fr base, uint64_t e) { return base.pow_u64(e); }

inline void poly_trim(Poly& p) {
    while (!p.empty() && p.back().is_zero()) p.pop_back();
}
inline size_t poly_degree(const Poly& p) {
    size_t n = p.size();
    while (n > 0 && p[n - 1].is_zero()) n--;
    return n ? n - 1 : 0;
}
inline Fr poly_eval(const Poly& p, const Fr& x) {
    // Horner in parallel chunks: sum_c x^(c*L) * chunk_c(x)
    const size_t n = p.size();
    if (n == 0) return Fr::zero();
    const size_t L = 1 << 12;
    const size_t nchunks = (n + L - 1) / L;
    std::vector<Fr> part(nchunks);
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < nchunks; c++) {
        size_t lo = c * L, hi = std::min(n, lo + L);
        Fr acc = Fr::zero();
        for (size_t i = hi; i-- > lo;) acc = acc * x + p[i];
        part[c] = acc;
    }
    const Fr xl = fr_pow(x, L);
    Fr acc = Fr::zero();
    for (size_t c = nchunks; c-- > 0;) acc = acc * xl + part[c];
    return acc;
}
// r = a + c * b
inline void poly_add_scaled(Poly& a, const Fr& c, const Poly& b) {
    if (a.size() < b.size()) a.resize(b.size(), Fr::zero());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < b.size(); i++) a[i] = a[i] + c * b[i];
}
inline void poly_add(Poly& a, const Poly& b) {
    if (a.size() < b.size()) a.resize(b.size(), Fr::zero());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < b.size(); i++) a[i] = a[i] + b[i];
}
inline void poly_sub(Poly& a, const Poly& b) {
    if (a.size() < b.size()) a.resize(b.size(), Fr::zero());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < b.size(); i++) a[i] = a[i] - b[i];
}
inline void poly_scale(Poly& a, const Fr& c) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < a.size(); i++) a[i] = a[i] * c;
}
// p = q * (X^n - 1) + r  (DensePolynomial::divide_by_vanishing_poly)
inline void poly_divide_by_vanishing(const Poly& p, size_t n, Poly* q, Poly* r) {
    r->assign(n, Fr::zero());
    if (p.size() <= n) {
        q->clear();
        std::copy(p.begin(), p.end(), r->begin());
        return;
    }
    q->assign(p.size() - n, Fr::zero());
    // q[i] = sum_{k>=1} p[i + k n]: one backward pass, q[i] = p[i+n] + q[i+n]
    for (size_t i = q->size(); i-- > 0;) {
        Fr v = p[i + n];
        if (i + n < q->size()) v = v + (*q)[i + n];
        (*q)[i] = v;
    }
    for (size_t i = 0; i < n; i++) {
        Fr v = p[i];
        if (i < q->size()) v = v + (*q)[i];
        (*r)[i] = v;
    }
}
// (p - p(z)) / (X - z), synthetic division
inline Poly poly_divide_by_linear(const Poly& p, const Fr& z) {
    const size_t n = p.size();
    if (n <= 1) return Poly();
    Poly q(n - 1);
    Fr carry = p[n - 1];
    q[n - 2] = carry;
    for (size_t i = n - 1; i-- > 1;) {
        carry = p[i] + z * carry;
        q[i - 1] = carry;
    }
    return q;
}

struct Domain {
    uint32_t log_n = 0;
    size_t n = 1;
    Fr gen, gen_inv, size_inv, size_fr;
    Domain() : Domain(1) {}
    // Radix2EvaluationDomain::new returns None beyond the field's two-adicity (2^47); here: throws
    explicit Domain(size_t min_size) {
        if (min_size > ((size_t)1 << SWB_FR_TWO_ADICITY)) throw std::length_error("evaluation domain larger than 2^47");
        log_n = 0;
        while (((size_t)1 << log_n) < min_size) log_n++;
        n = (size_t)1 << log_n;
        const uint32_t root[8] = SWB_FR_ROOT_OF_UNITY_INIT;
        for (int i = 0; i < 8; i++) gen.l[i] = root[i];
        for (uint32_t i = log_n; i < SWB_FR_TWO_ADICITY; i++) gen = gen.sqr();
        gen_inv = gen.inverse();
        size_fr = fr_from_u64(n);
        size_inv = size_fr.inverse();
    }
    Fr element(size_t i) const { return fr_pow(gen, i); }
    std::vector<Fr> elements() const {
        std::vector<Fr> e(n);
        const size_t L = 1 << 12;
#pragma omp parallel for schedule(static)
        for (size_t c = 0; c < (n + L - 1) / L; c++) {
            size_t lo = c * L, hi = std::min(n, lo + L);
            Fr v = fr_pow(gen, lo);
            for (size_t i = lo; i < hi; i++) { e[i] = v; v = v * gen; }
        }
        return e;
    }
    Fr vanishing_at(const Fr& x) const { return fr_pow(x, n) - Fr::one(); }
    // EvaluationDomainExt::eval_unnormalized_bivariate_lagrange_poly(x, y) = (v(x) - v(y)) / (x - y)
    Fr bivariate(const Fr& x, const Fr& y) const {
        if (x == y) return size_fr * fr_pow(x, n - 1);
        return (vanishing_at(x) - vanishing_at(y)) * (x - y).inverse();
    }
    // Radix2EvaluationDomain::reindex_by_subdomain(other, index)
    size_t reindex_by_subdomain(const Domain& other, size_t index) const {
        const size_t period = n / other.n;
        if (index < other.n) return index * period;
        const size_t i = index - other.n, x = period - 1;
        return i + i / x + 1;
    }
    // all Lagrange coefficients L_i(tau), i < n  (evaluate_all_lagrange_coefficients)
    std::vector<Fr> lagrange_at(const Fr& tau) const {
        std::vector<Fr> out(n);
        const Fr z = vanishing_at(tau);
        if (z.is_zero()) {
            Fr e = Fr::one();
            for (size_t i = 0; i < n; i++) { out[i] = (e == tau) ? Fr::one() : Fr::zero(); e = e * gen; }
            return out;
        }
        // L_i(tau) = z * w^i / (n * (tau - w^i))
        Fr e = Fr::one();
        const Fr zn = z * size_inv;
        for (size_t i = 0; i < n; i++) { out[i] = zn * e * (tau - e).inverse(); e = e * gen; }
        return out;
    }
};

// Montgomery batch inversion in parallel chunks (ark_ff::batch_inversion); zeros stay zero
inline void batch_inverse(std::vector<Fr>& v) {
    const size_t n = v.size(), L = 1 << 10;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < (n + L - 1) / L; c++) {
        size_t lo = c * L, hi = std::min(n, lo + L);
        std::vector<Fr> pre(hi - lo);
        Fr acc = Fr::one();
        for (size_t i = lo; i < hi; i++) {
            pre[i - lo] = acc;
            if (!v[i].is_zero()) acc = acc * v[i];
        }
        acc = acc.inverse();
        for (size_t i = hi; i-- > lo;) {
            if (v[i].is_zero()) continue;
            Fr x = v[i];
            v[i] = acc * pre[i - lo];
            acc = acc * x;
        }
    }
}

}  // namespace marlin
}  // namespace swb
