/* TEST INFRASTRUCTURE ONLY -- see swb_oracle.h.  CPU restatement of the arkworks 0.3 algorithms
 * on the Marlin hot path.  PARITY UNPINNED vs real arkworks; pinned
 * against oracle/golden.py and tests/golden/.  Never linked into libswb200. */
#include "swb_oracle.h"
#include "orc_constants.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- field instantiations ----------------------------------------------------------------- */
const uint64_t fr_MOD[4] = ORC_FR_MOD;
const uint64_t fr_R1[4] = ORC_FR_R1;
const uint64_t fr_R2[4] = ORC_FR_R2;
const uint64_t fr_INV = ORC_FR_INV;
const uint64_t fq_MOD[6] = ORC_FQ_MOD;
const uint64_t fq_R1[6] = ORC_FQ_R1;
const uint64_t fq_R2[6] = ORC_FQ_R2;
const uint64_t fq_INV = ORC_FQ_INV;

#define FP fr
#define FPN 4
#include "fp_tmpl.h"
#undef FP
#undef FPN
#define FP fq
#define FPN 6
#include "fp_tmpl.h"
#undef FP
#undef FPN

static const uint64_t FR_ROOT[4] = ORC_FR_ROOT_OF_UNITY;
static const uint64_t FR_GEN[4] = ORC_FR_GENERATOR;
static const uint64_t FR_GEN_INV[4] = ORC_FR_GENERATOR_INV;
static const uint64_t G1X[6] = ORC_FQ_G1_GEN_X;
static const uint64_t G1Y[6] = ORC_FQ_G1_GEN_Y;

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static int pick_threads(int t) { return t > 0 ? t : orc_num_threads(); }

void orc_fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { fr_mul(r, a, b); }
void orc_fr_add(fr_t *r, const fr_t *a, const fr_t *b) { fr_add(r, a, b); }
void orc_fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { fr_sub(r, a, b); }
void orc_fr_inv(fr_t *r, const fr_t *a) { fr_inv(r, a); }
void orc_fr_from_canon(fr_t *r, const big256_t *a) { fr_from_canon(r, a->l); }
void orc_fr_to_canon(big256_t *r, const fr_t *a) { fr_to_canon(r->l, a); }
void orc_fq_mul(fq_t *r, const fq_t *a, const fq_t *b) { fq_mul(r, a, b); }
void orc_fq_add(fq_t *r, const fq_t *a, const fq_t *b) { fq_add(r, a, b); }
void orc_fq_sub(fq_t *r, const fq_t *a, const fq_t *b) { fq_sub(r, a, b); }
void orc_fq_inv(fq_t *r, const fq_t *a) { fq_inv(r, a); }
void orc_fq_from_canon(fq_t *r, const uint64_t a[6]) { fq_from_canon(r, a); }
void orc_fq_to_canon(uint64_t r[6], const fq_t *a) { fq_to_canon(r, a); }

void orc_fr_mul_vec(fr_t *r, const fr_t *a, const fr_t *b, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_mul(&r[i], &a[i], &b[i]);
}
void orc_fq_mul_vec(fq_t *r, const fq_t *a, const fq_t *b, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fq_mul(&r[i], &a[i], &b[i]);
}

/* ark_ff::batch_inversion: Montgomery's trick, zero entries left untouched */
void orc_fr_batch_inverse(fr_t *v, size_t n) {
    fr_t *prod = (fr_t *)malloc(sizeof(fr_t) * (n + 1));
    fr_t acc;
    size_t cnt = 0;
    fr_one(&acc);
    prod[cnt++] = acc;                       /* prod[k] = product of the first k non-zero entries */
    for (size_t i = 0; i < n; i++) {
        if (fr_is_zero(&v[i])) continue;
        fr_mul(&acc, &acc, &v[i]);
        prod[cnt++] = acc;
    }
    fr_inv(&acc, &acc);
    size_t k = cnt - 1;
    for (size_t i = n; i-- > 0;) {
        if (fr_is_zero(&v[i])) continue;
        fr_t newacc, inv_i;
        fr_mul(&newacc, &acc, &v[i]);
        fr_mul(&inv_i, &acc, &prod[k - 1]);
        k--;
        v[i] = inv_i;
        acc = newacc;
    }
    free(prod);
}

/* ---- G1 ------------------------------------------------------------------------------------ */
void orc_g1_generator(g1_affine_t *g) {
    memset(g, 0, sizeof *g);
    memcpy(g->x.l, G1X, sizeof G1X);
    memcpy(g->y.l, G1Y, sizeof G1Y);
}
void orc_g1_jac_zero(g1_jac_t *r) {
    /* GroupProjective::zero() = (0, 1, 0) */
    memset(r, 0, sizeof *r);
    fq_one(&r->y);
}
static int jac_is_zero(const g1_jac_t *p) { return fq_is_zero(&p->z); }
void orc_g1_from_affine(g1_jac_t *r, const g1_affine_t *a) {
    if (a->infinity) { orc_g1_jac_zero(r); return; }
    r->x = a->x;
    r->y = a->y;
    fq_one(&r->z);
}
/* dbl-2009-l, a = 0 (double_in_place) */
void orc_g1_double(g1_jac_t *p) {
    if (jac_is_zero(p)) return;
    fq_t A, B, C, D, E, Fv, t;
    fq_sqr(&A, &p->x);
    fq_sqr(&B, &p->y);
    fq_sqr(&C, &B);
    fq_add(&t, &p->x, &B);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &A);
    fq_sub(&t, &t, &C);
    fq_dbl(&D, &t);
    fq_dbl(&E, &A);
    fq_add(&E, &E, &A);
    fq_sqr(&Fv, &E);
    fq_mul(&t, &p->y, &p->z);
    fq_dbl(&p->z, &t);
    fq_sub(&p->x, &Fv, &D);
    fq_sub(&p->x, &p->x, &D);
    fq_sub(&t, &D, &p->x);
    fq_mul(&t, &E, &t);
    fq_dbl(&C, &C);
    fq_dbl(&C, &C);
    fq_dbl(&C, &C);
    fq_sub(&p->y, &t, &C);
}
/* madd-2007-bl (add_assign_mixed) */
void orc_g1_add_mixed(g1_jac_t *p, const g1_affine_t *q) {
    if (q->infinity) return;
    if (jac_is_zero(p)) { orc_g1_from_affine(p, q); return; }
    fq_t z1z1, u2, s2, h, hh, i, j, r, v, t;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&p->x, &u2) && fq_eq(&p->y, &s2)) { orc_g1_double(p); return; }
    fq_sub(&h, &u2, &p->x);
    fq_sqr(&hh, &h);
    fq_dbl(&i, &hh);
    fq_dbl(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &p->y);
    fq_dbl(&r, &r);
    fq_mul(&v, &p->x, &i);
    /* Z3 = (Z1+H)^2 - Z1Z1 - HH */
    fq_add(&t, &p->z, &h);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &z1z1);
    fq_sub(&p->z, &t, &hh);
    /* X3 = r^2 - J - 2V */
    fq_sqr(&p->x, &r);
    fq_sub(&p->x, &p->x, &j);
    fq_sub(&p->x, &p->x, &v);
    fq_sub(&p->x, &p->x, &v);
    /* Y3 = r (V - X3) - 2 Y1 J */
    fq_mul(&j, &p->y, &j);
    fq_dbl(&j, &j);
    fq_sub(&v, &v, &p->x);
    fq_mul(&v, &v, &r);
    fq_sub(&p->y, &v, &j);
}
/* add-2007-bl (add_assign) */
void orc_g1_add(g1_jac_t *p, const g1_jac_t *q) {
    if (jac_is_zero(p)) { *p = *q; return; }
    if (jac_is_zero(q)) return;
    fq_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    fq_sqr(&z1z1, &p->z);
    fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z);
    fq_mul(&s1, &s1, &z2z2);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&u1, &u2) && fq_eq(&s1, &s2)) { orc_g1_double(p); return; }
    fq_sub(&h, &u2, &u1);
    fq_dbl(&i, &h);
    fq_sqr(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &s1);
    fq_dbl(&r, &r);
    fq_mul(&v, &u1, &i);
    fq_add(&t, &p->z, &q->z);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &z1z1);
    fq_sub(&t, &t, &z2z2);
    fq_mul(&p->z, &t, &h);
    fq_sqr(&p->x, &r);
    fq_sub(&p->x, &p->x, &j);
    fq_sub(&p->x, &p->x, &v);
    fq_sub(&p->x, &p->x, &v);
    fq_mul(&s1, &s1, &j);
    fq_dbl(&s1, &s1);
    fq_sub(&v, &v, &p->x);
    fq_mul(&v, &v, &r);
    fq_sub(&p->y, &v, &s1);
}
void orc_g1_to_affine(g1_affine_t *r, const g1_jac_t *p) {
    memset(r, 0, sizeof *r);
    if (jac_is_zero(p)) { r->infinity = 1; return; }   /* GroupAffine::zero() = (0,0,true)... */
    fq_t zi, zi2, zi3;
    fq_inv(&zi, &p->z);
    fq_sqr(&zi2, &zi);
    fq_mul(&zi3, &zi2, &zi);
    fq_mul(&r->x, &p->x, &zi2);
    fq_mul(&r->y, &p->y, &zi3);
}
int orc_g1_affine_on_curve(const g1_affine_t *a) {
    if (a->infinity) return 1;
    fq_t y2, x3, one;
    fq_sqr(&y2, &a->y);
    fq_sqr(&x3, &a->x);
    fq_mul(&x3, &x3, &a->x);
    fq_one(&one);
    fq_add(&x3, &x3, &one);
    return fq_eq(&y2, &x3);
}
void orc_g1_mul(g1_jac_t *r, const g1_affine_t *base, const big256_t *k) {
    g1_jac_t acc;
    orc_g1_jac_zero(&acc);
    for (int i = 255; i >= 0; i--) {
        orc_g1_double(&acc);
        if ((k->l[i / 64] >> (i % 64)) & 1) orc_g1_add_mixed(&acc, base);
    }
    *r = acc;
}
/* ProjectiveCurve::batch_normalization_into_affine: one batched inversion of all Z */
void orc_g1_batch_normalize(g1_affine_t *out, const g1_jac_t *in, size_t n) {
    fq_t *prod = (fq_t *)malloc(sizeof(fq_t) * (n ? n : 1));
    fq_t acc;
    fq_one(&acc);
    for (size_t i = 0; i < n; i++) {
        prod[i] = acc;                       /* product of non-zero z before i */
        if (!jac_is_zero(&in[i])) fq_mul(&acc, &acc, &in[i].z);
    }
    fq_inv(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        memset(&out[i], 0, sizeof out[i]);
        if (jac_is_zero(&in[i])) { out[i].infinity = 1; continue; }
        fq_t zi, zi2, zi3;
        fq_mul(&zi, &acc, &prod[i]);
        fq_mul(&acc, &acc, &in[i].z);
        fq_sqr(&zi2, &zi);
        fq_mul(&zi3, &zi2, &zi);
        fq_mul(&out[i].x, &in[i].x, &zi2);
        fq_mul(&out[i].y, &in[i].y, &zi3);
    }
    free(prod);
}

/* ---- VariableBaseMSM::multi_scalar_mul (ark-ec/src/msm/variable_base.rs) ------------------- */
static uint32_t log2_ceil(size_t x) {           /* ark_std::log2 */
    if (x <= 1) return 0;
    uint32_t l = 0;
    size_t v = x - 1;
    while (v) { l++; v >>= 1; }
    return l;
}
static size_t ln_without_floats(size_t a) { return (size_t)log2_ceil(a) * 69 / 100; }

static int big_is_zero(const big256_t *s) { return (s->l[0] | s->l[1] | s->l[2] | s->l[3]) == 0; }
static int big_is_one(const big256_t *s) { return s->l[0] == 1 && (s->l[1] | s->l[2] | s->l[3]) == 0; }
static uint64_t big_window(const big256_t *s, size_t start, size_t c) {   /* divn + low limb % 2^c */
    size_t limb = start / 64, off = start % 64;
    uint64_t v = s->l[limb] >> off;
    if (off && limb + 1 < 4) v |= s->l[limb + 1] << (64 - off);
    return c >= 64 ? v : (v & ((1ull << c) - 1));
}

void orc_msm_variable_base(g1_jac_t *out, const g1_affine_t *bases, const big256_t *scalars,
                           size_t n, int threads) {
    const size_t c = n < 32 ? 3 : ln_without_floats(n) + 2;
    const size_t num_bits = 253;                 /* FrParameters::MODULUS_BITS */
    const size_t nwin = (num_bits + c - 1) / c;
    g1_jac_t *window_sums = (g1_jac_t *)malloc(sizeof(g1_jac_t) * nwin);
    threads = pick_threads(threads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (size_t w = 0; w < nwin; w++) {
        const size_t w_start = w * c;
        g1_jac_t res;
        orc_g1_jac_zero(&res);
        const size_t nb = ((size_t)1 << c) - 1;
        g1_jac_t *buckets = (g1_jac_t *)malloc(sizeof(g1_jac_t) * nb);
        for (size_t b = 0; b < nb; b++) orc_g1_jac_zero(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            if (big_is_zero(&scalars[i])) continue;
            if (big_is_one(&scalars[i])) {
                if (w_start == 0) orc_g1_add_mixed(&res, &bases[i]);
            } else {
                uint64_t d = big_window(&scalars[i], w_start, c);
                if (d) orc_g1_add_mixed(&buckets[d - 1], &bases[i]);
            }
        }
        g1_jac_t running;
        orc_g1_jac_zero(&running);
        for (size_t b = nb; b-- > 0;) {
            orc_g1_add(&running, &buckets[b]);
            orc_g1_add(&res, &running);
        }
        free(buckets);
        window_sums[w] = res;
    }
    g1_jac_t total;
    orc_g1_jac_zero(&total);
    for (size_t w = nwin; w-- > 1;) {
        orc_g1_add(&total, &window_sums[w]);
        for (size_t k = 0; k < c; k++) orc_g1_double(&total);
    }
    g1_jac_t lowest = window_sums[0];
    orc_g1_add(&lowest, &total);
    *out = lowest;
    free(window_sums);
}

/* ---- FixedBaseMSM (ark-ec/src/msm/fixed_base.rs) as used by KZG10::setup ------------------- */
static size_t fixed_window_size(size_t num_scalars) {   /* get_mul_window_size */
    return num_scalars < 32 ? 3 : ln_without_floats(num_scalars);
}
void orc_fixed_base_powers(g1_affine_t *out, const g1_jac_t *g, const fr_t *beta, size_t n, int threads) {
    if (n == 0) return;
    threads = pick_threads(threads);
    const size_t scalar_bits = 253;
    const size_t window = fixed_window_size(n);
    const size_t in_window = (size_t)1 << window;
    const size_t outerc = (scalar_bits + window - 1) / window;
    const size_t last_in_window = (size_t)1 << (scalar_bits - (outerc - 1) * window);
    /* get_window_table: table[outer][inner] = inner * 2^(outer*window) * g, batch-normalised */
    g1_jac_t *tbl_j = (g1_jac_t *)malloc(sizeof(g1_jac_t) * outerc * in_window);
    g1_jac_t g_outer = *g;
    for (size_t o = 0; o < outerc; o++) {
        g1_jac_t g_inner;
        orc_g1_jac_zero(&g_inner);
        size_t cur = (o == outerc - 1) ? last_in_window : in_window;
        for (size_t i = 0; i < in_window; i++) {
            if (i < cur) {
                tbl_j[o * in_window + i] = g_inner;
                orc_g1_add(&g_inner, &g_outer);
            } else {
                orc_g1_jac_zero(&tbl_j[o * in_window + i]);
            }
        }
        for (size_t k = 0; k < window; k++) orc_g1_double(&g_outer);
    }
    g1_affine_t *tbl = (g1_affine_t *)malloc(sizeof(g1_affine_t) * outerc * in_window);
    orc_g1_batch_normalize(tbl, tbl_j, outerc * in_window);
    free(tbl_j);
    /* powers of beta (sequential, like KZG10::setup) */
    fr_t *pw = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_one(&pw[0]);
    for (size_t i = 1; i < n; i++) fr_mul(&pw[i], &pw[i - 1], beta);
    g1_jac_t *res = (g1_jac_t *)malloc(sizeof(g1_jac_t) * n);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) {
        big256_t s;
        fr_to_canon(s.l, &pw[i]);
        g1_jac_t acc;                      /* windowed_mul: res = table[0][w0]; res += table[o][wo] */
        orc_g1_jac_zero(&acc);
        for (size_t o = 0; o < outerc; o++) {
            uint64_t d = big_window(&s, o * window, window);
            if (o * window + window > scalar_bits) d &= (last_in_window - 1);
            orc_g1_add_mixed(&acc, &tbl[o * in_window + d]);
        }
        res[i] = acc;
    }
    /* batch normalisation in chunks so it parallelises */
    const size_t chunk = 4096;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t s = 0; s < (n + chunk - 1) / chunk; s++) {
        size_t lo = s * chunk, hi = lo + chunk > n ? n : lo + chunk;
        orc_g1_batch_normalize(out + lo, res + lo, hi - lo);
    }
    free(res);
    free(pw);
    free(tbl);
}

/* ---- Radix2EvaluationDomain ----------------------------------------------------------------- */
void orc_domain_generator(fr_t *w, uint32_t log_n) {
    /* group_gen = ROOT_OF_UNITY ^ (2^(TWO_ADICITY - log_n)) */
    fr_t g;
    memcpy(g.l, FR_ROOT, sizeof FR_ROOT);
    for (uint32_t i = log_n; i < ORC_FR_TWO_ADICITY; i++) fr_sqr(&g, &g);
    *w = g;
}
static void distribute_powers(fr_t *v, size_t n, const fr_t *g, int threads) {
    /* v[j] *= g^j */
    const size_t chunk = 1 << 12;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t s = 0; s < (n + chunk - 1) / chunk; s++) {
        size_t lo = s * chunk, hi = lo + chunk > n ? n : lo + chunk;
        uint64_t e[1] = {lo};
        fr_t p;
        fr_pow(&p, g, e, 1);
        for (size_t j = lo; j < hi; j++) {
            fr_mul(&v[j], &v[j], &p);
            fr_mul(&p, &p, g);
        }
    }
}
void orc_ntt(fr_t *v, uint32_t log_n, int inverse, int coset, int threads) {
    const size_t n = (size_t)1 << log_n;
    threads = pick_threads(threads);
    fr_t w;
    orc_domain_generator(&w, log_n);
    if (inverse) fr_inv(&w, &w);
    if (coset && !inverse) {
        fr_t g;
        memcpy(g.l, FR_GEN, sizeof FR_GEN);
        distribute_powers(v, n, &g, threads);
    }
    /* bit reversal (derange) then in-place decimation-in-time butterflies */
    for (size_t i = 0; i < n; i++) {
        size_t r = 0;
        for (uint32_t b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < r) { fr_t t = v[i]; v[i] = v[r]; v[r] = t; }
    }
    /* roots[k] = w^k, k < n/2 */
    size_t half = n / 2 ? n / 2 : 1;
    fr_t *roots = (fr_t *)malloc(sizeof(fr_t) * half);
    fr_one(&roots[0]);
    for (size_t k = 1; k < half; k++) fr_mul(&roots[k], &roots[k - 1], &w);
    for (uint32_t s = 1; s <= log_n; s++) {
        const size_t m = (size_t)1 << s, hm = m >> 1, stride = n / m;
#pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t idx = 0; idx < n / 2; idx++) {
            size_t blk = idx / hm, j = idx % hm;
            fr_t *lo = &v[blk * m + j], *hi = lo + hm, t;
            fr_mul(&t, hi, &roots[j * stride]);
            fr_sub(hi, lo, &t);
            fr_add(lo, lo, &t);
        }
    }
    free(roots);
    if (inverse) {
        fr_t ninv;
        fr_from_u64(&ninv, (uint64_t)n);
        fr_inv(&ninv, &ninv);
#pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t i = 0; i < n; i++) fr_mul(&v[i], &v[i], &ninv);
        if (coset) {
            fr_t gi;
            memcpy(gi.l, FR_GEN_INV, sizeof FR_GEN_INV);
            distribute_powers(v, n, &gi, threads);
        }
    }
}
