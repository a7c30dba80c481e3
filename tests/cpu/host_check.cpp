// CPU check of the device code paths that are also valid host code:
//   * the 32-bit-limb Montgomery schedule (run through the emulated carry flag),
//   * XYZZ point arithmetic and the signed-digit bucket method as the kernels use them,
// against the C oracle.  Built and run by tests/test_host_schedule.py (no GPU needed).
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "g1.cuh"
#include "swb_oracle.h"

using namespace swb;

static Fq fq_from(const fq_t& a) { Fq r; memcpy(r.l, a.l, 48); return r; }
static fq_t fq_to(const Fq& a) { fq_t r; memcpy(r.l, a.l, 48); return r; }

static bool same_point(const G1Xyzz& p, const g1_jac_t& j) {
    g1_affine_t a;
    orc_g1_to_affine(&a, &j);
    if (p.is_identity()) return a.infinity;
    if (a.infinity) return false;
    Fq inv = (p.zz * p.zzz).inverse();
    Fq x = p.x * (p.zzz * inv), y = p.y * (p.zz * inv);
    return memcmp(x.l, a.x.l, 48) == 0 && memcmp(y.l, a.y.l, 48) == 0;
}

int main() {
    std::mt19937_64 rng(7);
    int bad = 0;
    // 1. limb schedule vs oracle vs host64
    for (int it = 0; it < 100000; it++) {
        fq_t a, b, r;
        for (int i = 0; i < 6; i++) { a.l[i] = rng(); b.l[i] = rng(); }
        a.l[5] &= 0x00ffffffffffffffull; b.l[5] &= 0x00ffffffffffffffull;
        orc_fq_mul(&r, &a, &b);
        Fq x = fq_from(a), y = fq_from(b);
        Fq z1 = Fq::mul_limb_schedule(x, y), z2 = Fq::mul_host64(x, y);
        if (memcmp(z1.l, r.l, 48) || memcmp(z2.l, r.l, 48)) bad++;
        fr_t c, d, e;
        for (int i = 0; i < 4; i++) { c.l[i] = rng(); d.l[i] = rng(); }
        c.l[3] &= 0x0fffffffffffffffull; d.l[3] &= 0x0fffffffffffffffull;
        orc_fr_mul(&e, &c, &d);
        Fr u, v; memcpy(u.l, c.l, 32); memcpy(v.l, d.l, 32);
        Fr w1 = Fr::mul_limb_schedule(u, v), w2 = Fr::mul_host64(u, v);
        if (memcmp(w1.l, e.l, 32) || memcmp(w2.l, e.l, 32)) bad++;
    }
    printf("field mismatches: %d\n", bad);
    // 2. XYZZ arithmetic vs oracle Jacobian
    g1_affine_t g; orc_g1_generator(&g);
    std::vector<g1_affine_t> pts(64);
    for (size_t i = 0; i < pts.size(); i++) {
        big256_t k = {{rng(), rng(), rng(), rng() & 0x0fffffffffffffffull}};
        g1_jac_t j; orc_g1_mul(&j, &g, &k); orc_g1_to_affine(&pts[i], &j);
    }
    G1Xyzz acc = G1Xyzz::identity();
    g1_jac_t jac; orc_g1_jac_zero(&jac);
    for (size_t i = 0; i < pts.size(); i++) {
        acc.add_affine(fq_from(pts[i].x), fq_from(pts[i].y));
        orc_g1_add_mixed(&jac, &pts[i]);
        if (!same_point(acc, jac)) bad++;
    }
    // doubling through add_affine of the same point, cancellation, xyzz+xyzz, dbl
    G1Xyzz d2 = G1Xyzz::identity();
    d2.add_affine(fq_from(pts[0].x), fq_from(pts[0].y));
    d2.add_affine(fq_from(pts[0].x), fq_from(pts[0].y));
    g1_jac_t j2; orc_g1_from_affine(&j2, &pts[0]); orc_g1_double(&j2);
    if (!same_point(d2, j2)) bad++;
    G1Xyzz canc = d2; canc.add_affine(fq_from(pts[0].x), fq_from(pts[0].y).neg());
    canc.add_affine(fq_from(pts[0].x), fq_from(pts[0].y).neg());
    if (!canc.is_identity()) bad++;
    G1Xyzz s = acc; s.add(d2);
    g1_jac_t js = jac; orc_g1_add(&js, &j2);
    if (!same_point(s, js)) bad++;
    G1Xyzz s2 = acc; s2.add(acc);
    g1_jac_t js2 = jac; orc_g1_double(&js2);
    if (!same_point(s2, js2)) bad++;
    if (!same_point(acc.dbl(), js2)) bad++;
    // 4. binary-GCD inversion (msm_pairs.cu inverts with it) against the Fermat inverse, both fields, edge values
    {
        int before = bad;
        for (int it = 0; it < 3000; it++) {
            Fq a;
            for (int i = 0; i < 12; i++) a.l[i] = (uint32_t)rng();
            a.l[11] &= 0x00ffffffu;
            if (it == 0) a = Fq::one();
            if (it == 1) a = Fq::one().neg();
            if (it == 2) { a = Fq::zero(); a.l[0] = 1; }          // raw limb 1 = R^-1
            if (it == 3) a = Fq::one() + Fq::one();
            const Fq i1 = a.inverse(), i2 = a.inverse_bingcd();
            if (!(i1 == i2) || !((a * i2) == Fq::one())) bad++;
            Fr b;
            for (int i = 0; i < 8; i++) b.l[i] = (uint32_t)rng();
            b.l[7] &= 0x0fffffffu;
            const Fr j1 = b.inverse(), j2 = b.inverse_bingcd();
            if (!(j1 == j2)) bad++;
        }
        if (!Fq::zero().inverse_bingcd().is_zero()) bad++;
        printf("inverse mismatches: %d\n", bad - before);
    }
    // lazy range [0, 2p) used by the NTT butterflies: add_lazy / sub_lazy / mul_lazy / final_sub2 against the canonical
    // operators, on values pushed to the top of their ranges (a + p < 2p, u - v + 2p close to 4p)
    {
        int before = bad;
        Fr p_minus_1 = Fr::zero() - Fr::one();                    // p - 1 as a Montgomery value: any canonical pattern works
        for (int it = 0; it < 4000; it++) {
            Fr a, b, w;
            for (int i = 0; i < 8; i++) { a.l[i] = (uint32_t)rng(); b.l[i] = (uint32_t)rng(); w.l[i] = (uint32_t)rng(); }
            a.l[7] &= 0x0fffffffu; b.l[7] &= 0x0fffffffu; w.l[7] &= 0x0fffffffu;
            if (it % 5 == 0) a = p_minus_1;
            if (it % 7 == 0) b = Fr::zero();
            if (it % 11 == 0) b = p_minus_1;
            // lift a and b into [p, 2p) half of the time: same residues, the other representative
            Fr al = a, bl = b;
            auto lift = [](Fr& x) {                                // x + p without reduction (fits: p < 2^253)
                uint64_t c = 0;
                for (int i = 0; i < 8; i++) { c += (uint64_t)x.l[i] + FrParams::mod(i); x.l[i] = (uint32_t)c; c >>= 32; }
            };
            if (it & 1) lift(al);
            if (it & 2) lift(bl);
            const Fr s = Fr::reduce_lazy(Fr::add_lazy(al, bl));
            if (!(s == a + b)) bad++;
            Fr d = Fr::sub_lazy(al, bl);                           // in (0, 4p)
            const Fr dm = Fr::reduce_lazy(Fr::mul_lazy(d, w));
            if (!(dm == (a - b) * w)) bad++;
            Fr::final_sub2(d);                                     // in [0, 2p)
            if (!(Fr::reduce_lazy(d) == a - b)) bad++;
            const Fr m2 = Fr::reduce_lazy(Fr::mul_lazy(al, w));
            if (!(m2 == a * w)) bad++;
        }
        printf("lazy-range mismatches: %d\n", bad - before);
    }
    printf("total mismatches: %d\n", bad);
    return bad != 0;
}
