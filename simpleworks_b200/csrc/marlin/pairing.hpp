// BLS12-377 pairing on the host, for the verifier only: what
// MarlinKZG10::check / batch_check evaluates under Marlin::verify (reference src/marlin/mod.rs:79-86).
// Restates the tower and group of ark-bls12-377 0.3 (fields/{fq2,fq6,fq12}.rs, curves/g2.rs):
//   Fq2  = Fq[u]  / (u^2 + 5)
//   Fq6  = Fq2[v] / (v^3 - u)
//   Fq12 = Fq6[w] / (w^2 - v)
//   G2   : y^2 = x^3 - u/5 over Fq2 (D-type sextic twist), cofactor derived in tools/gen_constants.py
// The pairing is the ate pairing f_{x,Q}(P)^((q^12-1)/r), x = 0x8508c00000000001.  Three Miller loops are kept:
//   miller_loop         T on the twist in homogeneous projective coordinates, sparse line products (the formulas of
//                       ark-ec 0.3 models/bls12/g2.rs) -- the one the verifier uses;
//   miller_loop_affine  T on the twist in affine Fq2 coordinates, one inversion per step;
//   miller_loop_plain   Q mapped to E(Fq12) through the untwisting map (x', y') -> (x' w^2, y' w^3), affine
//                       arithmetic over Fq12 with nothing curve-specific in it.
// The last two agree bit for bit, the first agrees with them after the final exponentiation (its lines carry Fq2
// factors); orc_pairing_selftest checks both, and bilinearity, on random points.  The final exponentiation uses the
// BLS12 decomposition (checked against the plain 2009-bit exponentiation).  Verification checks pairing *equations*,
// for which any bilinear non-degenerate pairing gives the same verdict as arkworks' own.  It stays on the host:
// 7.4 ms per proof on the GPU box (the two Miller loops share their squarings, the hard part of the final
// exponentiation squares cyclotomically; the rest is subgroup tests of the proof's points and the commitment MSM).
#pragma once
#include "curve_host.hpp"

namespace swb {
namespace marlin {

inline Fq fq_small(uint32_t v) {
    Fq r = Fq::zero();
    Fq one = Fq::one();
    for (int b = 31; b >= 0; b--) {
        r = r.dbl();
        if ((v >> b) & 1u) r = r + one;
    }
    return r;
}

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    static Fq mul5(const Fq& a) { Fq t = a.dbl().dbl(); return t + a; }
    Fq2 operator*(const Fq2& o) const {                     // u^2 = -5; Karatsuba: 3 Fq products
        const Fq a = c0 * o.c0, b = c1 * o.c1;
        return {a - mul5(b), (c0 + c1) * (o.c0 + o.c1) - a - b};
    }
    Fq2 conj() const { return {c0, c1.neg()}; }             // the q-power Frobenius (u^q = -u)
    Fq2 pow_words(const uint32_t* e, int nwords) const {
        Fq2 acc = one();
        for (int i = nwords * 32 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1u) acc = acc * (*this);
        }
        return acc;
    }
    Fq2 sqr() const { return (*this) * (*this); }
    Fq2 scale(const Fq& s) const { return {c0 * s, c1 * s}; }
    Fq norm() const { return c0.sqr() + mul5(c1.sqr()); }
    Fq2 inverse() const {
        Fq ni = norm().inverse_bingcd();
        return {c0 * ni, (c1 * ni).neg()};
    }
    Fq2 mul_by_u() const { return {mul5(c1).neg(), c0}; }   // (c0 + c1 u) u = -5 c1 + c0 u
};

struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 operator*(const Fq6& o) const {                     // v^3 = u; Karatsuba: 6 Fq2 products
        const Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
        const Fq2 r0 = t0 + ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_by_u();
        const Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_by_u();
        const Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    Fq6 scale2(const Fq2& s) const { return {c0 * s, c1 * s, c2 * s}; }
    Fq6 mul_by_v() const { return {c2.mul_by_u(), c0, c1}; }
    Fq6 inverse() const {
        Fq2 t0 = c0.sqr() - (c1 * c2).mul_by_u();
        Fq2 t1 = c2.sqr().mul_by_u() - c0 * c1;
        Fq2 t2 = c1.sqr() - c0 * c2;
        Fq2 d = c0 * t0 + (c2 * t1 + c1 * t2).mul_by_u();
        Fq2 di = d.inverse();
        return {t0 * di, t1 * di, t2 * di};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    static Fq12 zero() { return {Fq6::zero(), Fq6::zero()}; }
    static Fq12 from_fq(const Fq& a) { Fq12 r = zero(); r.c0.c0.c0 = a; return r; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    Fq12 operator+(const Fq12& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq12 operator-(const Fq12& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq12 operator*(const Fq12& o) const {                   // w^2 = v; Karatsuba: 3 Fq6 products
        const Fq6 a = c0 * o.c0, b = c1 * o.c1;
        return {a + b.mul_by_v(), (c0 + c1) * (o.c0 + o.c1) - a - b};
    }
    Fq12 conj() const { return {c0, c1.neg()}; }            // the q^6-power Frobenius
    Fq12 sqr() const {                                      // complex squaring: 2 Fq6 products
        const Fq6 ab = c0 * c1;
        const Fq6 t = (c0 + c1) * (c0 + c1.mul_by_v()) - ab - ab.mul_by_v();      // c0^2 + v c1^2
        return {t, ab + ab};
    }
    // squaring of an element of the cyclotomic subgroup (anything after the easy part of the final exponentiation):
    // Granger-Scott, "Faster squaring in the cyclotomic subgroup of sixth degree extensions" -- three Fq4 squarings
    // (6 Fq2 products) instead of 12; ark-ff's cyclotomic_square for this tower.  Wrong outside the subgroup; the
    // self-test compares it with sqr() on subgroup elements.
    Fq12 cyclotomic_sqr() const {
        const Fq2 &r0 = c0.c0, &r4 = c0.c1, &r3 = c0.c2, &r2 = c1.c0, &r1 = c1.c1, &r5 = c1.c2;
        auto fq4_sqr = [](const Fq2& a, const Fq2& b, Fq2* t0, Fq2* t1) {     // (a + b y)^2, y^2 = u
            const Fq2 ab = a * b;
            *t0 = (a + b) * (b.mul_by_u() + a) - ab - ab.mul_by_u();
            *t1 = ab + ab;
        };
        Fq2 t0, t1, t2, t3, t4, t5;
        fq4_sqr(r0, r1, &t0, &t1);
        fq4_sqr(r2, r3, &t2, &t3);
        fq4_sqr(r4, r5, &t4, &t5);
        auto three_minus_two = [](const Fq2& t, const Fq2& z) { Fq2 x = t - z; x = x + x; return x + t; };   // 3t - 2z
        auto three_plus_two = [](const Fq2& t, const Fq2& z) { Fq2 x = t + z; x = x + x; return x + t; };    // 3t + 2z
        Fq12 r;
        r.c0.c0 = three_minus_two(t0, r0);
        r.c1.c1 = three_plus_two(t1, r1);
        r.c1.c0 = three_plus_two(t5.mul_by_u(), r2);
        r.c0.c2 = three_minus_two(t4, r3);
        r.c0.c1 = three_minus_two(t2, r4);
        r.c1.c2 = three_plus_two(t3, r5);
        return r;
    }
    // times the sparse element a + (b + c v) w, a, b, c in Fq2 -- the shape of a line of the Miller loop
    // (arkworks' mul_by_034)
    Fq12 mul_by_line(const Fq2& a, const Fq2& b, const Fq2& c) const {
        auto mul_b_c = [](const Fq6& x, const Fq2& p, const Fq2& q) {             // x * (p + q v): 5 Fq2 products
            const Fq2 t0 = x.c0 * p, t1 = x.c1 * q;
            return Fq6{t0 + (x.c2 * q).mul_by_u(), (x.c0 + x.c1) * (p + q) - t0 - t1, x.c2 * p + t1};
        };
        const Fq6 t0 = c0.scale2(a);                        // c0 * a
        const Fq6 t1 = mul_b_c(c1, b, c);                   // c1 * (b + c v)
        const Fq6 mid = mul_b_c(c0 + c1, a + b, c);         // (c0 + c1) * (a + b + c v)
        return {t0 + t1.mul_by_v(), mid - t0 - t1};
    }
    Fq12 inverse() const {
        Fq6 d = (c0 * c0 - (c1 * c1).mul_by_v()).inverse();
        return {c0 * d, (c1 * d).neg()};
    }
    Fq12 pow_words(const uint32_t* e, int nwords) const {
        Fq12 acc = one();
        bool started = false;
        for (int i = nwords * 32 - 1; i >= 0; i--) {
            if (started) acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1u) { acc = acc * (*this); started = true; }
        }
        return acc;
    }
};

// ---- G2: affine points on the twist over Fq2 -------------------------------------------------
struct G2Point {
    Fq2 x, y;
    bool infinity = true;
    static G2Point identity() { return {Fq2::zero(), Fq2::zero(), true}; }
    bool operator==(const G2Point& o) const {
        if (infinity || o.infinity) return infinity == o.infinity;
        return x == o.x && y == o.y;
    }
};
inline Fq2 g2_coeff_b() { return {Fq::zero(), fq_small(5).inverse().neg()}; }      // -u/5
inline bool g2_on_curve(const G2Point& p) { return p.infinity || p.y.sqr() == p.x.sqr() * p.x + g2_coeff_b(); }
inline G2Point g2_neg(const G2Point& p) { return {p.x, p.y.neg(), p.infinity}; }
inline G2Point g2_add(const G2Point& a, const G2Point& b) {
    if (a.infinity) return b;
    if (b.infinity) return a;
    Fq2 lam;
    if (a.x == b.x) {
        if (!(a.y == b.y) || a.y.is_zero()) return G2Point::identity();
        Fq2 xx = a.x.sqr();
        lam = (xx + xx + xx) * (a.y + a.y).inverse();
    } else {
        lam = (b.y - a.y) * (b.x - a.x).inverse();
    }
    Fq2 x3 = lam.sqr() - a.x - b.x;
    return {x3, lam * (a.x - x3) - a.y, false};
}
// the plain ladder on affine points (one Fq2 inversion per step): reference for the projective one below
inline G2Point g2_mul_words_affine(const G2Point& p, const uint32_t* k, int nwords) {
    G2Point acc = G2Point::identity();
    for (int i = nwords * 32 - 1; i >= 0; i--) {
        acc = g2_add(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) acc = g2_add(acc, p);
    }
    return acc;
}
// [k]P with the accumulator in XYZZ coordinates over Fq2 (dbl-2008-s-1 / madd-2008-s, the formulas of g1.cuh with
// a = 0): no inversion until the end.  The cofactor clearing of G2Projective::rand is a 500-bit multiplication and
// the subgroup test of a deserialised point a 253-bit one: 22 ms -> 1.5 ms each.
inline G2Point g2_mul_words(const G2Point& p, const uint32_t* k, int nwords) {
    if (p.infinity) return G2Point::identity();
    Fq2 X = Fq2::zero(), Y = Fq2::zero(), ZZ = Fq2::zero(), ZZZ = Fq2::zero();    // ZZ == 0: the identity
    auto dbl = [&]() {
        if (ZZ.is_zero()) return;
        if (Y.is_zero()) { ZZ = Fq2::zero(); return; }                            // order-2 point
        const Fq2 U = Y + Y, V = U.sqr(), W = U * V, S = X * V, xx = X.sqr(), M = xx + xx + xx;
        const Fq2 X3 = M.sqr() - (S + S);
        Y = M * (S - X3) - W * Y;
        X = X3;
        ZZ = V * ZZ;
        ZZZ = W * ZZZ;
    };
    auto add_p = [&]() {
        if (ZZ.is_zero()) { X = p.x; Y = p.y; ZZ = Fq2::one(); ZZZ = Fq2::one(); return; }
        const Fq2 U2 = p.x * ZZ, S2 = p.y * ZZZ, P = U2 - X, R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) dbl(); else ZZ = Fq2::zero();
            return;
        }
        const Fq2 PP = P.sqr(), PPP = P * PP, Q = X * PP;
        const Fq2 X3 = R.sqr() - PPP - (Q + Q);
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    };
    for (int i = nwords * 32 - 1; i >= 0; i--) {
        dbl();
        if ((k[i >> 5] >> (i & 31)) & 1u) add_p();
    }
    if (ZZ.is_zero()) return G2Point::identity();
    const Fq2 inv = (ZZ * ZZZ).inverse();                 // x = X / ZZ, y = Y / ZZZ
    return {X * (ZZZ * inv), Y * (ZZ * inv), false};
}
inline G2Point g2_mul_fr(const G2Point& p, const Fr& k_mont) {
    Fr c = k_mont.to_canonical();
    return g2_mul_words(p, c.l, 8);
}
// square root in Fq2 (complex method over the norm); false for non-residues.  Every candidate is
// verified by squaring, so a returned root is always correct.
inline bool fq2_sqrt(const Fq2& a, Fq2* out) {
    if (a.is_zero()) { *out = a; return true; }
    const FqSqrtCtx& sq = fq_sqrt_ctx();
    const Fq two_inv = fq_small(2).inverse();
    if (a.c1.is_zero()) {
        // a in Fq: either sqrt(a) in Fq, or sqrt(-a/5) * u
        Fq r;
        if (sq.sqrt(a.c0, &r)) { *out = {r, Fq::zero()}; return true; }
        if (sq.sqrt(a.c0.neg() * fq_small(5).inverse(), &r)) { *out = {Fq::zero(), r}; return out->sqr() == a; }
        return false;
    }
    // a = c0 + c1 u, u^2 = -5:  x0^2 = (c0 +- alpha)/2 with alpha = sqrt(norm(a)),  x1 = c1 / (2 x0)
    Fq alpha;
    if (!sq.sqrt(a.norm(), &alpha)) return false;
    for (int sign = 0; sign < 2; sign++) {
        Fq delta = (sign ? a.c0 - alpha : a.c0 + alpha) * two_inv, x0;
        if (!sq.sqrt(delta, &x0) || x0.is_zero()) continue;
        Fq2 cand = {x0, a.c1 * x0.dbl().inverse()};
        if (cand.sqr() == a) { *out = cand; return true; }
    }
    return false;
}
// lexicographic "greater" on (c1, c0) as canonical integers, the order ark-ff uses for Fp2
inline bool fq2_gt(const Fq2& a, const Fq2& b) {
    if (!(a.c1 == b.c1)) return fq_canonical_gt(a.c1, b.c1);
    return fq_canonical_gt(a.c0, b.c0);
}
// G2Projective::rand: x <- Fq2::rand (c0 then c1), greatest <- bool, until on-curve; cofactor
inline G2Point g2_rand(ChaChaRng& rng) {
    static const uint32_t cof[SWB_G2_COFACTOR_WORDS] = SWB_G2_COFACTOR_INIT;
    for (;;) {
        Fq2 x = {rand_fq(rng), Fq::zero()};
        x.c1 = rand_fq(rng);
        const bool greatest = rng.next_bool();
        Fq2 y;
        if (!fq2_sqrt(x.sqr() * x + g2_coeff_b(), &y)) continue;
        Fq2 negy = y.neg();
        const bool y_lt_negy = fq2_gt(negy, y);
        G2Point p = {x, (y_lt_negy ^ greatest) ? y : negy, false};
        return g2_mul_words(p, cof, SWB_G2_COFACTOR_WORDS);
    }
}

// CanonicalSerialize for G2Affine, compressed: x = c0 (48 B) || c1 (48 B), SWFlags in the top bits of
// the last byte (bit 7: y is the larger of {y, -y}; bit 6: infinity)
inline void put_g2_compressed(std::vector<uint8_t>& out, const G2Point& p) {
    const size_t at = out.size();
    put_fq_canonical(out, p.infinity ? Fq::zero() : p.x.c0);
    put_fq_canonical(out, p.infinity ? Fq::zero() : p.x.c1);
    if (p.infinity) out[at + 95] |= 0x40;
    else if (fq2_gt(p.y, p.y.neg())) out[at + 95] |= 0x80;
}
inline bool get_g2_compressed(const uint8_t*& p, const uint8_t* end, G2Point* out) {
    if (end - p < 96) return false;
    uint8_t buf[96];
    memcpy(buf, p, 96);
    const uint8_t flags = buf[95] & 0xC0;
    buf[95] &= 0x3F;
    p += 96;
    if (flags & 0x40) {
        if (flags & 0x80) return false;
        for (int i = 0; i < 96; i++)
            if (buf[i]) return false;
        *out = G2Point::identity();
        return true;
    }
    Fq c[2];
    for (int k = 0; k < 2; k++)
        if (!fq_from_canonical_bytes(buf + 48 * k, &c[k])) return false;      // refuses coordinates >= q
    Fq2 x = {c[0], c[1]}, y;
    if (!fq2_sqrt(x.sqr() * x + g2_coeff_b(), &y)) return false;
    const Fq2 negy = y.neg();
    const bool want_larger = (flags & 0x80) != 0;
    *out = {x, (fq2_gt(y, negy) == want_larger) ? y : negy, false};
    // prime-order subgroup: [r]Q == O (a small-order point in a verifying key would otherwise reach the Miller loop)
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = FrParams::mod(i);
    return g2_mul_words(*out, r, 8).infinity;
}

// ---- ate pairing --------------------------------------------------------------------------------
struct E12Point {
    Fq12 x, y;
    bool infinity;
};
inline E12Point untwist(const G2Point& q) {
    E12Point r;
    r.infinity = q.infinity;
    r.x = Fq12::zero();
    r.y = Fq12::zero();
    r.x.c0.c1 = q.x;        // x' * v   (w^2 = v)
    r.y.c1.c1 = q.y;        // y' * v w (w^3 = v w)
    return r;
}
// Miller function f_{x,Q}(P) with T on the twist in homogeneous projective coordinates (Costello-Lange-Naehrig, the
// formulas of ark-ec 0.3 models/bls12/g2.rs for a D-type twist): no inversion at all, lines scaled by Fq2 factors
// that the final exponentiation removes.  A doubling step costs 2 products + 7 squarings in Fq2, an addition step
// 11 products + 2 squarings; the line a + (b + c v) w = (-2YZ py) + (3X^2 px) w + (3b'Z^2 - Y^2) v w multiplies f
// sparsely.  Checked against miller_loop_affine / miller_loop_plain after the final exponentiation
// (orc_pairing_selftest).
// prod_k f_{x,Q_k}(P_k): the squarings of f are shared by all pairs (a batch check has two)
inline Fq12 miller_loop_multi(const std::vector<std::pair<G1Point, G2Point>>& pairs) {
    static const Fq two_inv = fq_small(2).inverse();
    static const Fq2 bt = g2_coeff_b();
    struct T { Fq2 X, Y, Z; const G1Point* p; const G2Point* q; };
    std::vector<T> ts;
    for (auto& pq : pairs)
        if (!pq.first.infinity && !pq.second.infinity) ts.push_back({pq.second.x, pq.second.y, Fq2::one(), &pq.first, &pq.second});
    Fq12 f = Fq12::one();
    if (ts.empty()) return f;
    const uint64_t x = SWB_BLS_X;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        f = f.sqr();
        for (auto& t : ts) {   // doubling step
            Fq2 &X = t.X, &Y = t.Y, &Z = t.Z;
            const Fq2 a = (X * Y).scale(two_inv), b = Y.sqr(), c = Z.sqr();
            const Fq2 e = bt * (c + c + c), f3 = e + e + e;
            const Fq2 g = (b + f3).scale(two_inv), h = (Y + Z).sqr() - (b + c), i2 = e - b, j = X.sqr(), e2 = e.sqr();
            X = a * (b - f3);
            Y = g.sqr() - (e2 + e2 + e2);
            Z = b * h;
            f = f.mul_by_line(h.neg().scale(t.p->y), (j + j + j).scale(t.p->x), i2);
        }
        if ((x >> i) & 1) {
            for (auto& t : ts) {   // addition step: T <- T + Q
                Fq2 &X = t.X, &Y = t.Y, &Z = t.Z;
                const G2Point& q = *t.q;
                const Fq2 theta = Y - q.y * Z, lambda = X - q.x * Z;
                const Fq2 c = theta.sqr(), d = lambda.sqr(), e = lambda * d, ff = Z * c, g = X * d;
                const Fq2 h = e + ff - (g + g);
                X = lambda * h;
                Y = theta * (g - h) - e * Y;
                Z = Z * e;
                const Fq2 j = theta * q.x - lambda * q.y;
                f = f.mul_by_line(lambda.scale(t.p->y), theta.neg().scale(t.p->x), j);
            }
        }
    }
    return f;
}
inline Fq12 miller_loop(const G1Point& p, const G2Point& q) { return miller_loop_multi({{p, q}}); }
// The same Miller function with T kept
// on the TWIST in affine Fq2 coordinates.  The untwisted point is (x' w^2, y' w^3), so a slope on E is lambda' w
// with lambda' = 3 x'^2 / (2 y') (or the chord's) in Fq2, the new point is (lambda'^2 - x1' - x2',
// lambda' (x1' - x3') - y1') again on the twist, and the line through T evaluated at P = (px, py) is
//     py - lambda' px * w + (lambda' x' - y') * w^3,
// i.e. the Fq12 element c0 = (py, 0, 0), c1 = (-lambda' px, lambda' x' - y', 0) -- the same VALUES as the loop
// over E(Fq12) below produces (miller_loop_plain, kept as the reference; tests compare the two bit for bit), at
// one Fq2 inversion and a handful of Fq2 products per step instead of an Fq12 inversion and five Fq12 products.
inline Fq12 miller_loop_affine(const G1Point& p, const G2Point& q) {
    if (p.infinity || q.infinity) return Fq12::one();
    Fq2 tx = q.x, ty = q.y;
    Fq12 f = Fq12::one();
    const uint64_t x = SWB_BLS_X;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    auto line_and_step = [&](const Fq2& lam, const Fq2& ax) {
        Fq12 l = Fq12::zero();
        l.c0.c0.c0 = p.y;
        l.c1.c0 = lam.scale(p.x).neg();
        l.c1.c1 = lam * tx - ty;
        const Fq2 nx = lam.sqr() - tx - ax;
        ty = lam * (tx - nx) - ty;
        tx = nx;
        return l;
    };
    for (int i = top - 1; i >= 0; i--) {
        const Fq2 xx = tx.sqr();
        const Fq2 lam = (xx + xx + xx) * (ty + ty).inverse();
        f = f.sqr() * line_and_step(lam, tx);
        if ((x >> i) & 1) {
            const Fq2 lam2 = (q.y - ty) * (q.x - tx).inverse();
            f = f * line_and_step(lam2, q.x);
        }
    }
    return f;
}
// the same function computed over E(Fq12) with nothing curve-specific in it
inline Fq12 miller_loop_plain(const G1Point& p, const G2Point& q) {
    if (p.infinity || q.infinity) return Fq12::one();
    const E12Point Q = untwist(q);
    const Fq12 px = Fq12::from_fq(p.x), py = Fq12::from_fq(p.y);
    Fq12 tx = Q.x, ty = Q.y;
    Fq12 f = Fq12::one();
    const uint64_t x = SWB_BLS_X;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    auto line_and_step = [&](const Fq12& lam, const Fq12& ax, const Fq12& ay) {
        // line through (tx,ty) with slope lam evaluated at P, then T <- T + A
        Fq12 l = (py - ty) - lam * (px - tx);
        Fq12 nx = lam.sqr() - tx - ax;
        Fq12 ny = lam * (tx - nx) - ty;
        tx = nx;
        ty = ny;
        (void)ay;
        return l;
    };
    for (int i = top - 1; i >= 0; i--) {
        Fq12 xx = tx.sqr();
        Fq12 lam = (xx + xx + xx) * (ty + ty).inverse();
        f = f.sqr() * line_and_step(lam, tx, ty);
        if ((x >> i) & 1) {
            Fq12 lam2 = (Q.y - ty) * (Q.x - tx).inverse();
            f = f * line_and_step(lam2, Q.x, Q.y);
        }
    }
    return f;
}
// ---- Frobenius on the tower --------------------------------------------------------------------
// x -> x^q.  On Fq2 it is conjugation; v^q = v * u^((q-1)/3) and w^q = w * u^((q-1)/6), so on Fq6 / Fq12
// it conjugates the Fq2 coefficients and scales them by powers of those two constants, which are
// computed once from the modulus.
struct FrobeniusCtx {
    Fq2 g1, g2, d;          // u^((q-1)/3), its square, u^((q-1)/6)
    FrobeniusCtx() {
        uint32_t qm1[12], e3[12], e6[12];
        for (int i = 0; i < 12; i++) qm1[i] = FqParams::mod(i);
        qm1[0] -= 1;
        div_small(e3, qm1, 3);
        div_small(e6, qm1, 6);
        const Fq2 u = {Fq::zero(), Fq::one()};
        g1 = u.pow_words(e3, 12);
        g2 = g1 * g1;
        d = u.pow_words(e6, 12);
    }
    static void div_small(uint32_t* out, const uint32_t* in, uint32_t k) {       // exact division of 12 limbs
        uint64_t rem = 0;
        for (int i = 11; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | in[i];
            out[i] = (uint32_t)(cur / k);
            rem = cur % k;
        }
    }
};
inline const FrobeniusCtx& frobenius_ctx() {
    static const FrobeniusCtx c;
    return c;
}
inline Fq6 frobenius6(const Fq6& a) {
    const FrobeniusCtx& c = frobenius_ctx();
    return {a.c0.conj(), a.c1.conj() * c.g1, a.c2.conj() * c.g2};
}
inline Fq12 frobenius(const Fq12& a, int power = 1) {
    Fq12 r = a;
    for (int k = 0; k < power; k++) r = {frobenius6(r.c0), frobenius6(r.c1).scale2(frobenius_ctx().d)};
    return r;
}

// f^((q^12 - 1)/r), the plain way: (conj(f) / f)^((q^6 + 1)/r) -- conjugation over Fq6 is the q^6-power
// Frobenius.  3000 Fq12 products; kept as the reference the fast version is checked against.
inline Fq12 final_exponentiation_plain(const Fq12& f) {
    static const uint32_t e[SWB_FINAL_EXP_WORDS] = SWB_FINAL_EXP_INIT;
    Fq12 g = f.conj() * f.inverse();
    return g.pow_words(e, SWB_FINAL_EXP_WORDS);
}
// f^x, x = 0x8508c00000000001 (positive); cyclotomic: f lies in the cyclotomic subgroup (cheaper squarings)
inline Fq12 exp_by_x(const Fq12& f, bool cyclotomic = false) {
    const uint64_t x = SWB_BLS_X;
    Fq12 acc = f;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        acc = cyclotomic ? acc.cyclotomic_sqr() : acc.sqr();
        if ((x >> i) & 1) acc = acc * f;
    }
    return acc;
}
// The BLS12 decomposition (easy part by Frobenius, hard part as in Fuentes-Castaneda, Knapp, Rodriguez-
// Henriquez, "Faster hashing to G2", the schedule ark-ec 0.3 bls12/mod.rs uses): five exponentiations by
// the 64-bit x instead of one by a 2009-bit number.  It returns the CUBE of final_exponentiation_plain
// (the hard part is raised to 3 (q^4 - q^2 + 1) / r); 3 is prime to r, so "== 1" tests and equalities
// between values computed this way are unaffected.  orc_pairing_selftest checks exactly that relation.
inline Fq12 final_exponentiation(const Fq12& f) {
    Fq12 r = f.conj() * f.inverse();                        // f^(q^6 - 1)
    r = frobenius(r, 2) * r;                                // ^(q^2 + 1): now in the cyclotomic subgroup, inverse = conj
    Fq12 y0 = r.cyclotomic_sqr().conj();
    Fq12 y5 = exp_by_x(r, true);
    Fq12 y1 = y5.cyclotomic_sqr();
    Fq12 y3 = y0 * y5;
    y0 = exp_by_x(y3, true);
    const Fq12 y2 = exp_by_x(y0, true);
    Fq12 y4 = exp_by_x(y2, true);
    y4 = y4 * y1;
    y1 = exp_by_x(y4, true);
    y3 = y3.conj();
    y1 = y1 * y3;
    y1 = y1 * r;
    y3 = r.conj();
    y0 = y0 * r;
    y0 = frobenius(y0, 3);
    y4 = y4 * y3;
    y4 = frobenius(y4, 1);
    y5 = y5 * y2;
    y5 = frobenius(y5, 2);
    y5 = y5 * y0;
    y5 = y5 * y4;
    y5 = y5 * y1;
    return y5;
}
inline Fq12 pairing(const G1Point& p, const G2Point& q) { return final_exponentiation(miller_loop(p, q)); }
// prod_i e(P_i, Q_i) == 1 ?
inline bool pairing_product_is_one(const std::vector<std::pair<G1Point, G2Point>>& pairs) {
    return final_exponentiation(miller_loop_multi(pairs)) == Fq12::one();
}

}  // namespace marlin
}  // namespace swb
