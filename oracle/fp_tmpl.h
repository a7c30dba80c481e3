/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into the product library.
 *
 * Prime-field template restating ark-ff 0.3 `Fp256<P>` / `Fp384<P>` (ark-ff/src/fields/macros.rs,
 * arithmetic.rs; crate not vendored in /root/reference -- reference Cargo.toml:15 pins ^0.3.0).
 * Elements are N little-endian u64 limbs holding a*R mod p (R = 2^(64N)), always fully reduced,
 * exactly the in-memory representation of ark_ff::FpXXX.  PARITY UNPINNED vs real arkworks (no
 * golden vectors exist in the reference); pinned against oracle/golden.py big-int arithmetic by
 * tests/test_oracle_*.py and tests/golden/.
 *
 * Instantiate with:  #define FP  name   #define FPN limbs   then include.
 * Provides  FP_mul/sqr/add/sub/neg/dbl/inv/pow/from_u64/to_canon/from_canon/is_zero/eq/cmp.
 */
#include <stdint.h>
#include <string.h>

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define F(name) CAT(FP, name)
#define FT CAT(FP, t)

typedef unsigned __int128 u128;

extern const uint64_t F(MOD)[FPN];
extern const uint64_t F(R1)[FPN];   /* R mod p  (Montgomery one) */
extern const uint64_t F(R2)[FPN];   /* R^2 mod p */
extern const uint64_t F(INV);       /* -p^-1 mod 2^64 */

static inline int F(is_zero)(const FT *a) {
    uint64_t o = 0;
    for (int i = 0; i < FPN; i++) o |= a->l[i];
    return o == 0;
}
static inline int F(eq)(const FT *a, const FT *b) { return memcmp(a, b, sizeof(FT)) == 0; }
static inline int F(cmp_raw)(const uint64_t *a, const uint64_t *b) {
    for (int i = FPN - 1; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static inline uint64_t F(sub_raw)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < FPN; i++) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
    return borrow;
}
static inline uint64_t F(add_raw)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t carry = 0;
    for (int i = 0; i < FPN; i++) {
        u128 t = (u128)a[i] + b[i] + carry;
        r[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    return carry;
}
/* ark-ff add_assign: add then subtract modulus if >= p (moduli here have spare top bits) */
static inline void F(add)(FT *r, const FT *a, const FT *b) {
    F(add_raw)(r->l, a->l, b->l);
    if (F(cmp_raw)(r->l, F(MOD)) >= 0) F(sub_raw)(r->l, r->l, F(MOD));
}
static inline void F(dbl)(FT *r, const FT *a) { F(add)(r, a, a); }
/* ark-ff sub_assign: if b > a add modulus first */
static inline void F(sub)(FT *r, const FT *a, const FT *b) {
    if (F(sub_raw)(r->l, a->l, b->l)) F(add_raw)(r->l, r->l, F(MOD));
}
static inline void F(neg)(FT *r, const FT *a) {
    if (F(is_zero)(a)) { *r = *a; return; }
    F(sub_raw)(r->l, F(MOD), a->l);
}
/* ark-ff mul_assign: CIOS Montgomery multiplication, one interleaved reduction per outer limb;
 * output in [0, p). */
static inline void F(mul)(FT *out, const FT *a, const FT *b) {
    uint64_t r[FPN + 2];
    memset(r, 0, sizeof r);
    for (int i = 0; i < FPN; i++) {
        u128 c = 0;
        for (int j = 0; j < FPN; j++) {
            c += (u128)a->l[j] * b->l[i] + r[j];
            r[j] = (uint64_t)c;
            c >>= 64;
        }
        c += r[FPN];
        r[FPN] = (uint64_t)c;
        r[FPN + 1] = (uint64_t)(c >> 64);
        uint64_t k = r[0] * F(INV);
        c = (u128)k * F(MOD)[0] + r[0];
        c >>= 64;
        for (int j = 1; j < FPN; j++) {
            c += (u128)k * F(MOD)[j] + r[j];
            r[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += r[FPN];
        r[FPN - 1] = (uint64_t)c;
        r[FPN] = r[FPN + 1] + (uint64_t)(c >> 64);
    }
    if (r[FPN] || F(cmp_raw)(r, F(MOD)) >= 0) F(sub_raw)(r, r, F(MOD));
    memcpy(out->l, r, sizeof(out->l));
}
static inline void F(sqr)(FT *r, const FT *a) { F(mul)(r, a, a); }
static inline void F(one)(FT *r) { memcpy(r->l, F(R1), sizeof r->l); }
static inline void F(zero)(FT *r) { memset(r->l, 0, sizeof r->l); }
/* into_repr(): one Montgomery reduction = multiply by the integer 1 */
static inline void F(to_canon)(uint64_t *out, const FT *a) {
    FT one_raw, t;
    memset(&one_raw, 0, sizeof one_raw);
    one_raw.l[0] = 1;
    F(mul)(&t, a, &one_raw);
    memcpy(out, t.l, sizeof t.l);
}
/* from_repr(): multiply by R^2 */
static inline void F(from_canon)(FT *r, const uint64_t *in) {
    FT t, r2;
    memcpy(t.l, in, sizeof t.l);
    memcpy(r2.l, F(R2), sizeof r2.l);
    F(mul)(r, &t, &r2);
}
static inline void F(from_u64)(FT *r, uint64_t v) {
    uint64_t t[FPN];
    memset(t, 0, sizeof t);
    t[0] = v;
    F(from_canon)(r, t);
}
/* a^e, e given as nl little-endian u64 limbs (square-and-multiply, MSB first) */
static inline void F(pow)(FT *r, const FT *a, const uint64_t *e, int nl) {
    FT acc;
    F(one)(&acc);
    int started = 0;
    for (int i = nl * 64 - 1; i >= 0; i--) {
        if (started) F(sqr)(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            F(mul)(&acc, &acc, a);
            started = 1;
        }
    }
    *r = acc;
}
/* field inverse; arkworks uses a binary EEA, the value is unique so Fermat is equivalent.
 * inverse of zero is reported as zero (arkworks returns None; callers here check first). */
static inline void F(inv)(FT *r, const FT *a) {
    uint64_t e[FPN];
    uint64_t two[FPN];
    memset(two, 0, sizeof two);
    two[0] = 2;
    F(sub_raw)(e, F(MOD), two);
    F(pow)(r, a, e, FPN);
}

#undef FT
#undef F
#undef CAT
#undef CAT_
