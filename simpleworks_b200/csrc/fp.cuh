// BLS12-377 prime fields on 32-bit limbs: Fr (8 limbs) and Fq (12 limbs), Montgomery form.
//
// Replaces ark_ff::Fp256<FrParameters> / Fp384<FqParameters> arithmetic (ark-ff 0.3
// fields/macros.rs; reached from reference src/marlin/mod.rs:52,75,92 through ark-marlin).
// The in-memory value is bit-identical to arkworks': a*R mod p, R = 2^(32N) = 2^256 / 2^384,
// little-endian limbs, always fully reduced into [0, p).
//
// Multiplication is an operand-scanning Montgomery product on two interleaved accumulators
// ("even" and "odd" columns).  Each 32x32 product is a (mad.lo.cc, madc.hi.cc) pair that ptxas
// fuses into one IMAD.WIDE.U32.X, and splitting even/odd columns keeps every carry chain a
// straight line of IMAD.WIDE instructions: 2*N^2 of them per product (128 for Fr, 288 for Fq),
// the "limb-product" unit used by the roofline in DESIGN.md.  -p^-1 mod 2^32 = 0xFFFFFFFF for
// both moduli, so the per-row reduction factor is just m = -acc[0].
#pragma once
#include "constants.cuh"
#include "ptx_ops.cuh"

namespace swb {

#if defined(__CUDACC__)
// -p^-1 mod 2^32 kept in constant memory on purpose.  As a literal, ptxas rewrites m = t0 * (-1)
// into a negation and folds the sign into the immediates of the following multiply-accumulates,
// after which the low and high halves no longer share operands and every IMAD.WIDE of the
// reduction chain is split into IMAD + IMAD.HI (measured: 422 vs 300 pipe instructions per Fq
// product).  An opaque multiplier keeps the chains as single IMAD.WIDE.U32.X instructions.
static __constant__ uint32_t c_mont_inv32 = 0xFFFFFFFFu;
#endif

struct FrParams {
    static constexpr int N = SWB_FR_LIMBS;
    static SWB_HD constexpr uint32_t mod(int i) { constexpr uint32_t v[N] = SWB_FR_MOD_INIT; return v[i]; }
    static SWB_HD constexpr uint32_t r1(int i) { constexpr uint32_t v[N] = SWB_FR_R1_INIT; return v[i]; }
    static SWB_HD constexpr uint32_t r2(int i) { constexpr uint32_t v[N] = SWB_FR_R2_INIT; return v[i]; }
    static constexpr uint32_t INV = SWB_FR_INV32;
    static constexpr uint64_t INV64 = SWB_FR_INV64;
};
struct FqParams {
    static constexpr int N = SWB_FQ_LIMBS;
    static SWB_HD constexpr uint32_t mod(int i) { constexpr uint32_t v[N] = SWB_FQ_MOD_INIT; return v[i]; }
    static SWB_HD constexpr uint32_t r1(int i) { constexpr uint32_t v[N] = SWB_FQ_R1_INIT; return v[i]; }
    static SWB_HD constexpr uint32_t r2(int i) { constexpr uint32_t v[N] = SWB_FQ_R2_INIT; return v[i]; }
    static constexpr uint32_t INV = SWB_FQ_INV32;
    static constexpr uint64_t INV64 = SWB_FQ_INV64;
};

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    // ---- constructors -------------------------------------------------------------------
    static SWB_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    static SWB_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r1(i);
        return r;
    }
    static SWB_HD Fp r_squared() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    // R^3 mod p as a raw limb pattern: mont_mul(R^2, R^2) = R^4 R^-1 = R^3
    static SWB_HD Fp r_cubed() { return r_squared() * r_squared(); }
    SWB_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    SWB_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    SWB_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // ---- r = (t >= p) ? t - p : t ------------------------------------------------------
    static SWB_HD void final_sub(Fp& t) {
        uint32_t s[N];
        s[0] = ptx::sub_cc(t.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) s[i] = ptx::subc_cc(t.l[i], P::mod(i));
        uint32_t borrow = ptx::subc(0u, 0u);   // 0xFFFFFFFF when t < p
#pragma unroll
        for (int i = 0; i < N; i++) t.l[i] = borrow ? t.l[i] : s[i];
    }

    // ---- lazy range [0, 2p) (Harvey butterflies; the moduli leave >= 3 spare top bits, so 4p fits) ----------
    // limb i of 2p
    static SWB_HD constexpr uint32_t mod2(int i) { return (P::mod(i) << 1) | (i ? (P::mod(i - 1) >> 31) : 0u); }
    // t < 4p -> t mod-ish 2p: (t >= 2p) ? t - 2p : t
    static SWB_HD void final_sub2(Fp& t) {
        uint32_t s[N];
        s[0] = ptx::sub_cc(t.l[0], mod2(0));
#pragma unroll
        for (int i = 1; i < N; i++) s[i] = ptx::subc_cc(t.l[i], mod2(i));
        uint32_t borrow = ptx::subc(0u, 0u);
#pragma unroll
        for (int i = 0; i < N; i++) t.l[i] = borrow ? t.l[i] : s[i];
    }
    // a, b in [0, 2p) -> a + b in [0, 2p)
    static SWB_HD Fp add_lazy(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);
        final_sub2(r);
        return r;
    }
    // a, b in [0, 2p) -> a - b + 2p in (0, 4p): no comparison at all; feed it to mul_lazy or final_sub2
    static SWB_HD Fp sub_lazy(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = ptx::add_cc(a.l[0], mod2(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], mod2(i));
        r.l[N - 1] = ptx::addc(a.l[N - 1], mod2(N - 1));
        r.l[0] = ptx::sub_cc(r.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(r.l[i], b.l[i]);
        r.l[N - 1] = ptx::subc(r.l[N - 1], b.l[N - 1]);
        return r;
    }
    // t in [0, 2p) -> canonical
    static SWB_HD Fp reduce_lazy(const Fp& t) {
        Fp r = t;
        final_sub(r);
        return r;
    }

    friend SWB_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);   // moduli leave >= 3 spare top bits
        final_sub(r);
        return r;
    }
    friend SWB_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        uint32_t mask = ptx::subc(0u, 0u);     // all ones when a < b
        r.l[0] = ptx::add_cc(r.l[0], P::mod(0) & mask);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], P::mod(i) & mask);
        r.l[N - 1] = ptx::addc(r.l[N - 1], P::mod(N - 1) & mask);
        return r;
    }
    SWB_HD Fp neg() const { return zero() - *this; }
    SWB_HD Fp dbl() const { return *this + *this; }

    // ---- Montgomery product -----------------------------------------------------------
    // acc pairs (x[j], x[j+1]) += a[j+off] * bi  for j = 0,2,..,N-2 as one carry chain; the
    // carry out of the top pair is left in CC.
    template <bool CARRY_IN>
    static SWB_HD void chain_mad(uint32_t* x, const uint32_t* a, uint32_t bi) {
        x[0] = CARRY_IN ? ptx::madc_lo_cc(a[0], bi, x[0]) : ptx::mad_lo_cc(a[0], bi, x[0]);
        x[1] = ptx::madc_hi_cc(a[0], bi, x[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            x[j] = ptx::madc_lo_cc(a[j], bi, x[j]);
            x[j + 1] = ptx::madc_hi_cc(a[j], bi, x[j + 1]);
        }
    }
    // same with the modulus as (compile-time) multiplicand, starting at limb `off`
    template <int OFF>
    static SWB_HD void chain_mad_mod(uint32_t* x, uint32_t m) {
        if (OFF == 0 && P::mod(0) == 1u) {
            // both moduli are 1 mod 2^32: x[0] + m*1 is 0 by construction, only its carry
            // matters.  (Spelling it as a multiplication by the literal 1 makes ptxas strength-
            // reduce the pair and then split every IMAD.WIDE of the chain into IMAD + IMAD.HI.)
            (void)ptx::add_cc(x[0], m);
            x[1] = ptx::addc_cc(x[1], 0u);
        } else {
            x[0] = ptx::mad_lo_cc(P::mod(OFF), m, x[0]);
            x[1] = ptx::madc_hi_cc(P::mod(OFF), m, x[1]);
        }
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            x[j] = ptx::madc_lo_cc(P::mod(OFF + j), m, x[j]);
            x[j + 1] = ptx::madc_hi_cc(P::mod(OFF + j), m, x[j + 1]);
        }
    }
    // add m*p with m chosen so the lowest limb of (ev + 2^32*od) becomes zero
    static SWB_HD void reduce_row(uint32_t* ev, uint32_t* od) {
#if defined(__CUDA_ARCH__)
        static_assert(P::INV == 0xFFFFFFFFu, "c_mont_inv32 holds the shared -p^-1 mod 2^32");
        uint32_t m = ev[0] * c_mont_inv32;
#else
        uint32_t m = ptx::mul_lo(ev[0], P::INV);
#endif
        chain_mad_mod<1>(od, m);               // odd limbs of p -> odd accumulator (no carry out)
        chain_mad_mod<0>(ev, m);               // even limbs of p -> even accumulator
        od[N - 1] = ptx::addc(od[N - 1], 0u);  // its carry has the weight of od's top limb
    }
    // One row for i >= 1.  On entry `x` is the even accumulator with x[0] == 0 and `y` the odd
    // one; dividing by 2^32 swaps their roles: y becomes even, x (shifted down two limbs)
    // becomes odd.  Adds a*bi, then reduces.
    static SWB_HD void row(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi) {
        y[0] = ptx::add_cc(y[0], x[1]);        // x[1] has the weight of y[0]; its carry feeds x's chain
#pragma unroll
        for (int j = 1; j < N - 1; j += 2) {
            x[j - 1] = ptx::madc_lo_cc(a[j], bi, x[j + 1]);
            x[j] = ptx::madc_hi_cc(a[j], bi, x[j + 2]);
        }
        x[N - 2] = ptx::madc_lo_cc(a[N - 1], bi, 0u);
        x[N - 1] = ptx::madc_hi(a[N - 1], bi, 0u);
        chain_mad<false>(y, a, bi);            // even limbs of a
        x[N - 1] = ptx::addc(x[N - 1], 0u);
        reduce_row(y, x);
    }

    // the device schedule (also runs on the host through the emulated carry flag, for tests)
    static SWB_HD Fp mul_limb_schedule(const Fp& a, const Fp& b) { return mul_limb_schedule_t<true>(a, b); }
    // Montgomery product without the final conditional subtraction: for a < 4p and b < p (or the other way round) the
    // running value stays below a + p < 2^(32 N) and the result (a b + m p) / R is below 2p
    static SWB_HD Fp mul_lazy(const Fp& a, const Fp& b) { return mul_limb_schedule_t<false>(a, b); }
    template <bool REDUCE>
    static SWB_HD Fp mul_limb_schedule_t(const Fp& a, const Fp& b) {
        uint32_t ev[N], od[N];
        // row 0: nothing to accumulate onto
        {
            const uint32_t bi = b.l[0];
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                ev[j] = ptx::mul_lo(a.l[j], bi);
                ev[j + 1] = ptx::mul_hi(a.l[j], bi);
                od[j] = ptx::mul_lo(a.l[j + 1], bi);
                od[j + 1] = ptx::mul_hi(a.l[j + 1], bi);
            }
            reduce_row(ev, od);
        }
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row(ev, od, a.l, b.l[i]);          // afterwards od is even, ev is odd
            if (i + 1 < N) row(od, ev, a.l, b.l[i + 1]);
        }
        // N is even: after N-1 further rows `od` is the even accumulator (od[0] == 0), `ev` odd.
        Fp r;
        r.l[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(ev[k], od[k + 1]);
        r.l[N - 1] = ptx::addc(ev[N - 1], 0u);
        if (REDUCE) final_sub(r);
        return r;
    }
#if !defined(__CUDA_ARCH__)
    // host path: the same Montgomery product on 64-bit limbs (CIOS), used by the small amount of
    // host-side curve arithmetic at the end of an MSM and by the host protocol code.
    static inline Fp mul_host64(const Fp& a, const Fp& b) {
        constexpr int M = N / 2;
        typedef unsigned __int128 u128;
        uint64_t x[M], y[M], p[M], t[M + 2];
        for (int i = 0; i < M; i++) {
            x[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
            y[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
            p[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
        }
        for (int i = 0; i < M + 2; i++) t[i] = 0;
        const uint64_t inv = P::INV64;
        for (int i = 0; i < M; i++) {
            u128 c = 0;
            for (int j = 0; j < M; j++) {
                c += (u128)x[j] * y[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M] = (uint64_t)c;
            t[M + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * inv;
            c = ((u128)m * p[0] + t[0]) >> 64;
            for (int j = 1; j < M; j++) {
                c += (u128)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M - 1] = (uint64_t)c;
            t[M] = t[M + 1] + (uint64_t)(c >> 64);
        }
        bool ge = t[M] != 0;
        if (!ge) {
            ge = true;
            for (int i = M - 1; i >= 0; i--) {
                if (t[i] != p[i]) { ge = t[i] > p[i]; break; }
            }
        }
        if (ge) {
            uint64_t br = 0;
            for (int i = 0; i < M; i++) {
                u128 d = (u128)t[i] - p[i] - br;
                t[i] = (uint64_t)d;
                br = (uint64_t)(d >> 64) & 1;
            }
        }
        Fp r;
        for (int i = 0; i < M; i++) {
            r.l[2 * i] = (uint32_t)t[i];
            r.l[2 * i + 1] = (uint32_t)(t[i] >> 32);
        }
        return r;
    }
#endif
#if defined(__CUDACC__)
    // out-of-line copy for translation units that define SWB_FP_NOINLINE_MUL: cold kernels (bucket
    // reduction, table building) call the product instead of inlining ~300 instructions per use
    static __device__ __noinline__ Fp mul_outlined(const Fp& a, const Fp& b) { return mul_limb_schedule(a, b); }
#endif
    friend SWB_HD Fp operator*(const Fp& a, const Fp& b) {
#if defined(__CUDA_ARCH__) && defined(SWB_FP_NOINLINE_MUL)
        return mul_outlined(a, b);
#elif defined(__CUDA_ARCH__)
        return mul_limb_schedule(a, b);
#else
        return mul_host64(a, b);
#endif
    }
    SWB_HD Fp sqr() const { return (*this) * (*this); }

    // canonical integer <-> Montgomery (ark_ff into_repr / from_repr)
    SWB_HD Fp to_canonical() const {
        Fp o = zero();
        o.l[0] = 1;
        return (*this) * o;
    }
    SWB_HD Fp from_canonical() const { return (*this) * r_squared(); }

    // a^e for a small public exponent
    SWB_HD Fp pow_u64(uint64_t e) const {
        Fp acc = one(), base = *this;
        while (e) {
            if (e & 1) acc = acc * base;
            base = base.sqr();
            e >>= 1;
        }
        return acc;
    }
    // Inverse by the binary extended Euclidean algorithm on the raw limbs: ~3 * bits shift/add steps on the ALU
    // pipe instead of ~1.5 * bits Montgomery products on the multiplier -- for the lone thread that inverts on behalf
    // of a whole block (msm_pairs.cu) it is several times faster and leaves the multiplier to the other warps.
    // Input and output in Montgomery form; zero maps to zero.  Variable time (nothing secret is inverted here).
    SWB_HD Fp inverse_bingcd() const {
        if (is_zero()) return *this;
        uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = l[i]; v[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
        x1[0] = 1;
        auto is_one = [](const uint32_t* a) {
            uint32_t o = a[0] ^ 1u;
#pragma unroll
            for (int i = 1; i < N; i++) o |= a[i];
            return o == 0;
        };
        auto shr1 = [](uint32_t* a, uint32_t top) {           // a = (top:a) >> 1
#pragma unroll
            for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
            a[N - 1] = (a[N - 1] >> 1) | (top << 31);
        };
        auto halve_mod = [&](uint32_t* x) {                   // x = x / 2 mod p
            uint32_t carry = 0;
            if (x[0] & 1u) {
                uint64_t c = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    c += (uint64_t)x[i] + P::mod(i);
                    x[i] = (uint32_t)c;
                    c >>= 32;
                }
                carry = (uint32_t)c;
            }
            shr1(x, carry);
        };
        auto geq = [](const uint32_t* a, const uint32_t* b) {
            for (int i = N - 1; i >= 0; i--)
                if (a[i] != b[i]) return a[i] > b[i];
            return true;
        };
        auto sub = [](uint32_t* a, const uint32_t* b) {        // a -= b (a >= b)
            uint64_t br = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                const uint64_t d = (uint64_t)a[i] - b[i] - br;
                a[i] = (uint32_t)d;
                br = (d >> 32) & 1u;
            }
        };
        auto sub_mod = [&](uint32_t* a, const uint32_t* b) {   // a = a - b mod p
            uint64_t br = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                const uint64_t d = (uint64_t)a[i] - b[i] - br;
                a[i] = (uint32_t)d;
                br = (d >> 32) & 1u;
            }
            if (br) {
                uint64_t c = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    c += (uint64_t)a[i] + P::mod(i);
                    a[i] = (uint32_t)c;
                    c >>= 32;
                }
            }
        };
        while (!is_one(u) && !is_one(v)) {
            while (!(u[0] & 1u)) { shr1(u, 0); halve_mod(x1); }
            while (!(v[0] & 1u)) { shr1(v, 0); halve_mod(x2); }
            if (geq(u, v)) { sub(u, v); sub_mod(x1, x2); }
            else { sub(v, u); sub_mod(x2, x1); }
        }
        // x = (aR)^-1 as a plain integer; one product with R^3 turns it into a^-1 R
        Fp x;
        const uint32_t* src = is_one(u) ? x1 : x2;
#pragma unroll
        for (int i = 0; i < N; i++) x.l[i] = src[i];
        return x * r_cubed();
    }
    // Fermat inverse (p - 2); zero maps to zero.
    SWB_HD Fp inverse() const {
        uint32_t e[N];
        e[0] = ptx::sub_cc(P::mod(0), 2u);
#pragma unroll
        for (int i = 1; i < N; i++) e[i] = ptx::subc_cc(P::mod(i), 0u);
        Fp acc = one();
        for (int i = N * 32 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * (*this);
        }
        return acc;
    }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace swb
