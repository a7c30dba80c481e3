/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the swb200 CUDA library.
 *
 * Plain-C restatement of the arkworks 0.3 algorithms that simpleworks' Marlin wrapper reaches
 * through src/marlin/mod.rs:52,75,85,92 (universal_setup / prove / verify / index).  The crates
 * themselves (ark-ff, ark-ec, ark-poly, ark-bls12-377 ^0.3.0 -- reference Cargo.toml:15-27) are
 * not vendored under /root/reference and cannot be built here (no Rust toolchain), so:
 *
 *     PARITY UNPINNED against real arkworks.  Pinned against oracle/golden.py (python big
 *     ints) and the fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (simpleworks_b200/, libswb200.so) must never link or call it.
 */
#ifndef SWB_ORACLE_H
#define SWB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } fr_t;   /* ark_ff::Fp256<FrParameters>, Montgomery form */
typedef struct { uint64_t l[6]; } fq_t;   /* ark_ff::Fp384<FqParameters>, Montgomery form */
typedef struct { uint64_t l[4]; } big256_t; /* ark_ff::BigInteger256, canonical integer   */
typedef struct { fq_t x, y; uint8_t infinity; uint8_t _pad[7]; } g1_affine_t;  /* GroupAffine */
typedef struct { fq_t x, y, z; } g1_jac_t;                                     /* GroupProjective (Jacobian) */

/* --- field (ark-ff/src/fields/macros.rs) ------------------------------------------------ */
void orc_fr_mul(fr_t *r, const fr_t *a, const fr_t *b);
void orc_fr_add(fr_t *r, const fr_t *a, const fr_t *b);
void orc_fr_sub(fr_t *r, const fr_t *a, const fr_t *b);
void orc_fr_inv(fr_t *r, const fr_t *a);
void orc_fr_from_canon(fr_t *r, const big256_t *a);
void orc_fr_to_canon(big256_t *r, const fr_t *a);
void orc_fq_mul(fq_t *r, const fq_t *a, const fq_t *b);
void orc_fq_add(fq_t *r, const fq_t *a, const fq_t *b);
void orc_fq_sub(fq_t *r, const fq_t *a, const fq_t *b);
void orc_fq_inv(fq_t *r, const fq_t *a);
void orc_fq_from_canon(fq_t *r, const uint64_t a[6]);
void orc_fq_to_canon(uint64_t r[6], const fq_t *a);
/* vector forms, for bulk parity tests */
void orc_fr_mul_vec(fr_t *r, const fr_t *a, const fr_t *b, size_t n);
void orc_fq_mul_vec(fq_t *r, const fq_t *a, const fq_t *b, size_t n);
void orc_fr_batch_inverse(fr_t *v, size_t n);      /* ark_ff::batch_inversion (zeros skipped) */

/* --- G1 (ark-ec/src/models/short_weierstrass_jacobian.rs) ------------------------------- */
void orc_g1_generator(g1_affine_t *g);
void orc_g1_jac_zero(g1_jac_t *r);
void orc_g1_from_affine(g1_jac_t *r, const g1_affine_t *a);
void orc_g1_add_mixed(g1_jac_t *r, const g1_affine_t *b);       /* add_assign_mixed */
void orc_g1_add(g1_jac_t *r, const g1_jac_t *b);                /* add_assign       */
void orc_g1_double(g1_jac_t *r);                                /* double_in_place  */
void orc_g1_to_affine(g1_affine_t *r, const g1_jac_t *a);       /* into_affine      */
void orc_g1_mul(g1_jac_t *r, const g1_affine_t *base, const big256_t *k);
void orc_g1_batch_normalize(g1_affine_t *out, const g1_jac_t *in, size_t n);
int  orc_g1_affine_on_curve(const g1_affine_t *a);

/* --- MSM (ark-ec/src/msm/{variable_base,fixed_base}.rs) --------------------------------- */
/* VariableBaseMSM::multi_scalar_mul(bases, scalars): Pippenger, unsigned windows of
 * c = 3 if n<32 else ceil(log2 n)*69/100+2 bits, zero skipped, one fast-pathed,
 * windows processed in parallel (rayon -> OpenMP here).  threads<=0: all cores. */
void orc_msm_variable_base(g1_jac_t *out, const g1_affine_t *bases, const big256_t *scalars,
                           size_t n, int threads);
/* powers: out[i] = beta^i * g for i<n, affine; FixedBaseMSM window table + batch normalise
 * (what KZG10::setup does for powers_of_g). */
void orc_fixed_base_powers(g1_affine_t *out, const g1_jac_t *g, const fr_t *beta, size_t n, int threads);

/* --- Radix2EvaluationDomain (ark-poly/src/domain/radix2) -------------------------------- */
/* in place, natural order in and out; inverse scales by n^-1; coset multiplies by 22^j before
 * a forward transform / by 22^-j after an inverse one.  */
void orc_ntt(fr_t *v, uint32_t log_n, int inverse, int coset, int threads);
void orc_domain_generator(fr_t *w, uint32_t log_n);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
