// Internal launchers of the device-resident polynomial kernels (polyops.cu).
#pragma once
#include "ctx.hpp"

namespace swb {

int poly_mul_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n);                       // a[i] *= b[i]
int poly_add_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n);                       // a[i] += b[i]
int poly_sub_ew(swb_ctx* c, Fr* a, const Fr* b, size_t n);                       // a[i] -= b[i]
int poly_add_scaled_ew(swb_ctx* c, Fr* a, const Fr& s, const Fr* b, size_t n);   // a[i] += s * b[i]
int poly_scale_ew(swb_ctx* c, Fr* a, const Fr& s, size_t n);                     // a[i] *= s
int poly_lin_ew(swb_ctx* c, Fr* a, const Fr& c0, const Fr& c1, size_t n);        // a[i] = c0 + c1 * a[i]
int poly_eval_dev(swb_ctx* c, const Fr* p, size_t n, const Fr& x, Fr* out);      // Horner, result on the host
int poly_div_vanishing_dev(swb_ctx* c, Fr* q, Fr* r, const Fr* p, size_t len, size_t n);   // p = q (X^n - 1) + r
int poly_div_linear_dev(swb_ctx* c, Fr* q, const Fr* p, size_t n, const Fr& z);  // q = (p - p(z)) / (X - z)
int poly_len_dev(swb_ctx* c, const Fr* p, size_t n, size_t* len);                // degree + 1 (0 for zero)
int poly_powers_dev(swb_ctx* c, Fr* out, size_t n, const Fr& g);                 // out[i] = g^i
// implemented in ntt.cu / msm.cu / vec.cu: device-pointer entry points used by the engine
}  // namespace swb
