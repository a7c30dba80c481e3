// Internal launchers of the prover-side device kernels that are not plain polynomial arithmetic
// (marlin_ops.cu): bulk field-element sampling from a ChaCha stream, sparse matrix-vector products
// over Fr, and the witness layout on H.
#pragma once
#include "ctx.hpp"

namespace swb {

// n consecutive `Fr::rand` draws from the ChaCha word stream (key, rounds) starting at word `pos`:
// candidate k is words [pos + 8k, pos + 8k + 8) with the top 3 bits cleared, accepted when < r and
// kept as the Montgomery representation (ark-ff 0.3 `UniformRand for Fp256`).  *words_used is how far
// the stream advanced (8 x candidates consumed, rejected ones included).
int rand_fr_dev(swb_ctx* c, Fr* out, size_t n, const uint32_t key[8], int rounds, uint64_t pos, uint64_t* words_used);

// out[r] = sum_{k in [start[r], start[r+1])} w(k) * coef[k] * x[col[k]] for r < nrows, and 0 for
// nrows <= r < nout; w(k) = weights[tag[k]] when tag != nullptr, else 1
int csr_spmv_dev(swb_ctx* c, Fr* out, size_t nout, size_t nrows, const uint32_t* start, const uint32_t* col, const Fr* coef,
                 const uint8_t* tag, const Fr* x, const Fr weights[3]);

// Marlin's w on H before interpolation: position k of H holds 0 when k is in the X-subdomain
// (k % ratio == 0), else z[ninst + (k - k/ratio - 1)] - xh[k] (0 for the witness beyond nvars)
int witness_evals_dev(swb_ctx* c, Fr* out, size_t nh, size_t ratio, const Fr* z, size_t ninst, size_t nvars, const Fr* xh);

}  // namespace swb
