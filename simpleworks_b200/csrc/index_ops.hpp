// Device-side Marlin indexer (index_ops.cu): joint arithmetisation of three CSR matrices.
#pragma once
#include "ctx.hpp"

namespace swb {

struct IndexCsr {               // one constraint matrix on the device; columns already shifted by the instance padding
    const uint32_t* start;      // [nrows + 1]
    const uint32_t* col;        // [nnz]
    const Fr* coef;             // [nnz]
    size_t nnz;
};
struct IndexJoint {             // result of the first phase (scratch of the context, valid until the next index call)
    uint32_t total = 0;         // matrix entries of A, B, C together
    uint32_t nnz = 0;           // joint non-zero positions = num_non_zero of the index
    size_t kn = 1;              // |K|
    uint32_t *er = nullptr, *epos = nullptr;
    const uint32_t *ids = nullptr, *scan = nullptr, *ent_row = nullptr;
};

// phase 1: sorts the entries by (row, column) and counts the joint non-zero positions (one device->host word)
int index_joint_dev(swb_ctx* c, const IndexCsr in[3], uint32_t nrows, uint32_t nvar, uint32_t nx, uint32_t nh, IndexJoint* out);
// phase 2: the six evaluation vectors on K (kn elements each; va, vb, vc are also used as scratch) and the
// column-grouped copy t_start [nh + 1], t_row / t_tag / t_coef [total]
int index_fill_dev(swb_ctx* c, const IndexCsr in[3], uint32_t nrows, uint32_t nx, uint32_t nh, const IndexJoint& j, const Fr* hel,
                   const Fr& size_inv, Fr* row, Fr* col, Fr* va, Fr* vb, Fr* vc, Fr* rowcol, uint32_t* t_start, uint32_t* t_row,
                   uint8_t* t_tag, Fr* t_coef);

}  // namespace swb
