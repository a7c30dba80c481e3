// Device primitives of the MSM bucketing stage (radix_sort.cu).
#pragma once
#include "ctx.hpp"

namespace swb {

// exclusive prefix sum, in place
int exclusive_scan_u32(swb_ctx* c, uint32_t* data, size_t n);
// stable LSD radix sort of (key, value) pairs on the low key_bits bits, independently inside each of
// nseg consecutive segments of seg_len pairs; ping-pongs between the given buffers and reports where
// the sorted arrays ended up
int radix_sort_segmented(swb_ctx* c, uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t seg_len,
                         uint32_t nseg, int key_bits, uint32_t** sorted_keys, uint32_t** sorted_vals);

}  // namespace swb
