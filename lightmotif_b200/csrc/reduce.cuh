// reduce.cuh -- argmax / max / threshold over materialised score matrices.
#pragma once

#include "common.cuh"

namespace lm {

// colkeys[32] (zero-initialised): max over rows of (score_key << 32 | row);
// u8 scores are keyed as score + 1.  first_cell = raw bits of cell (0, 0).
cudaError_t launch_colmax(const void* scores, bool is_u8, uint64_t rows, uint64_t stride,
                          unsigned long long* colkeys, uint32_t* first_cell, int sm_count,
                          cudaStream_t stream);
// all cells >= t (thr_bits: f32 bits or u8 value) appended to hits[0..cap);
// *count is the total.
cudaError_t launch_threshold(const void* scores, bool is_u8, uint64_t rows, uint64_t stride,
                             uint32_t thr_bits, lm_hit* hits, unsigned long long cap,
                             unsigned long long* count, int sm_count, cudaStream_t stream);

}  // namespace lm
