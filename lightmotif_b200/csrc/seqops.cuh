// seqops.cuh -- launchers of the encode / stripe / wrap / synth kernels.
#pragma once

#include "common.cuh"

namespace lm {

// dst[i] = code(src[i]); invalid bytes -> default symbol, *first_bad =
// min(*first_bad, index of an invalid byte) (initialise to ~0ull).
cudaError_t launch_encode(const uint8_t* src, uint64_t n, uint8_t* dst, int alphabet,
                          unsigned long long* first_bad, int sm_count, cudaStream_t stream);
// Rows [row0, row0 + out_rows) of the striped matrix (global row count
// `rows`) of the n-symbol sequence `src` (symbols, or ASCII when from_ascii).
cudaError_t launch_stripe(const uint8_t* src, bool from_ascii, uint64_t n, uint64_t rows, uint64_t row0,
                          uint64_t out_rows, uint8_t* dst, int alphabet,
                          unsigned long long* first_bad, int sm_count, cudaStream_t stream);
// Fill striped rows [r_begin, r_end) (>= rows) from the base rows.
cudaError_t launch_wrap(uint8_t* data, uint64_t rows, uint64_t r_begin, uint64_t r_end, int alphabet,
                        cudaStream_t stream);
// dst[i] = lut32[src[i] & 31] over n_bytes (a multiple of 16) bytes: residue -> class code.
cudaError_t launch_classmap(const uint8_t* src, uint64_t n_bytes, uint8_t* dst, const uint8_t* lut32, int sm_count,
                            cudaStream_t stream);
cudaError_t launch_synth(uint8_t* dst, uint64_t out_rows, uint64_t rows, uint64_t row0,
                         uint64_t total_length, uint64_t seed, int alphabet, int sm_count,
                         cudaStream_t stream);

}  // namespace lm
