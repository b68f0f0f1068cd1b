// reduce.cu -- stand-alone reductions over a materialised StripedScores matrix.
//
// Replaces  Maximum::argmax / max   lightmotif/src/pli/mod.rs:135-160
//           argmax_f32_avx2         lightmotif/src/pli/platform/avx2.rs:349-426
//           argmax_u8_avx2/max_u8   avx2.rs:469-559
//           Threshold::threshold    pli/mod.rs:210-221
//
// HBM-bound streaming passes: lane = striped column, so a warp reads one
// 128-byte score row (f32) per load.  The per-column winner is the maximum of
// the packed key (order-preserving score key << 32 | row): highest score, and
// among equal scores the LAST row -- what every reference back-end keeps per
// column (avx2.rs:387 _CMP_LE_OS, pli/mod.rs:146 `>=`).
#include "reduce.cuh"

namespace lm {

namespace {

constexpr int RED_THREADS = 256;

template <bool U8>
__global__ void __launch_bounds__(RED_THREADS) colmax_kernel(const void* __restrict__ scores,
                                                             uint64_t rows, uint64_t stride,
                                                             unsigned long long* colkeys,
                                                             uint32_t* first_cell) {
  __shared__ unsigned long long red[RED_THREADS / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t wstride = (uint64_t)gridDim.x * (RED_THREADS / 32);
  unsigned long long best = 0;
  for (uint64_t r = (uint64_t)blockIdx.x * (RED_THREADS / 32) + warp; r < rows; r += wstride) {
    uint32_t sk;
    if (U8) sk = (uint32_t)reinterpret_cast<const uint8_t*>(scores)[r * stride + lane] + 1u;
    else sk = f32_to_key(reinterpret_cast<const float*>(scores)[r * stride + lane]);
    unsigned long long k = ((unsigned long long)sk << 32) | (uint32_t)r;
    best = k > best ? k : best;
  }
  red[warp][lane] = best;
  __syncthreads();
  if (warp == 0) {
    unsigned long long m = red[0][lane];
#pragma unroll
    for (int w = 1; w < RED_THREADS / 32; ++w) m = red[w][lane] > m ? red[w][lane] : m;
    if (m) atomicMax(&colkeys[lane], m);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && rows > 0) {
    if (U8) *first_cell = reinterpret_cast<const uint8_t*>(scores)[0];
    else *first_cell = __float_as_uint(reinterpret_cast<const float*>(scores)[0]);
  }
}

template <bool U8>
__global__ void __launch_bounds__(RED_THREADS) threshold_kernel(const void* __restrict__ scores,
                                                                uint64_t rows, uint64_t stride,
                                                                uint32_t thr_bits, lm_hit* hits,
                                                                unsigned long long cap,
                                                                unsigned long long* count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t wstride = (uint64_t)gridDim.x * (RED_THREADS / 32);
  for (uint64_t r = (uint64_t)blockIdx.x * (RED_THREADS / 32) + warp; r < rows; r += wstride) {
    bool hit;
    float sc;
    if (U8) {
      uint32_t s = reinterpret_cast<const uint8_t*>(scores)[r * stride + lane];
      hit = s >= thr_bits;
      sc = (float)s;
    } else {
      sc = reinterpret_cast<const float*>(scores)[r * stride + lane];
      hit = sc >= __uint_as_float(thr_bits);
    }
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) {
        unsigned long long slot = base + __popc(mask & ((1u << lane) - 1));
        if (slot < cap) {
          lm_hit h;
          h.row = r;
          h.col = lane;
          h.position = (uint64_t)lane * rows + r;
          h.score = sc;
          hits[slot] = h;
        }
      }
    }
  }
}

inline unsigned red_grid(uint64_t rows, int sm_count) {
  uint64_t g = (rows + RED_THREADS / 32 - 1) / (RED_THREADS / 32);
  uint64_t cap = (uint64_t)sm_count * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t launch_colmax(const void* scores, bool is_u8, uint64_t rows, uint64_t stride,
                          unsigned long long* colkeys, uint32_t* first_cell, int sm_count,
                          cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (is_u8) colmax_kernel<true><<<red_grid(rows, sm_count), RED_THREADS, 0, stream>>>(scores, rows, stride, colkeys, first_cell);
  else colmax_kernel<false><<<red_grid(rows, sm_count), RED_THREADS, 0, stream>>>(scores, rows, stride, colkeys, first_cell);
  return cudaGetLastError();
}

cudaError_t launch_threshold(const void* scores, bool is_u8, uint64_t rows, uint64_t stride,
                             uint32_t thr_bits, lm_hit* hits, unsigned long long cap,
                             unsigned long long* count, int sm_count, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (is_u8) threshold_kernel<true><<<red_grid(rows, sm_count), RED_THREADS, 0, stream>>>(scores, rows, stride, thr_bits, hits, cap, count);
  else threshold_kernel<false><<<red_grid(rows, sm_count), RED_THREADS, 0, stream>>>(scores, rows, stride, thr_bits, hits, cap, count);
  return cudaGetLastError();
}

}  // namespace lm
