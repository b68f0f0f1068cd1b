// seqops.cu -- device kernels for the integer (bit-exact) sequence stages.
//
// Replaces  Encode::encode_into   lightmotif/src/pli/mod.rs:56-66, avx2.rs:37-99
//           encode_lossy          lightmotif/src/seq.rs:122-129
//           Stripe::stripe_into   pli/mod.rs:178-200, avx2.rs:561-799
//           configure_wrap        seq.rs:369-381
//
// Layout fact used throughout (proved against the reference's literal test
// matrices in tests/): after stripe + configure_wrap, cell (r, c) of the
// [rows + wrap][32] matrix holds S'(c * rows + r), where S'(p) is symbol p of
// the sequence, or the default symbol (N / X) for p >= n.  This closed form
// also covers the cascading case wrap > rows of seq.rs:373-378.
#include "seqops.cuh"

namespace lm {

namespace {

constexpr int ENC_THREADS = 256;

// ASCII -> symbol LUT; 0xFF = not in the alphabet (abc.rs:161-171, 298-324:
// exact upper-case match only)
__device__ __forceinline__ void build_lut(uint8_t* lut, int alphabet) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = 0xFF;
  __syncthreads();
  const char* dna = "ACTGN";
  const char* prot = "ACDEFGHIKLMNPQRSTVWYX";
  const char* s = alphabet == LM_DNA ? dna : prot;
  const int k = alphabet_k(alphabet);
  if ((int)threadIdx.x < k) lut[(uint8_t)s[threadIdx.x]] = (uint8_t)threadIdx.x;
  __syncthreads();
}

// Encode: 16 bytes per thread.  Invalid bytes map to `pad` (lossy) and the
// smallest invalid index is recorded with atomicMin.
__global__ void __launch_bounds__(ENC_THREADS) encode_kernel(const uint8_t* __restrict__ src,
                                                             uint64_t n, uint8_t* __restrict__ dst,
                                                             int alphabet, uint8_t pad,
                                                             unsigned long long* first_bad) {
  __shared__ uint8_t lut[256];
  build_lut(lut, alphabet);
  unsigned long long bad = ~0ull;
  const uint64_t nvec = n / 16;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t start = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    for (uint64_t v = start; v < nvec; v += stride) {
      uint4 x = reinterpret_cast<const uint4*>(src)[v];
      uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint8_t code = lut[(w[q] >> (8 * b)) & 0xFF];
          if (code == 0xFF) {
            unsigned long long idx = v * 16 + q * 4 + b;
            bad = idx < bad ? idx : bad;
            code = pad;
          }
          o |= (uint32_t)code << (8 * b);
        }
        w[q] = o;
      }
      reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  // tail (or everything when the buffers are not 16-byte aligned)
  for (uint64_t i = (aligned ? nvec * 16 : 0) + start; i < n; i += stride) {
    uint8_t code = lut[src[i]];
    if (code == 0xFF) { bad = i < bad ? i : bad; code = pad; }
    dst[i] = code;
  }
  if (bad != ~0ull) atomicMin(first_bad, bad);
}

// Fused (encode +) stripe + wrap: writes rows [0, out_rows) of the striped
// matrix of a sequence whose GLOBAL striped matrix has `rows` rows; the slab
// starts at global row `row0`.  src holds symbols (FROM_ASCII = false) or raw
// ASCII (true) of the whole sequence (n bytes).
constexpr int ST_ROWS = 256;  // rows per CTA tile
constexpr int ST_THREADS = 256;

template <bool FROM_ASCII>
__global__ void __launch_bounds__(ST_THREADS) stripe_kernel(const uint8_t* __restrict__ src, uint64_t n,
                                                            uint64_t rows, uint64_t row0,
                                                            uint64_t out_rows, uint8_t* __restrict__ dst,
                                                            int alphabet, uint8_t pad,
                                                            unsigned long long* first_bad) {
  __shared__ uint8_t lut[256];
  __shared__ __align__(16) uint8_t tile[32][ST_ROWS + 4];  // [column][row], +4 keeps word alignment, shifts banks
  if (FROM_ASCII) build_lut(lut, alphabet);
  unsigned long long bad = ~0ull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t ntiles = (out_rows + ST_ROWS - 1) / ST_ROWS;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t r0 = t * ST_ROWS;  // local row of the tile
    // phase 1: 32 column segments of ST_ROWS consecutive symbols each
    for (int c = warp; c < 32; c += ST_THREADS / 32) {
      for (int k = lane; k < ST_ROWS; k += 32) {
        uint64_t r = row0 + r0 + k;             // global striped row
        uint64_t p = (uint64_t)c * rows + r;    // sequence index S'(c*rows + r)
        uint8_t code = pad;
        if (p < n && r0 + k < out_rows) {
          uint8_t x = src[p];
          if (FROM_ASCII) {
            code = lut[x];
            if (code == 0xFF) { bad = p < bad ? p : bad; code = pad; }
          } else {
            code = x;
          }
        }
        tile[c][k] = code;
      }
    }
    __syncthreads();
    // phase 2: 4x4 byte transposes, 32-byte rows written as 8 x 4 bytes
    for (int blk = threadIdx.x; blk < (ST_ROWS / 4) * 8; blk += ST_THREADS) {
      int cb = blk & 7, rb = blk >> 3;
      uint32_t w0 = *reinterpret_cast<const uint32_t*>(&tile[4 * cb + 0][4 * rb]);
      uint32_t w1 = *reinterpret_cast<const uint32_t*>(&tile[4 * cb + 1][4 * rb]);
      uint32_t w2 = *reinterpret_cast<const uint32_t*>(&tile[4 * cb + 2][4 * rb]);
      uint32_t w3 = *reinterpret_cast<const uint32_t*>(&tile[4 * cb + 3][4 * rb]);
      uint32_t u = __byte_perm(w0, w1, 0x5140), v = __byte_perm(w2, w3, 0x5140);
      uint32_t x = __byte_perm(w0, w1, 0x7362), y = __byte_perm(w2, w3, 0x7362);
      uint32_t o[4] = {__byte_perm(u, v, 0x5410), __byte_perm(u, v, 0x7632),
                       __byte_perm(x, y, 0x5410), __byte_perm(x, y, 0x7632)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t r = r0 + 4 * rb + k;
        if (r < out_rows) *reinterpret_cast<uint32_t*>(dst + r * 32 + 4 * cb) = o[k];
      }
    }
    __syncthreads();
  }
  if (FROM_ASCII && bad != ~0ull) atomicMin(first_bad, bad);
}

// configure_wrap on an existing device matrix: rows [rows + old_wrap, rows + new_wrap)
// cell (r, c) = base cell (r % rows, c + r / rows), default symbol past column 31.
__global__ void wrap_kernel(uint8_t* data, uint64_t rows, uint64_t r_begin, uint64_t r_end, uint8_t pad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t n = (r_end - r_begin) * 32;
  if (i >= n) return;
  uint64_t r = r_begin + i / 32, c = i % 32;
  uint64_t cc = c + r / rows, rr = r % rows;
  data[r * 32 + c] = cc < 32 ? data[rr * 32 + cc] : pad;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// synthetic striped slab: cell (r, c) = synth(c * rows + row0 + r)
__global__ void synth_kernel(uint8_t* dst, uint64_t out_rows, uint64_t rows, uint64_t row0,
                             uint64_t total_length, uint64_t seed, int alphabet, uint8_t pad) {
  uint64_t nwords = out_rows * 8;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = w / 8;
    uint32_t cb = (uint32_t)(w % 8) * 4;
    uint32_t o = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint64_t p = (uint64_t)(cb + b) * rows + row0 + r;
      uint8_t code = pad;
      if (p < total_length) {
        if (alphabet == LM_DNA) {
          uint64_t h = splitmix64(seed + (p >> 5));
          code = (uint8_t)((h >> (2 * (p & 31))) & 3);
        } else {
          code = (uint8_t)((splitmix64(seed + p) >> 11) % 20);
        }
      }
      o |= (uint32_t)code << (8 * b);
    }
    reinterpret_cast<uint32_t*>(dst)[w] = o;
  }
}

inline unsigned grid_for(uint64_t work_items, int threads, int sm_count, int per_sm = 8) {
  uint64_t g = (work_items + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

cudaError_t launch_encode(const uint8_t* src, uint64_t n, uint8_t* dst, int alphabet,
                          unsigned long long* first_bad, int sm_count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  uint8_t pad = (uint8_t)(alphabet_k(alphabet) - 1);
  encode_kernel<<<grid_for((n + 15) / 16, ENC_THREADS, sm_count), ENC_THREADS, 0, stream>>>(
      src, n, dst, alphabet, pad, first_bad);
  return cudaGetLastError();
}

cudaError_t launch_stripe(const uint8_t* src, bool from_ascii, uint64_t n, uint64_t rows, uint64_t row0,
                          uint64_t out_rows, uint8_t* dst, int alphabet,
                          unsigned long long* first_bad, int sm_count, cudaStream_t stream) {
  if (out_rows == 0) return cudaSuccess;
  uint8_t pad = (uint8_t)(alphabet_k(alphabet) - 1);
  uint64_t ntiles = (out_rows + ST_ROWS - 1) / ST_ROWS;
  unsigned grid = (unsigned)(ntiles < (uint64_t)sm_count * 4 ? ntiles : (uint64_t)sm_count * 4);
  if (from_ascii)
    stripe_kernel<true><<<grid, ST_THREADS, 0, stream>>>(src, n, rows, row0, out_rows, dst, alphabet, pad, first_bad);
  else
    stripe_kernel<false><<<grid, ST_THREADS, 0, stream>>>(src, n, rows, row0, out_rows, dst, alphabet, pad, first_bad);
  return cudaGetLastError();
}

cudaError_t launch_wrap(uint8_t* data, uint64_t rows, uint64_t r_begin, uint64_t r_end, int alphabet,
                        cudaStream_t stream) {
  if (r_end <= r_begin) return cudaSuccess;
  uint8_t pad = (uint8_t)(alphabet_k(alphabet) - 1);
  uint64_t n = (r_end - r_begin) * 32;
  wrap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(data, rows, r_begin, r_end, pad);
  return cudaGetLastError();
}

cudaError_t launch_synth(uint8_t* dst, uint64_t out_rows, uint64_t rows, uint64_t row0,
                         uint64_t total_length, uint64_t seed, int alphabet, int sm_count,
                         cudaStream_t stream) {
  if (out_rows == 0) return cudaSuccess;
  uint8_t pad = (uint8_t)(alphabet_k(alphabet) - 1);
  synth_kernel<<<grid_for(out_rows * 8, 256, sm_count, 16), 256, 0, stream>>>(
      dst, out_rows, rows, row0, total_length, seed, alphabet, pad);
  return cudaGetLastError();
}

}  // namespace lm
