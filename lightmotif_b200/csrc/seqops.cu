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
// ASCII (true) of the whole sequence (n bytes) and is 4-byte aligned.
//
// One CTA tile = ST_ROWS striped rows x 32 columns.  Column c of the tile is the
// contiguous source range S'[c * rows + r .. + ST_ROWS): phase 1 streams the 32
// ranges with fully coalesced 4-byte loads (a warp reads 128 contiguous bytes
// per instruction, starting at the enclosing 4-byte boundary), encodes a word at
// a time, and parks them in shared memory column-major; phase 2 re-aligns each
// column with a funnel shift, transposes 4x4 byte blocks with PRMT and writes
// 32-byte rows.  HBM-bound: 1 byte read + 1 byte written per symbol.
#ifndef LM_ST_ROWS
#define LM_ST_ROWS 512
#endif
constexpr int ST_ROWS = LM_ST_ROWS;        // rows per CTA tile
constexpr int ST_THREADS = 256;
constexpr int ST_COLW = ST_ROWS / 4 + 3;   // words per column in shared memory: odd => conflict-free

// 4 ASCII bytes -> 4 DNA codes, fast path: (x >> 1) & 3 maps A, C, T, G to 0, 1, 2, 3; the word is
// accepted iff rebuilding the letters from the codes gives the input back.  Returns false when any
// byte needs the table (N, lower case, anything else).
__device__ __forceinline__ bool dna_word_fast(uint32_t w, uint32_t& code) {
  code = (w >> 1) & 0x03030303u;
  uint32_t sel;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(sel) : "r"(code), "r"(0x00001001u), "r"(0u));  // nibbles c0 c2 c1 c3
  const uint32_t letters = __byte_perm(0x47544341u /* "ACTG" */, 0u, sel);
  return letters == __byte_perm(w, 0u, 0x3120);
}

#ifndef LM_ST_MINB
#define LM_ST_MINB 4
#endif
template <bool FROM_ASCII>
__global__ void __launch_bounds__(ST_THREADS, LM_ST_MINB) stripe_kernel(const uint8_t* __restrict__ src, uint64_t n,
                                                            uint64_t rows, uint64_t row0,
                                                            uint64_t out_rows, uint8_t* __restrict__ dst,
                                                            int alphabet, uint8_t pad,
                                                            unsigned long long* first_bad) {
  __shared__ uint8_t lut[256];
  __shared__ uint32_t tile[32][ST_COLW];  // [column][word]; column c starts `shift[c]` bytes into its first word
  __shared__ uint32_t shift[32];
  if (FROM_ASCII) build_lut(lut, alphabet);
  unsigned long long bad = ~0ull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t ntiles = (out_rows + ST_ROWS - 1) / ST_ROWS;
  const uint32_t pad4 = pad * 0x01010101u;
  const bool dna = alphabet == LM_DNA;
  const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
  constexpr int NJ = (ST_COLW + 31) / 32, NC = 32 / (ST_THREADS / 32);
  uint32_t raw[NC][NJ];

  // All of a thread's loads for a tile (4 columns x 5 words) are issued before any is consumed: the
  // bytes in flight, not the instruction rate, decide how much of the HBM bandwidth this kernel sees.
  auto fetch = [&](uint64_t t) {
    const uint64_t r0 = t * ST_ROWS;
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const int c = warp + ci * (ST_THREADS / 32);
      const uint64_t w_first = ((uint64_t)c * rows + row0 + r0) >> 2;  // enclosing source word of the segment
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int wi = j * 32 + lane;
        raw[ci][j] = (wi < ST_COLW && (w_first + wi) * 4 < n) ? __ldg(&src32[w_first + wi]) : pad4;
      }
    }
  };

  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t r0 = t * ST_ROWS;  // local row of the tile
    // Interior tiles (every word of all 32 column segments lies inside the sequence: all but the last
    // few of each column) take a path without any per-word bounds logic -- the general path below
    // spends ~18 instructions per symbol on 64-bit index checks and was issue-bound at 44 % of the
    // HBM rate.
    const uint64_t p_last = (uint64_t)31 * rows + row0 + r0;  // first symbol of the last column's segment
    const bool interior = ((p_last >> 2) + ST_COLW) * 4 <= n && (!FROM_ASCII || dna);
    if (interior) {
#pragma unroll
      for (int ci = 0; ci < NC; ++ci) {
        const int c = warp + ci * (ST_THREADS / 32);
        const uint32_t* seg = src32 + (((uint64_t)c * rows + row0 + r0) >> 2);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int wi = j * 32 + lane;
          raw[ci][j] = wi < ST_COLW ? __ldg(seg + wi) : 0u;
        }
      }
#pragma unroll
      for (int ci = 0; ci < NC; ++ci) {
        const int c = warp + ci * (ST_THREADS / 32);
        const uint64_t p_first = (uint64_t)c * rows + row0 + r0;
        if (lane == 0) shift[c] = (uint32_t)(p_first & 3);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int wi = j * 32 + lane;
          if (wi >= ST_COLW) break;
          uint32_t out = raw[ci][j];
          if (FROM_ASCII) {
            uint32_t code;
            if (dna_word_fast(out, code)) {
              out = code;
            } else {  // N, lower case, anything else: byte by byte through the table
              const uint32_t w = out;
              const uint64_t p0 = ((p_first >> 2) + wi) * 4;
              out = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                uint32_t cde = lut[(w >> (8 * b)) & 0xFF];
                if (cde == 0xFF) {
                  if (p0 + b < bad) bad = p0 + b;
                  cde = pad;
                }
                out |= cde << (8 * b);
              }
            }
          }
          tile[c][wi] = out;
        }
      }
    } else {
    fetch(t);

    // ---- phase 1: encode the 32 column segments (ST_ROWS consecutive symbols each) into shared memory
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const int c = warp + ci * (ST_THREADS / 32);
      const uint64_t p_first = (uint64_t)c * rows + row0 + r0;  // sequence index of the segment's first symbol
      const uint64_t w_first = p_first >> 2;
      if (lane == 0) shift[c] = (uint32_t)(p_first & 3);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int wi = j * 32 + lane;
        if (wi >= ST_COLW) break;
        const uint64_t p0 = (w_first + wi) * 4;  // sequence index of byte 0 of this word
        const uint32_t w = raw[ci][j];
        uint32_t out = pad4;
        if (p0 < n) {
          if (FROM_ASCII) {
            uint32_t code;
            if (dna && p0 + 3 < n && dna_word_fast(w, code)) {
              out = code;
            } else {
              out = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                uint32_t cde = pad;
                if (p0 + b < n) {
                  cde = lut[(w >> (8 * b)) & 0xFF];
                  if (cde == 0xFF) {
                    if (p0 + b < bad) bad = p0 + b;  // any invalid byte of the sequence may be reported: the minimum wins
                    cde = pad;
                  }
                }
                out |= cde << (8 * b);
              }
            }
          } else if (p0 + 3 < n) {
            out = w;
          } else {
            out = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) out |= (p0 + b < n ? (w >> (8 * b)) & 0xFFu : (uint32_t)pad) << (8 * b);
          }
        }
        tile[c][wi] = out;
      }
    }
    }  // general path
    __syncthreads();
    // ---- phase 2: re-align, 4x4 byte transposes; a thread assembles 16 bytes of 4 consecutive rows
    // (half a striped row each) and writes them with 16-byte stores
    for (int blk = threadIdx.x; blk < (ST_ROWS / 4) * 2; blk += ST_THREADS) {
      const int half = blk & 1, rb = blk >> 1;
      uint32_t o[4][4];  // [row k][column block]
#pragma unroll
      for (int cbl = 0; cbl < 4; ++cbl) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 16 * half + 4 * cbl + i;
          w[i] = __funnelshift_r(tile[c][rb], tile[c][rb + 1], 8 * shift[c]);
        }
        const uint32_t u = __byte_perm(w[0], w[1], 0x5140), v = __byte_perm(w[2], w[3], 0x5140);
        const uint32_t x = __byte_perm(w[0], w[1], 0x7362), y = __byte_perm(w[2], w[3], 0x7362);
        o[0][cbl] = __byte_perm(u, v, 0x5410);
        o[1][cbl] = __byte_perm(u, v, 0x7632);
        o[2][cbl] = __byte_perm(x, y, 0x5410);
        o[3][cbl] = __byte_perm(x, y, 0x7632);
      }
      const uint64_t rbase = r0 + 4 * rb;
      if (rbase + 3 < out_rows) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(dst + (rbase + k) * 32 + 16 * half) = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (rbase + k < out_rows)
            *reinterpret_cast<uint4*>(dst + (rbase + k) * 32 + 16 * half) = make_uint4(o[k][0], o[k][1], o[k][2], o[k][3]);
      }
    }
    __syncthreads();
  }
  if (FROM_ASCII && bad != ~0ull) atomicMin(first_bad, bad);
}

// Residue -> class code (protein filter, prune.cuh): dst[i] = lut[src[i]], 16 bytes per thread.
// HBM-bound 1 + 1 bytes per residue.
struct ClassLut { uint8_t v[32]; };
__global__ void __launch_bounds__(ENC_THREADS) classmap_kernel(const uint4* __restrict__ src, uint64_t nvec,
                                                                uint4* __restrict__ dst, const ClassLut lut) {
  __shared__ uint8_t sl[32];
  if (threadIdx.x < 32) sl[threadIdx.x] = lut.v[threadIdx.x];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;  // independent 16-byte loads in flight per thread
  for (uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v0 + u * stride < nvec) x[u] = __ldg(&src[v0 + u * stride]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (v0 + u * stride >= nvec) break;
      uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) o |= (uint32_t)sl[(w[q] >> (8 * b)) & 31] << (8 * b);
        w[q] = o;
      }
      dst[v0 + u * stride] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
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
  unsigned grid = (unsigned)(ntiles < (uint64_t)sm_count * 16 ? ntiles : (uint64_t)sm_count * 16);
  if (from_ascii)
    stripe_kernel<true><<<grid, ST_THREADS, 0, stream>>>(src, n, rows, row0, out_rows, dst, alphabet, pad, first_bad);
  else
    stripe_kernel<false><<<grid, ST_THREADS, 0, stream>>>(src, n, rows, row0, out_rows, dst, alphabet, pad, first_bad);
  return cudaGetLastError();
}

cudaError_t launch_classmap(const uint8_t* src, uint64_t n_bytes, uint8_t* dst, const uint8_t* lut32, int sm_count,
                            cudaStream_t stream) {
  if (n_bytes == 0) return cudaSuccess;
  ClassLut lut;
  for (int i = 0; i < 32; ++i) lut.v[i] = lut32[i];
  const uint64_t nvec = n_bytes / 16;  // striped matrices are multiples of 32 bytes
  classmap_kernel<<<grid_for((nvec + 3) / 4, ENC_THREADS, sm_count, 8), ENC_THREADS, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), nvec, reinterpret_cast<uint4*>(dst), lut);
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
