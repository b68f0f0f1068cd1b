// score.cu -- exact PSSM scoring kernels for sm_100a.
//
// Replaces  Score::score_rows_into       lightmotif/src/pli/mod.rs:72-106
//           score_f32_avx2_permute       lightmotif/src/pli/platform/avx2.rs:101-199
//           score_f32_avx2_gather        avx2.rs:201-290
//           score_u8_avx2_shuffle        avx2.rs:292-347
// fused (optionally) with
//           Maximum::argmax              pli/mod.rs:135-155, avx2.rs:349-426
//           Threshold::threshold         pli/mod.rs:210-221
//
// Numerics: every f32 score is the reference's strictly sequential sum
// ((0 + w0) + w1) + ... + w(M-1); one accumulator per position, motif rows
// added in order, never re-associated => bit-identical to all reference
// back-ends.  M is padded to a multiple of 4 with +0.0 rows (x + 0.0 == x for
// every x the chain can produce: the chain never yields -0.0).
//
// Structure (one CTA = 8 warps, a tile = 128 striped rows x 32 columns):
//   1. one thread stages the tile's contiguous byte range
//      [(128 + m_pad) x 32 B] from HBM with a 1-D bulk TMA copy
//      (cp.async.bulk + mbarrier), double-buffered against compute;
//   2. the CTA transposes it in shared memory to column-major, pre-multiplied
//      by 4 (4x4 byte blocks, PRMT network) so a lane can read 4 consecutive
//      rows of ITS column with one conflict-free LDS.32;
//   3. lane = striped column, each thread owns 16 consecutive rows and keeps
//      their 16 accumulators in registers.  Per (position, motif row):
//         PRMT  addr = {table-group base[31:8], 4*symbol}   (ALU pipe)
//         LDS   w    = [addr + jj*KP*4]                      (<= 21 distinct
//                       consecutive words per warp: 1 wavefront)
//         FADD  acc += w                                     (FMA pipe)
//      The kernel is bound by the shared-memory pipe (1 warp-LDS/clk/SM).
//   4. epilogue per tile: coalesced 128-B stores and/or running per-column
//      best (>=, so the last row wins ties) and/or ballot-compacted hits.
//
// Paired positions (DNA f32, tiles without the default symbol N).  Step 3 costs one LDS.32 and one
// FADD per cell update: the LSU delivers 32 look-ups/clk/SM and the issue ports carry 2.3
// instructions per update, and the kernel sits at ~72 % of the former because of the latter.  Two
// ADJACENT positions of a thread (rows i, i + 1 of its column) read adjacent symbols for the same
// motif row, so a second table  T2[j][s + 4 s'] = (w_j[s], w_j[s'])  (16 float2 entries = 128 bytes
// = one bank pass per motif row, built in shared memory by the kernel itself) delivers both weights
// with ONE LDS.64, and ONE packed add (add.rn.f32x2: two independent IEEE adds) applies them to the
// accumulator pair: half the instructions per update, the same sequential order per position --
// still bit-identical.  The pair codes 8 s + 32 s' come from the pre-multiplied window words with
// a funnel shift and a multiply-add per 4 symbols.
//
// Prefix table (same tiles).  The sequential sum starts at +0.0, so a position's accumulator after its first
// four motif rows is a function of four symbols only: a 256-entry table of those exact partial sums (built by
// the kernel with the same __fadd_rn chain, one copy per lane = bank, 32 KB) replaces 4 of the M look-ups of
// every position with ONE conflict-free LDS.32 -- order-preserving, hence still bit-identical.  Later steps
// depend on the running value and cannot be tabulated.  M = 20: 341 -> 391 Gbp/s (the LSU bound moves from
// 32/20 to 32/17 positions/clk/SM).
#include <algorithm>

#include "score.cuh"

namespace lm {

namespace {

#ifndef LM_SCORE_MINBLOCKS
#define LM_SCORE_MINBLOCKS 3
#endif

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int RPT = 16;                 // rows (positions) per thread per tile
constexpr int TILE_ROWS = WARPS * RPT;  // 128
constexpr int NWIN = (RPT + 3 + 3) / 4; // 5 window words: rows i + jj, i < 16, jj < 4

struct SmemLayout {
  uint32_t table_bytes, tilet_off, tsw, raw_off, raw_bytes, bar_off, red_off, table2_off, table4_off, total;
};

constexpr uint32_t TABLE4_BYTES = 256 * 32 * 4;  // prefix table: 4^4 entries, one copy per lane (bank)

// `prefix`: the launch uses the prefix table (DNA f32, one motif)
__host__ __device__ inline SmemLayout smem_layout(int kp, int m_pad, bool prefix) {
  SmemLayout L;
  uint32_t tt = TILE_ROWS + m_pad;  // rows staged per tile (multiple of 4)
  L.table_bytes = (uint32_t)m_pad * kp * 4;  // multiple of 256
  L.tsw = tt / 4;
  if ((L.tsw & 1) == 0) L.tsw += 1;  // odd word stride => conflict-free column reads
  L.tilet_off = L.table_bytes;
  L.raw_off = L.tilet_off + 32 * L.tsw * 4;
  L.raw_off = (L.raw_off + 127) & ~127u;
  L.raw_bytes = tt * 32;
  L.bar_off = L.raw_off + 2 * L.raw_bytes;
  L.red_off = L.bar_off + 16;
  L.table2_off = (L.red_off + WARPS * 32 * 8 + 511) & ~511u;  // pair table [m_pad][16] float2 (DNA f32 only)
  L.table4_off = (L.table2_off + (kp == 16 ? (uint32_t)m_pad * 128 : 0) + 127) & ~127u;
  L.total = L.table4_off + (prefix ? TABLE4_BYTES : 0);
  return L;
}
constexpr uint32_t SMEM_LIMIT = 200 * 1024;
// the prefix table is used when the launch is DNA f32 with one motif AND it still fits (very long motifs:
// their tables take the room)
__host__ __device__ inline bool wants_prefix(int kp, bool u8, int n_motifs, int m_pad) {
  return kp == 16 && !u8 && n_motifs == 1 && smem_layout(kp, m_pad, true).total <= SMEM_LIMIT;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
  return __byte_perm(a, b, s);
}

template <int IMM, bool U8>
__device__ __forceinline__ void lookup_add(float& accf, uint32_t& accu, uint32_t word,
                                           uint32_t group_base, uint32_t sel) {
  uint32_t addr;
  asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(addr) : "r"(word), "r"(group_base), "r"(sel));
  if (U8) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    accu += v;
  } else {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    accf = __fadd_rn(accf, v);
  }
}

template <int JJ, int KP, bool U8>
__device__ __forceinline__ void row_step(float (&accf)[RPT], uint32_t (&accu)[RPT],
                                         const uint32_t (&W)[NWIN], uint32_t gbase) {
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int b = i + JJ;  // window byte = row offset of (position i, motif row 4*j4 + JJ)
    lookup_add<JJ * KP * 4, U8>(accf[i], accu[i], W[b >> 2], gbase, 0x7650u | (b & 3));
  }
}

// ---- paired positions: acc2[k] = (accumulator of position 2k, of position 2k + 1) as f32x2
template <int IMM>
__device__ __forceinline__ void pair_lookup_add(unsigned long long& acc, uint32_t pword, uint32_t group_base,
                                                uint32_t sel) {
  uint32_t addr;
  unsigned long long v;
  asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(addr) : "r"(pword), "r"(group_base), "r"(sel));
  asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(IMM));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(acc), "l"(v));
}
template <int JJ>
__device__ __forceinline__ void pair_row_step(unsigned long long (&acc2)[RPT / 2], const uint32_t (&P)[NWIN],
                                              uint32_t gbase2) {
#pragma unroll
  for (int i = 0; i < RPT; i += 2) {
    const int b = i + JJ;  // window byte of (position i, motif row 4*j4 + JJ); its pair code covers bytes b, b + 1
    pair_lookup_add<JJ * 128>(acc2[i >> 1], P[b >> 2], gbase2, 0x7650u | (b & 3));
  }
}
// pair codes 8 s + 32 s' of 4 window symbols and their successors, from the pre-multiplied words
// (bytes 4 s): 2 * w + 8 * (w shifted down one byte); no carries between bytes while s, s' <= 3
__device__ __forceinline__ uint32_t pair_codes(uint32_t w, uint32_t w_next) {
  return 2u * w + 8u * __funnelshift_r(w, w_next, 8);
}

template <int KP, int MODE, bool U8>
__global__ void __launch_bounds__(THREADS, LM_SCORE_MINBLOCKS) score_kernel(const ScoreArgs a, uint64_t num_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool PAIRS = KP == 16 && !U8;  // the paired-position path exists for DNA f32
  const bool prefix = wants_prefix(KP, U8, a.n_motifs, a.m_pad);
  const SmemLayout L = smem_layout(KP, a.m_pad, prefix);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem);
  uint32_t* tilet = reinterpret_cast<uint32_t*>(smem + L.tilet_off);
  uint8_t* raw = smem + L.raw_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  unsigned long long* red = reinterpret_cast<unsigned long long*>(smem + L.red_off);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t tt_max = TILE_ROWS + a.m_pad;
  const uint64_t n_score_rows = a.row_end - a.row_begin;
  float2* table2 = reinterpret_cast<float2*>(smem + L.table2_off);
  float* table4 = reinterpret_cast<float*>(smem + L.table4_off);
  constexpr uint32_t SYM_MASK = (KP == 16) ? 0x0f0f0f0fu : 0x1f1f1f1fu;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // tile staging: rows [row_begin + tile*128, +tt) clamped to what is readable
  auto stage = [&](uint64_t tile, int buf) {
    uint64_t r0 = a.row_begin + tile * TILE_ROWS;
    uint64_t avail = a.avail_rows > r0 ? a.avail_rows - r0 : 0;
    uint32_t nrows = avail < tt_max ? (uint32_t)avail : tt_max;
    fence_proxy_async();
    mbar_expect_tx(&bars[buf], nrows * 32);
    if (nrows) tma_load_1d(raw + (size_t)buf * L.raw_bytes, a.seq + r0 * 32, nrows * 32, &bars[buf]);
  };

  // many motifs over few tiles (the seeding launch of a batch): the grid is num_tiles x motif_chunks
  // CTAs, CTA (tile, chunk) applies its chunk of the motifs to its one tile
  const uint32_t chunks = a.motif_chunks > 1 ? a.motif_chunks : 1;
  const uint32_t chunk = chunks > 1 ? (uint32_t)(blockIdx.x / num_tiles) : 0;
  const int q_begin = (int)((uint64_t)a.n_motifs * chunk / chunks), q_end = (int)((uint64_t)a.n_motifs * (chunk + 1) / chunks);
  const uint64_t step = chunks > 1 ? num_tiles : gridDim.x;
  uint64_t tile = chunks > 1 ? blockIdx.x % num_tiles : blockIdx.x;
  if (tid == 0 && tile < num_tiles) stage(tile, 0);

  // running per-column best of this thread's column (single-motif argmax):
  // score + row, updated with '>=' while rows increase => last row wins ties
  float best_f = 0.f;
  uint32_t best_u = 0, best_row = 0;
  bool best_any = false;

  uint32_t it = 0;
  for (; tile < num_tiles; tile += step, ++it) {
    const int buf = it & 1;
    const uint64_t next = tile + step;
    if (tid == 0 && next < num_tiles) stage(next, buf ^ 1);
    mbar_wait(&bars[buf], (it >> 1) & 1);

    const uint64_t tile_row0 = a.row_begin + tile * TILE_ROWS;
    uint8_t* rawb = raw + (size_t)buf * L.raw_bytes;
    {  // rows past the readable extent (end of the sequence only): default symbol
      uint64_t avail = a.avail_rows > tile_row0 ? a.avail_rows - tile_row0 : 0;
      if (avail < tt_max) {
        uint32_t w0 = (uint32_t)avail * 8, w1 = tt_max * 8;
        uint32_t fill = a.pad_sym * 0x01010101u;
        for (uint32_t w = w0 + tid; w < w1; w += THREADS) reinterpret_cast<uint32_t*>(rawb)[w] = fill;
        __syncthreads();
      }
    }
    // ---- transpose to column-major, symbols pre-multiplied by 4
    uint32_t seen = 0;  // OR of the symbols this thread moves (bit 2 of a byte: the default symbol N)
    {
      const uint32_t* raw32 = reinterpret_cast<const uint32_t*>(rawb);
      const uint32_t nblk = (tt_max / 4) * 8;
      for (uint32_t blk = tid; blk < nblk; blk += THREADS) {
        uint32_t cb = blk & 7, rb = blk >> 3;
        uint32_t w0 = raw32[(4 * rb + 0) * 8 + cb];
        uint32_t w1 = raw32[(4 * rb + 1) * 8 + cb];
        uint32_t w2 = raw32[(4 * rb + 2) * 8 + cb];
        uint32_t w3 = raw32[(4 * rb + 3) * 8 + cb];
        uint32_t u = prmt(w0, w1, 0x5140), v = prmt(w2, w3, 0x5140);
        uint32_t x = prmt(w0, w1, 0x7362), y = prmt(w2, w3, 0x7362);
        uint32_t t0 = prmt(u, v, 0x5410), t1 = prmt(u, v, 0x7632);
        uint32_t t2 = prmt(x, y, 0x5410), t3 = prmt(x, y, 0x7632);
        if (PAIRS) seen |= w0 | w1 | w2 | w3;
        uint32_t* dst = tilet + (4 * cb) * L.tsw + rb;
        dst[0 * L.tsw] = (t0 & SYM_MASK) << 2;
        dst[1 * L.tsw] = (t1 & SYM_MASK) << 2;
        dst[2 * L.tsw] = (t2 & SYM_MASK) << 2;
        dst[3 * L.tsw] = (t3 & SYM_MASK) << 2;
      }
    }
    // tiles holding an N anywhere (sequence ends, assembly gaps) take the one-look-up-per-update path
    bool use_pairs = false;
    if (PAIRS) use_pairs = !__syncthreads_or((seen & 0x04040404u) != 0) && a.n_motifs == 1;
    else __syncthreads();

    const uint32_t* colp = tilet + lane * L.tsw + warp * (RPT / 4);
    const uint64_t my_row0 = tile_row0 + (uint64_t)warp * RPT;  // absolute striped row of acc[0]

    for (int q = q_begin; q < q_end; ++q) {
      // ---- motif table -> shared memory
      int m_pad = a.m_pads ? a.m_pads[q] : a.m_pad;
      {
        const uint32_t* src = a.tables + (a.table_off ? a.table_off[q] : 0);
        if (it == 0 || a.n_motifs > 1) {
          if (a.n_motifs > 1) __syncthreads();  // previous motif's look-ups are done
          for (int w = tid; w < m_pad * KP; w += THREADS) table[w] = src[w];
          if (PAIRS && a.n_motifs == 1) {
            const float* tf = reinterpret_cast<const float*>(src);
            for (int e = tid; e < m_pad * 16; e += THREADS) {
              const int j = e >> 4, s0 = e & 3, s1 = (e >> 2) & 3;
              table2[e] = make_float2(tf[j * KP + s0], tf[j * KP + s1]);
            }
            // prefix table: entry (s0, s1, s2, s3) = ((( +0.0 + w0[s0]) + w1[s1]) + w2[s2]) + w3[s3], the exact value
            // every position's accumulator has after its first four motif rows (rows >= m are +0.0), replicated
            // once per lane so that a warp reads 32 different entries in one bank pass
            for (int idx = tid; prefix && idx < 256 * 32; idx += THREADS) {
              const int e = idx >> 5;
              float v = __fadd_rn(0.0f, tf[0 * KP + (e & 3)]);
              v = __fadd_rn(v, tf[1 * KP + ((e >> 2) & 3)]);
              v = __fadd_rn(v, tf[2 * KP + ((e >> 4) & 3)]);
              v = __fadd_rn(v, tf[3 * KP + ((e >> 6) & 3)]);
              table4[idx] = v;
            }
          }
          __syncthreads();
        }
      }
      float accf[RPT];
      uint32_t accu[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) { accf[i] = 0.0f; accu[i] = 0u; }
      uint32_t W[NWIN];
#pragma unroll
      for (int k = 0; k < NWIN - 1; ++k) W[k] = colp[k];
      if (PAIRS && use_pairs) {
        unsigned long long acc2[RPT / 2];
#pragma unroll
        for (int k = 0; k < RPT / 2; ++k) acc2[k] = 0ull;  // (+0.0, +0.0)
        uint32_t P[NWIN];
        uint32_t gbase2 = smem_u32(table2);
        // Motif rows 0..3 through the prefix table: ONE conflict-free LDS.32 per position instead of four
        // look-ups.  The sequential sum starts at +0.0, so its first four steps depend on four symbols only and
        // can be tabulated exactly; later steps depend on the running value and cannot.  The index comes from the
        // window bytes (4 s each): dp4a with (1, 4, 16, 64) = 4 * (s0 + 4 s1 + 16 s2 + 64 s3).
        int j4 = 0;
        if (prefix) {
          W[NWIN - 1] = colp[NWIN - 1];
          const uint32_t t4 = smem_u32(table4) + lane * 4;
          float v[RPT];
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const uint32_t w = (i & 3) ? __funnelshift_r(W[i >> 2], W[(i >> 2) + 1], 8 * (i & 3)) : W[i >> 2];
            const uint32_t code = __dp4a(w, 0x40100401u, 0u);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(t4 + code * 32u));
          }
#pragma unroll
          for (int k = 0; k < RPT / 2; ++k)
            acc2[k] = (unsigned long long)__float_as_uint(v[2 * k]) | ((unsigned long long)__float_as_uint(v[2 * k + 1]) << 32);
#pragma unroll
          for (int k = 0; k < NWIN - 1; ++k) W[k] = W[k + 1];
          gbase2 += 4 * 128;
          j4 = 1;
        }
#pragma unroll
        for (int k = 0; k < NWIN - 2; ++k) P[k] = pair_codes(W[k], W[k + 1]);
        for (; j4 < m_pad / 4; ++j4) {
          W[NWIN - 1] = colp[j4 + NWIN - 1];
          P[NWIN - 2] = pair_codes(W[NWIN - 2], W[NWIN - 1]);
          P[NWIN - 1] = pair_codes(W[NWIN - 1], 0u);  // only its first two codes (bytes 16, 17) are used
          pair_row_step<0>(acc2, P, gbase2);
          pair_row_step<1>(acc2, P, gbase2);
          pair_row_step<2>(acc2, P, gbase2);
          pair_row_step<3>(acc2, P, gbase2);
#pragma unroll
          for (int k = 0; k < NWIN - 1; ++k) { W[k] = W[k + 1]; P[k] = P[k + 1]; }
          gbase2 += 4 * 128;
        }
#pragma unroll
        for (int k = 0; k < RPT / 2; ++k) {
          accf[2 * k] = __uint_as_float((uint32_t)acc2[k]);
          accf[2 * k + 1] = __uint_as_float((uint32_t)(acc2[k] >> 32));
        }
      } else {
      uint32_t gbase = smem_u32(table);
      for (int j4 = 0; j4 < m_pad / 4; ++j4) {
        W[NWIN - 1] = colp[j4 + NWIN - 1];
        row_step<0, KP, U8>(accf, accu, W, gbase);
        row_step<1, KP, U8>(accf, accu, W, gbase);
        row_step<2, KP, U8>(accf, accu, W, gbase);
        row_step<3, KP, U8>(accf, accu, W, gbase);
#pragma unroll
        for (int k = 0; k < NWIN - 1; ++k) W[k] = W[k + 1];
        gbase += 4 * KP * 4;
      }
      }

      // ---- epilogue
      if (MODE & MODE_STORE) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          uint64_t row = my_row0 + i;
          if (row < a.row_end) {
            uint64_t o = (row - a.row_begin) * a.out_stride + lane;
            if (U8) reinterpret_cast<uint8_t*>(a.out)[o] = (uint8_t)min(accu[i], 255u);
            else reinterpret_cast<float*>(a.out)[o] = accf[i];
          }
        }
      }
      if (MODE & MODE_ARGMAX) {
        float bf = best_f;
        uint32_t bu = best_u, brow = best_row;
        bool any = best_any;
        if (a.n_motifs > 1) any = false;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint64_t row = my_row0 + i;
          if (row < a.row_end) {  // warp-uniform
            const uint32_t rel = (uint32_t)(row - a.row_begin);
            if (U8) {
              const uint32_t s = min(accu[i], 255u);
              if (!any || s >= bu) { bu = s; brow = rel; }
            } else {
              if (!any || accf[i] >= bf) { bf = accf[i]; brow = rel; }
            }
            any = true;
          }
        }
        if (my_row0 == a.row_begin && lane == 0 && a.row_begin < a.row_end)
          a.first_cell[(size_t)q * a.first_cell_stride] = U8 ? min(accu[0], 255u) : __float_as_uint(accf[0]);
        if (a.n_motifs > 1) {
          // u8: +1 so that a valid candidate never encodes to key 0 ("none")
          const uint32_t sk = U8 ? bu + 1u : f32_to_key(bf);
          red[warp * 32 + lane] = any ? (((unsigned long long)sk << 32) | brow) : 0ull;
          __syncthreads();
          if (warp == 0) {
            unsigned long long m = red[lane];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) { unsigned long long o = red[w * 32 + lane]; m = o > m ? o : m; }
            if (m) atomicMax(&a.colkeys[(size_t)q * a.colkeys_stride + lane], m);
          }
        } else {
          best_f = bf; best_u = bu; best_row = brow; best_any = any;
        }
      }
      if (MODE & MODE_THRESHOLD) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          uint64_t row = my_row0 + i;
          bool hit = row < a.row_end;
          float sc;
          if (U8) { uint32_t s = min(accu[i], 255u); hit = hit && s >= a.thr_bits; sc = (float)s; }
          else { hit = hit && accf[i] >= __uint_as_float(a.thr_bits); sc = accf[i]; }
          unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.count, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
              unsigned long long slot = base + __popc(mask & ((1u << lane) - 1));
              if (slot < a.cap) {
                lm_hit h;
                h.row = row - a.row_begin + a.row_add;
                h.col = lane;
                h.position = (uint64_t)lane * (a.pos_rows ? a.pos_rows : n_score_rows) + h.row;
                h.score = sc;
                a.hits[slot] = h;
              }
            }
          }
        }
      }
    }
    __syncthreads();  // tile buffers are reused by the next iteration
  }

  if ((MODE & MODE_ARGMAX) && a.n_motifs == 1) {
    const uint32_t sk = U8 ? best_u + 1u : f32_to_key(best_f);
    red[warp * 32 + lane] = best_any ? (((unsigned long long)sk << 32) | best_row) : 0ull;
    __syncthreads();
    if (warp == 0) {
      unsigned long long m = red[lane];
#pragma unroll
      for (int w = 1; w < WARPS; ++w) { unsigned long long o = red[w * 32 + lane]; m = o > m ? o : m; }
      if (m) atomicMax(&a.colkeys[lane], m);
    }
  }
}

template <int KP, bool U8>
cudaError_t launch_mode(int mode, const ScoreArgs& a, int sm_count, cudaStream_t stream) {
  const uint64_t n_rows = a.row_end - a.row_begin;
  const uint64_t num_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  if (num_tiles == 0) return cudaSuccess;
  const SmemLayout L = smem_layout(KP, a.m_pad, wants_prefix(KP, U8, a.n_motifs, a.m_pad));
  if (L.total > SMEM_LIMIT) return cudaErrorInvalidValue;
  void (*kern)(const ScoreArgs, uint64_t) = nullptr;
  switch (mode) {
#define LM_CASE(M) case M: kern = score_kernel<KP, M, U8>; break;
    LM_CASE(1) LM_CASE(2) LM_CASE(3) LM_CASE(4) LM_CASE(5) LM_CASE(6) LM_CASE(7)
#undef LM_CASE
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, L.total);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = (uint64_t)sm_count * per_sm;
  ScoreArgs b = a;
  b.motif_chunks = 1;
  if (a.n_motifs > 1 && num_tiles * 2 <= grid) {  // fewer tiles than CTAs: split the motifs as well
    b.motif_chunks = (uint32_t)std::min<uint64_t>(grid / num_tiles, (uint64_t)a.n_motifs);
    grid = num_tiles * b.motif_chunks;
  } else if (grid > num_tiles) {
    grid = num_tiles;
  }
  kern<<<(unsigned)grid, THREADS, L.total, stream>>>(b, num_tiles);
  return cudaGetLastError();
}

}  // namespace

namespace {
template <int KP, bool U8>
cudaError_t preload_modes() {
  cudaFuncAttributes at;
  cudaError_t e = cudaSuccess;
#define LM_CASE(M) if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, score_kernel<KP, M, U8>);
  LM_CASE(1) LM_CASE(2) LM_CASE(3) LM_CASE(4) LM_CASE(5) LM_CASE(6) LM_CASE(7)
#undef LM_CASE
  return e;
}
}  // namespace

// load every instance now instead of at its first launch (see lmi::preload_kernels in capi.cu)
cudaError_t preload_score_kernels() {
  cudaError_t e = preload_modes<16, false>();
  if (e == cudaSuccess) e = preload_modes<16, true>();
  if (e == cudaSuccess) e = preload_modes<32, false>();
  if (e == cudaSuccess) e = preload_modes<32, true>();
  return e;
}

cudaError_t launch_score(int alphabet, int mode, bool is_u8, const ScoreArgs& a, int sm_count,
                         cudaStream_t stream) {
  if (alphabet == LM_DNA)
    return is_u8 ? launch_mode<16, true>(mode, a, sm_count, stream)
                 : launch_mode<16, false>(mode, a, sm_count, stream);
  return is_u8 ? launch_mode<32, true>(mode, a, sm_count, stream)
               : launch_mode<32, false>(mode, a, sm_count, stream);
}

}  // namespace lm
