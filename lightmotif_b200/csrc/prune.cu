// prune.cu -- exact filter + recheck scan (DNA), see prune.cuh.
//
// Kernel.  Every WARP is autonomous: it pulls chunks of consecutive warp tiles
// (4 * PL = 84 striped rows x 32 columns = 2688 positions) from a global
// counter, stages each tile (+ halo) from HBM into its own double-buffered
// shared-memory slot with one bulk-TMA copy (own mbarriers), filters it,
// rechecks its own candidates.  There is no CTA-wide barrier in steady state.
//
//  filter   lane = (column quad q = lane & 7, row group g = lane >> 3); a thread
//           walks PL (21) consecutive rows of its 4 columns.  One LDS.32 per row
//           yields 4 symbols (4 adjacent columns = 4 independent positions);
//           IMAD + SHF turn them into a PRMT selector that is reused by all R
//           filter rows:   acc = bias + sum_k PRMT(table[k], selector[i + k])
//           (byte lanes, no carries across lanes by construction; the adds are
//           issued as IMADs so that they run on the FMA pipe beside the PRMTs
//           on the ALU pipe).  A position passes iff bit 7 of its byte is set,
//           i.e. iff its integer upper bound reaches the current threshold.
//  recheck  passing positions go to the warp's shared-memory queue; after the
//           filter pass the warp rescans them EXACTLY (sequential f32 sum over
//           all M rows in the reference's order, from the tile already in
//           shared memory) and records those that really reach the threshold
//           (atomicMax on the per-column winner keys / hit list append).
//  fallback a tile whose queue overflows (low-complexity sequence, very low
//           threshold) is rescored exactly position by position.
//
// The bound:  UB(p) = c0 + delta * sum_q(p) >= real-valued score(p), and the
// fp32 sequential sum differs from the real sum by at most eps, so
//   fp32score(p) >= T  ==>  sum_q(p) >= ceil((T - eps - c0) / delta)  (- 1 safety)
#include <algorithm>
#include <cmath>
#include <vector>

#include "prune.cuh"

namespace lm {

namespace {

constexpr int PW = 8;                   // consumer (filter) warps per CTA
constexpr int PT = (PW + 1) * 32;       // + one producer warp
#ifndef LM_PRUNE_EXP
#define LM_PRUNE_EXP 0
#endif
#ifndef LM_PRUNE_NSTAGE
#define LM_PRUNE_NSTAGE 3
#endif
constexpr int NSTAGE = LM_PRUNE_NSTAGE; // tile buffers in flight per CTA
#ifndef LM_PRUNE_PL
#define LM_PRUNE_PL 21
#endif
#ifndef LM_PRUNE_MINBLOCKS
#define LM_PRUNE_MINBLOCKS 3
#endif
constexpr int PL = LM_PRUNE_PL;        // rows per thread per tile (odd => conflict-free LDS)
constexpr int WARP_ROWS = 4 * PL;
constexpr int PTILE_ROWS = PW * WARP_ROWS;
constexpr int QCAP = 64;                // queue entries (quads) per warp and tile; more => exact fallback
static_assert(PL % 2 == 1, "PL must be odd (bank-conflict-free row groups)");

struct PSmem {
  uint32_t table_off, ctl_off, queue_off, raw_off, raw_bytes, total;
};
// pipeline control block: full[s] (TMA landed + descriptor valid) / empty[s] (all consumer warps done)
struct Ctl {
  uint64_t full[NSTAGE], empty[NSTAGE];
  struct { unsigned long long tile; uint32_t bias4; float thr; } desc[NSTAGE];
};
__host__ __device__ inline PSmem psmem(int m) {
  PSmem s;
  uint32_t tt = PTILE_ROWS + ((m + 3) & ~3);
  s.table_off = 0;
  s.ctl_off = ((uint32_t)((m + 3) & ~3) * 8 * 4 + 127) & ~127u;
  s.queue_off = (s.ctl_off + (uint32_t)sizeof(Ctl) + 127) & ~127u;
  s.raw_off = s.queue_off + PW * QCAP * 4;
  s.raw_bytes = (tt * 32 + 127) & ~127u;
  s.total = s.raw_off + NSTAGE * s.raw_bytes;
  return s;
}

// raw PRMT (the __byte_perm intrinsic masks the selector with 0x7777 first)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// acc + p as `p * one + acc`: an IMAD on the (otherwise idle) FMA pipe instead of an IADD3 on the
// ALU pipe that the PRMT look-ups saturate.  `one` is a kernel parameter so ptxas cannot fold it.
#ifndef LM_PRUNE_IMAD
#define LM_PRUNE_IMAD 1
#endif
__device__ __forceinline__ uint32_t add_fma_pipe(uint32_t p, uint32_t one, uint32_t acc) {
#if LM_PRUNE_IMAD
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(p), "r"(one), "r"(acc));
  return d;
#else
  return acc + p;
#endif
}

// PRMT selector of a quad (4 symbols, one per byte, each <= 7):
//   x * 0x00100001 = x + (x << 20) puts s2@16, s0@20, s3@24, s1@28 (no overlaps),
// so the top half holds the four symbols as nibbles in the order (2, 0, 3, 1):
// one IMAD (FMA pipe) + one shift instead of three ALU ops.  Byte lane b of
// every look-up therefore belongs to column LANE_COL[b] of the quad.
__device__ __forceinline__ uint32_t make_sel(uint32_t x) { return (x * 0x00100001u) >> 16; }
__device__ __constant__ const uint8_t LANE_COL[4] = {2, 0, 3, 1};

template <bool THR>
__device__ __forceinline__ void record(const PruneArgs& a, float s, uint32_t rel_row, uint32_t col) {
  if (THR) {
    unsigned long long slot = atomicAdd(a.count, 1ull);
    if (slot < a.cap) {
      lm_hit h;
      h.row = rel_row;
      h.col = col;
      h.position = (uint64_t)col * (a.row_end - a.row_begin) + rel_row;
      h.score = s;
      a.hits[slot] = h;
    }
  } else {
    unsigned long long key = ((unsigned long long)f32_to_key(s) << 32) | rel_row;
    atomicMax(&a.colkeys[col], key);
  }
}

// exact sequential f32 score of the window starting at (row_in_tile, col).  The table is padded
// to a multiple of 4 rows with +0.0 (x + 0.0 == x: the chain starts at +0.0 and never yields -0.0).
__device__ __forceinline__ float exact_score(const uint8_t* rawb, const float* tab, int m_pad,
                                             uint32_t row_in_tile, uint32_t col) {
  float s = 0.0f;
  const uint8_t* p = rawb + row_in_tile * 32 + col;
  for (int j = 0; j < m_pad; j += 4) {
    const uint32_t s0 = p[0], s1 = p[32], s2 = p[64], s3 = p[96];
    const float w0 = tab[s0], w1 = tab[8 + s1], w2 = tab[16 + s2], w3 = tab[24 + s3];
    s = __fadd_rn(s, w0);
    s = __fadd_rn(s, w1);
    s = __fadd_rn(s, w2);
    s = __fadd_rn(s, w3);
    p += 128;
    tab += 32;
  }
  return s;
}

template <int R, bool THR>
__global__ void __launch_bounds__(PT, LM_PRUNE_MINBLOCKS) prune_kernel(const PruneArgs a, uint64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.plan.m;
  const PSmem S = psmem(m);
  const uint32_t tt = PTILE_ROWS + ((m + 3) & ~3);  // rows staged per tile (tile + halo)
  float* tab = reinterpret_cast<float*>(smem + S.table_off);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + S.ctl_off);
  uint8_t* raw = smem + S.raw_off;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], PW);
    }
    mbar_fence_init();
  }
  const int m_pad = (m + 3) & ~3;
  for (int i = tid; i < m_pad * 8; i += PT) tab[i] = a.table[i];  // [m_pad][8], rows >= m are +0.0
  __syncthreads();  // the only CTA-wide barrier: from here on warps meet through mbarriers only

  if (warp == PW) {
    // ===================== producer warp: claims tiles, stages them with bulk TMA, publishes the
    // threshold each tile is to be filtered at.  Tiles are claimed one at a time from a global
    // counter, so the CTAs of the whole grid sweep the sequence as one dense window.
    uint32_t thr_key = 0xffffffffu, bias4 = 0;
    float thr = THR ? a.thr : -INFINITY;
    for (uint32_t it = 0;; ++it) {
      const int s = it % NSTAGE;
      // claim the tile and fetch the running maximum first: both round trips overlap the wait
      unsigned long long tile = 0;
      if (lane == 0) tile = atomicAdd(&a.stats[2], 1ull);
      uint32_t sk = 0;
      if (!THR) sk = (uint32_t)(__ldcg(&a.colkeys[lane]) >> 32);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (!THR) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sk = max(sk, __shfl_xor_sync(0xffffffffu, sk, o));
      }
      if (it >= NSTAGE) mbar_wait_hint(&ctl->empty[s], ((it / NSTAGE) - 1) & 1, 4000);
      // ---- current threshold -> byte bias (recomputed only when the threshold moved)
      if (THR ? it == 0 : sk != thr_key) {
        const float t = THR ? a.thr : (sk ? key_to_f32(sk) : -INFINITY);
        // sum_q >= qthr is implied by fp32score >= t (see header); -1 guards the division
        double need = ((double)t - a.plan.eps - a.plan.c0) / a.plan.delta;
        int bias;
        if (!(need > 1.0)) bias = 128;                               // everything passes
        else if (need > (double)a.plan.smax + 2.0) bias = 0;         // nothing can pass
        else {
          int qthr = (int)ceil(need) - 1;
          qthr = qthr < 1 ? 1 : (qthr > 128 ? 128 : qthr);
          bias = 128 - qthr;
        }
        bias4 = (uint32_t)bias * 0x01010101u;
        thr = t;
        thr_key = sk;
      }
      if (lane == 0) {
        ctl->desc[s].tile = tile;
        ctl->desc[s].bias4 = bias4;
        ctl->desc[s].thr = thr;
        if (tile < num_tiles) {
#if LM_PRUNE_EXP == 2  // experiment: no DRAM traffic (every tile re-reads the first few tiles from L2)
          uint64_t r0 = a.row_first + (tile & 3) * PTILE_ROWS;
#else
          uint64_t r0 = a.row_first + tile * PTILE_ROWS;
#endif
          uint64_t avail = a.avail_rows > r0 ? a.avail_rows - r0 : 0;
          uint32_t nrows = avail < tt ? (uint32_t)avail : tt;
          fence_proxy_async();
          mbar_expect_tx(&ctl->full[s], nrows * 32);
          if (nrows) tma_load_1d(raw + (size_t)s * S.raw_bytes, a.seq + r0 * 32, nrows * 32, &ctl->full[s]);
        } else {
          mbar_arrive(&ctl->full[s]);  // end marker: consumers read desc.tile >= num_tiles and leave
        }
      }
      if (tile >= num_tiles) break;
    }
    return;
  }

  // ===================== consumer warps: warp w filters rows [w * WARP_ROWS, (w + 1) * WARP_ROWS) of
  // every tile and rechecks its own candidates
  const int q = lane & 7, g = lane >> 3;
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem + S.queue_off) + warp * QCAP;
  const uint32_t lanemask_lt = (1u << lane) - 1u;

  // filter tables in per-thread registers.  They are bounced through shared
  // memory on purpose: PRMT cannot take its table operand from a uniform
  // register, and ptxas would otherwise keep the (uniform) kernel-parameter
  // values in URs and re-materialise them with a UR->R move before every look-up.
  uint32_t tlo[R], thi[R];
  {
    if (lane < R) { queue[lane] = a.plan.tab_lo[lane]; queue[16 + lane] = a.plan.tab_hi[lane]; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
      tlo[k] = *reinterpret_cast<volatile uint32_t*>(&queue[k]);
      thi[k] = *reinterpret_cast<volatile uint32_t*>(&queue[16 + k]);
    }
    __syncwarp();
  }
  const uint32_t one = a.one;
  unsigned long long n_cand = 0, n_over = 0;

  for (uint32_t it = 0;; ++it) {
    const int s = it % NSTAGE;
    mbar_wait_hint(&ctl->full[s], (it / NSTAGE) & 1, 2000);
    const uint64_t tile = ctl->desc[s].tile;
    if (tile >= num_tiles) break;
    const uint32_t bias4 = ctl->desc[s].bias4;
    const float thr = ctl->desc[s].thr;
    const uint32_t wrow0 = warp * WARP_ROWS;                       // this warp's first row in the tile
    const uint64_t tile_row0 = a.row_first + tile * PTILE_ROWS;
    uint8_t* rawb = raw + (size_t)s * S.raw_bytes;
    {
      // rows past the readable extent (end of the sequence only): default symbol.  Every warp fills
      // the rows IT reads; overlapping halo rows get the same value from both neighbours.
      uint64_t avail = a.avail_rows > tile_row0 ? a.avail_rows - tile_row0 : 0;
      const uint32_t wend = wrow0 + WARP_ROWS + ((m + 3) & ~3);
      if (avail < wend) {
        uint32_t w0 = (avail > wrow0 ? (uint32_t)avail : wrow0) * 8, w1 = wend * 8;
        uint32_t fill = a.pad_sym * 0x01010101u;
        for (uint32_t w = w0 + lane; w < w1; w += 32) reinterpret_cast<uint32_t*>(rawb)[w] = fill;
        __syncwarp();
      }
    }
    const uint64_t left = a.row_end - tile_row0;  // > 0 for every claimed tile
    const uint32_t tile_valid = left < PTILE_ROWS ? (uint32_t)left : PTILE_ROWS;
    const uint32_t rows_valid = tile_valid > wrow0 ? min(tile_valid - wrow0, (uint32_t)WARP_ROWS) : 0;

#if LM_PRUNE_EXP == 1  // experiment: stream the tiles, no filtering
    if (rows_valid && bias4 == 0x12345678u) {
#else
    if (rows_valid) {
#endif
      // ---- filter
      const uint32_t row_base = wrow0 + g * PL;
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(rawb) + (row_base + a.plan.j0) * 8 + q;
      uint32_t sel[PL + R - 1];
#pragma unroll
      for (int i = 0; i < R - 1; ++i) sel[i] = make_sel(rp[i * 8]);
      // Candidates are appended to the warp's queue with a ballot (the queue length lives in a
      // warp-uniform register: no shared-memory atomics, and the branch is warp-uniform).  One
      // entry = a PAIR of consecutive quad rows: first row, column quad, 2 x 4 byte-lane flags.
      uint32_t nq = 0;
      auto push2 = [&](uint32_t lo, uint32_t hi, int i0) {
        const bool mine = ((lo | hi) & 0x80808080u) != 0;  // one LOP3 with a predicate result
        if (__any_sync(0xffffffffu, mine)) {
          const uint32_t votes = __ballot_sync(0xffffffffu, mine);
          if (mine) {
            const uint32_t mlo = lo & 0x80808080u, mhi = hi & 0x80808080u;
            const uint32_t slot = nq + __popc(votes & lanemask_lt);
            const uint32_t flags = (((mlo >> 7) * 0x10204080u) >> 28) | ((((mhi >> 7) * 0x10204080u) >> 28) << 4);
            if (slot < QCAP) queue[slot] = ((row_base + i0) << 11) | (q << 8) | flags;
          }
          nq += __popc(votes);
        }
      };
      uint32_t prev = 0;
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        sel[i + R - 1] = make_sel(rp[(i + R - 1) * 8]);
        uint32_t acc = bias4;
#pragma unroll
        for (int k = 0; k < R; ++k) acc = add_fma_pipe(prmt_raw(tlo[k], thi[k], sel[i + k]), one, acc);
        if (i & 1) push2(prev, acc, i - 1);          // one test per PAIR of quad rows
        else if (i == PL - 1) push2(acc, 0u, i);
        prev = acc;
      }
      __syncwarp();  // queue complete

      // ---- exact recheck of the candidates (or of the warp's whole slice on overflow)
      if (nq) {
        if (nq <= QCAP) {
          for (uint32_t k = lane; k < nq * 8; k += 32) {
            const uint32_t code = queue[k >> 3], b = k & 7;
            if (!((code >> b) & 1u)) continue;
            const uint32_t row = (code >> 11) + (b >> 2), col = ((code >> 8) & 7u) * 4 + LANE_COL[b & 3];
            if (row >= tile_valid) continue;
            float sc = exact_score(rawb, tab, m_pad, row, col);
            if (sc >= thr) record<THR>(a, sc, (uint32_t)(tile_row0 + row - a.row_begin), col);
          }
          n_cand += nq;
        } else {
          for (uint32_t p = lane; p < rows_valid * 32; p += 32) {
            uint32_t row = wrow0 + (p >> 5), col = p & 31;
            float sc = exact_score(rawb, tab, m_pad, row, col);
            if (sc >= thr) record<THR>(a, sc, (uint32_t)(tile_row0 + row - a.row_begin), col);
          }
          n_over += 1;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->empty[s]);  // this warp is done with the stage
  }
  if (lane == 0 && a.stats) {
    if (n_cand) atomicAdd(&a.stats[0], n_cand);
    if (n_over) atomicAdd(&a.stats[1], n_over);
  }
}

template <bool THR>
cudaError_t launch_r(const PruneArgs& a, int sm_count, cudaStream_t stream) {
  if (a.row_end <= a.row_first) return cudaSuccess;
  const uint64_t num_tiles = (a.row_end - a.row_first + PTILE_ROWS - 1) / PTILE_ROWS;
  const PSmem S = psmem(a.plan.m);
  void (*kern)(const PruneArgs, uint64_t) = nullptr;
  switch (a.plan.r) {
    case 4: kern = prune_kernel<4, THR>; break;
    case 6: kern = prune_kernel<6, THR>; break;
    case 8: kern = prune_kernel<8, THR>; break;
    case 10: kern = prune_kernel<10, THR>; break;
    case 12: kern = prune_kernel<12, THR>; break;
    case 16: kern = prune_kernel<16, THR>; break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT, S.total);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = std::min<uint64_t>((uint64_t)sm_count * per_sm, num_tiles);
  kern<<<(unsigned)grid, PT, S.total, stream>>>(a, num_tiles);
  return cudaGetLastError();
}

// ------------------------------------------------------------ host planning

constexpr double CLAMP_FLOOR = 40.0;  // -inf / very negative weights are treated as rowmax - 40 in the DPs

struct Rows {
  int m;
  std::vector<double> w;       // [m][5]
  std::vector<double> rowmax;  // over all 5 letters
};

bool load_rows(const float* pssm, size_t m, size_t stride, Rows& r) {
  r.m = (int)m;
  r.w.resize(m * 5);
  r.rowmax.resize(m);
  for (size_t j = 0; j < m; ++j) {
    double mx = -INFINITY;
    for (int a = 0; a < 5; ++a) {
      float x = pssm[j * stride + a];
      if (std::isnan(x) || x == INFINITY) return false;
      r.w[j * 5 + a] = x;
      if (x > mx) mx = x;
    }
    if (!std::isfinite(mx)) return false;
    r.rowmax[j] = mx;
  }
  return true;
}

}  // namespace

double estimate_max_score(const float* pssm, size_t m, size_t stride, double n_positions) {
  Rows r;
  if (!load_rows(pssm, m, stride, r)) return -INFINITY;
  // distribution of the score under uniform ACGT on a 1/32-bit grid (weights clamped from below)
  const double res = 1.0 / 32.0;
  std::vector<int> lo(m), hi(m);
  std::vector<int> qw(m * 4);
  int span = 0;
  for (size_t j = 0; j < m; ++j) {
    double mx = -INFINITY;
    for (int a = 0; a < 4; ++a) mx = std::max(mx, r.w[j * 5 + a]);
    if (!std::isfinite(mx)) return -INFINITY;
    int mn = 0;
    for (int a = 0; a < 4; ++a) {
      double x = std::max(r.w[j * 5 + a], mx - CLAMP_FLOOR);
      qw[j * 4 + a] = (int)std::floor((x - mx) / res);  // <= 0
      mn = std::min(mn, qw[j * 4 + a]);
    }
    span += -mn;
  }
  std::vector<double> d(span + 1, 0.0), nd(span + 1);
  d[0] = 1.0;  // index = deficit below the maximum, in grid units
  double maxsum = 0;
  for (size_t j = 0; j < m; ++j) {
    double mx = -INFINITY;
    for (int a = 0; a < 4; ++a) mx = std::max(mx, r.w[j * 5 + a]);
    maxsum += mx;
    std::fill(nd.begin(), nd.end(), 0.0);
    for (int a = 0; a < 4; ++a) {
      int s = -qw[j * 4 + a];
      for (int i = 0; i + s <= span; ++i) nd[i + s] += 0.25 * d[i];
    }
    d.swap(nd);
  }
  double cum = 0;
  for (int i = 0; i <= span; ++i) {
    cum += d[i];
    if (cum * n_positions >= 1.0) return maxsum - (i + 1) * res;
  }
  return maxsum - span * res;
}

PrunePlan plan_prune(const float* pssm, size_t m, size_t stride, double t_plan, bool t_is_estimate,
                     double n_positions, int force_r) {
  PrunePlan best;
  Rows r;
  if (m < 4 || m > PRUNE_MAX_M || !load_rows(pssm, m, stride, r)) return best;
  (void)n_positions;
  if (t_is_estimate) t_plan -= 1.0;  // the running maximum spends most of the scan just below the final one
  double abs_sum = 0;                // sum_j max_a |w| over finite entries, for the fp32 error bound
  double max_total = 0;
  for (size_t j = 0; j < m; ++j) {
    double am = 0;
    for (int a = 0; a < 5; ++a)
      if (std::isfinite(r.w[j * 5 + a])) am = std::max(am, std::fabs(r.w[j * 5 + a]));
    abs_sum += am;
    max_total += r.rowmax[j];
  }
  const double eps = 2.0 * (double)m * std::ldexp(1.0, -24) * abs_sum + 1e-30;
  static const int RS[] = {4, 6, 8, 10, 12, 16};
  static const double CS[] = {4.0, 6.0, 8.0, 12.0, CLAMP_FLOOR};
  double best_cost = 1e300;
  // cost model calibrated on B200 (profiles/README.md): per position, a filter of R rows costs
  // ~(1.5 R + 15.6) / 4 issue slots, a queued candidate ~870 (divergent recheck), exact scoring ~55
  const double EXACT_COST = 55.0, RECHECK_COST = 870.0;
  for (int R : RS) {
    if (R > (int)m || (force_r && R != force_r)) continue;
    for (int j0 = 0; j0 + R <= (int)m; ++j0) {
      double restmax = max_total;
      for (int k = 0; k < R; ++k) restmax -= r.rowmax[j0 + k];
      for (double C : CS) {
        // clamped weights, per-row base, step
        double wc[16][5], base[16], range_sum = 0, base_sum = 0;
        for (int k = 0; k < R; ++k) {
          const double mx = r.rowmax[j0 + k];
          double mn = INFINITY;
          for (int a = 0; a < 5; ++a) {
            wc[k][a] = std::max(r.w[(j0 + k) * 5 + a], mx - C);
            mn = std::min(mn, wc[k][a]);
          }
          base[k] = mn;
          base_sum += mn;
          range_sum += mx - mn;
        }
        if (!(range_sum > 0)) continue;
        const double delta = range_sum / (127.0 - R);
        int q[16][5], smax = 0;
        for (int k = 0; k < R; ++k) {
          int mxq = 0;
          for (int a = 0; a < 5; ++a) {
            q[k][a] = (int)std::ceil((wc[k][a] - base[k]) / delta - 1e-9);
            if (q[k][a] < 0) q[k][a] = 0;
            // make sure the bound holds despite the 1e-9 slack
            while (base[k] + q[k][a] * delta < wc[k][a]) q[k][a]++;
            mxq = std::max(mxq, q[k][a]);
          }
          smax += mxq;
        }
        if (smax > 127) continue;
        const double c0 = base_sum + restmax;
        // pass rate under uniform ACGT at t_plan
        double need = (t_plan - eps - c0) / delta;
        int qthr = need > 1.0 ? (int)std::ceil(need) - 1 : 0;
        double pass;
        if (qthr <= 0) pass = 1.0;
        else if (qthr > smax) pass = 0.0;
        else {
          double d[128] = {0}, nd[128];
          d[0] = 1.0;
          for (int k = 0; k < R; ++k) {
            std::fill(nd, nd + 128, 0.0);
            for (int a = 0; a < 4; ++a)
              for (int i = 0; i + q[k][a] < 128; ++i) nd[i + q[k][a]] += 0.25 * d[i];
            std::copy(nd, nd + 128, d);
          }
          pass = 0;
          for (int i = qthr; i < 128; ++i) pass += d[i];
        }
        const double cost = (1.5 * R + 15.6) / 4.0 + pass * RECHECK_COST;
        if (cost < best_cost) {
          best_cost = cost;
          best.m = (int)m;
          best.r = R;
          best.j0 = j0;
          best.smax = smax;
          best.c0 = c0;
          best.delta = delta;
          best.eps = eps;
          best.est_pass_rate = pass;
          for (int k = 0; k < 16; ++k) best.tab_lo[k] = best.tab_hi[k] = 0;
          for (int k = 0; k < R; ++k) {
            best.tab_lo[k] = (uint32_t)q[k][0] | ((uint32_t)q[k][1] << 8) | ((uint32_t)q[k][2] << 16) |
                             ((uint32_t)q[k][3] << 24);
            best.tab_hi[k] = (uint32_t)q[k][4];
          }
        }
      }
    }
  }
  best.ok = best_cost < EXACT_COST && best.r > 0;
  return best;
}


cudaError_t launch_prune(bool threshold_mode, const PruneArgs& a, int sm_count, cudaStream_t stream) {
  return threshold_mode ? launch_r<true>(a, sm_count, stream) : launch_r<false>(a, sm_count, stream);
}

}  // namespace lm
