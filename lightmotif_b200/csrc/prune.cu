// prune.cu -- exact filter + recheck scan (DNA), see prune.cuh.
//
// Kernel (one CTA = 8 warps, tile = 672 striped rows x 32 columns = 21 504
// positions, staged from HBM with one bulk-TMA copy, double-buffered):
//
//  filter   lane = (column quad q = lane & 7, row group g = lane >> 3); a thread
//           walks PL (21) consecutive rows of its 4 columns.  One LDS.32 per row
//           yields 4 symbols (4 adjacent columns = 4 independent positions);
//           3 ALU ops turn them into a PRMT selector that is reused by all R
//           filter rows:   acc = bias + sum_k PRMT(table[k], selector[i + k])
//           (byte lanes, no carries across lanes by construction).  A
//           position passes iff bit 7 of its byte is set, i.e. iff its integer
//           upper bound reaches the current threshold.
//  recheck  passing positions go to a shared-memory queue; after the filter
//           pass the CTA rescans them EXACTLY (sequential f32 sum over all M
//           rows in the reference's order, from the tile already in shared
//           memory) and records those that really reach the threshold
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

constexpr int PW = 8;
constexpr int PT = PW * 32;
#ifndef LM_PRUNE_PL
#define LM_PRUNE_PL 21
#endif
#ifndef LM_PRUNE_MINBLOCKS
#define LM_PRUNE_MINBLOCKS 3
#endif
constexpr int PL = LM_PRUNE_PL;        // rows per thread per tile (odd => conflict-free LDS)
constexpr int WARP_ROWS = 4 * PL;
constexpr int PTILE_ROWS = PW * WARP_ROWS;
constexpr int QCAP = 2048;
static_assert(PL % 2 == 1, "PL must be odd (bank-conflict-free row groups)");

struct PSmem {
  uint32_t raw_off, raw_bytes, queue_off, table_off, misc_off, total;
};
__host__ __device__ inline PSmem psmem(int m) {
  PSmem s;
  uint32_t tt = PTILE_ROWS + ((m + 3) & ~3);
  s.raw_off = 0;
  s.raw_bytes = tt * 32;
  s.queue_off = 2 * s.raw_bytes;
  s.table_off = s.queue_off + QCAP * 4;
  s.misc_off = s.table_off + (uint32_t)m * 8 * 4;
  s.total = s.misc_off + 64;
  return s;
}

struct Misc {
  uint64_t bars[2];
  unsigned long long next_tile[2];  // dynamically scheduled tile staged in each buffer
  uint32_t qcount[2];               // double-buffered queue length (saves one barrier per tile)
  uint32_t bias4;
  float thr;
  uint32_t thr_key;                 // threshold key the bias was computed for
  uint32_t pad;
};

// raw PRMT (the __byte_perm intrinsic masks the selector with 0x7777 first)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
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

// exact sequential f32 score of the window starting at (row_in_tile, col)
__device__ __forceinline__ float exact_score(const uint8_t* rawb, const float* tab, int m,
                                             uint32_t row_in_tile, uint32_t col) {
  float s = 0.0f;
  const uint8_t* p = rawb + row_in_tile * 32 + col;
  for (int j = 0; j < m; ++j) s = __fadd_rn(s, tab[j * 8 + p[j * 32]]);
  return s;
}

#ifdef LM_PRUNE_DEBUG
__device__ unsigned long long g_dbg[2048][4];
#endif

template <int R, bool THR>
__global__ void __launch_bounds__(PT, LM_PRUNE_MINBLOCKS) prune_kernel(const PruneArgs a, uint64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const PSmem S = psmem(a.plan.m);
  uint8_t* raw = smem + S.raw_off;
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem + S.queue_off);
  float* tab = reinterpret_cast<float*>(smem + S.table_off);
  Misc* misc = reinterpret_cast<Misc*>(smem + S.misc_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 7, g = lane >> 3;
  const int m = a.plan.m;
  const uint32_t tt = PTILE_ROWS + ((m + 3) & ~3);

  if (tid == 0) {
    mbar_init(&misc->bars[0], 1);
    mbar_init(&misc->bars[1], 1);
    mbar_fence_init();
    misc->qcount[0] = misc->qcount[1] = 0;
  }
  for (int i = tid; i < m * 8; i += PT) tab[i] = a.table[i];
  __syncthreads();

  auto stage = [&](uint64_t tile, int buf) {
    uint64_t r0 = a.row_first + tile * PTILE_ROWS;
    uint64_t avail = a.avail_rows > r0 ? a.avail_rows - r0 : 0;
    uint32_t nrows = avail < tt ? (uint32_t)avail : tt;
    fence_proxy_async();
    mbar_expect_tx(&misc->bars[buf], nrows * 32);
    if (nrows) tma_load_1d(raw + (size_t)buf * S.raw_bytes, a.seq + r0 * 32, nrows * 32, &misc->bars[buf]);
  };

  // filter tables in per-thread registers.  They are bounced through shared
  // memory on purpose: PRMT cannot take its table operand from a uniform
  // register, and ptxas would otherwise keep the (uniform) kernel-parameter
  // values in URs and re-materialise them with a UR->R move before every look-up.
  uint32_t tlo[R], thi[R];
  {
    uint32_t* tq = queue;  // the queue is unused until the first filter pass
    if (tid < R) { tq[tid] = a.plan.tab_lo[tid]; tq[16 + tid] = a.plan.tab_hi[tid]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < R; ++k) {
      tlo[k] = *reinterpret_cast<volatile uint32_t*>(&tq[k]);
      thi[k] = *reinterpret_cast<volatile uint32_t*>(&tq[16 + k]);
    }
    __syncthreads();
  }

  // dynamic tile scheduling: CTAs pull tiles from a global counter so that all of them finish
  // together (with a static split the first-launched CTA of an SM runs ahead and the SM ends with
  // too few warps); the tile for iteration it+1 is claimed when its prefetch is issued
  auto claim = [&]() -> unsigned long long { return atomicAdd(&a.stats[2], 1ull); };
  if (tid == 0) {
    unsigned long long t0 = claim();
    misc->next_tile[0] = t0;
    misc->thr_key = 0xffffffffu;
    if (t0 < num_tiles) stage(t0, 0);
  }
  unsigned long long n_cand = 0, n_over = 0;
#ifdef LM_PRUNE_DEBUG
  unsigned long long dbg_t0 = 0, dbg_g0 = 0;
  if (tid == 0) { dbg_t0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_g0)); }
#endif
  // threshold key prefetched one tile ahead (L2 latency off the critical path; stale is safe)
  uint32_t sk_next = 0;
  if (!THR && warp == 0) sk_next = (uint32_t)(__ldcg(&a.colkeys[lane]) >> 32);
  __syncthreads();

  uint32_t it = 0;
  for (;; ++it) {
    const int buf = it & 1;
    const uint64_t tile = misc->next_tile[buf];
    if (tile >= num_tiles) break;
    // ---- current threshold -> byte bias (recomputed only when the threshold moved)
    if (warp == 0) {
      uint32_t sk = sk_next;
      if (!THR) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sk = max(sk, __shfl_xor_sync(0xffffffffu, sk, o));
      }
      if (lane == 0 && (THR ? it == 0 : sk != misc->thr_key)) {
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
        misc->bias4 = (uint32_t)bias * 0x01010101u;
        misc->thr = t;
        misc->thr_key = sk;
      }
    }
    mbar_wait(&misc->bars[buf], (it >> 1) & 1);
    const uint64_t tile_row0 = a.row_first + tile * PTILE_ROWS;
    uint8_t* rawb = raw + (size_t)buf * S.raw_bytes;
    {
      uint64_t avail = a.avail_rows > tile_row0 ? a.avail_rows - tile_row0 : 0;
      if (avail < tt) {
        uint32_t w0 = (uint32_t)avail * 8, w1 = tt * 8;
        uint32_t fill = a.pad_sym * 0x01010101u;
        for (uint32_t w = w0 + tid; w < w1; w += PT) reinterpret_cast<uint32_t*>(rawb)[w] = fill;
      }
    }
    __syncthreads();  // (A) tile + bias visible; every thread is done with the previous tile
    {
      // claim + prefetch the next tile into the other buffer (free since barrier A)
      if (tid == 0) {
        unsigned long long nt = claim();
        misc->next_tile[buf ^ 1] = nt;
        if (nt < num_tiles) stage(nt, buf ^ 1);
      }
      if (!THR && warp == 0) sk_next = (uint32_t)(__ldcg(&a.colkeys[lane]) >> 32);
    }
    const uint32_t bias4 = misc->bias4;
    const float thr = misc->thr;
    const uint64_t left = a.row_end - tile_row0;
    const uint32_t rows_valid = left < PTILE_ROWS ? (uint32_t)left : PTILE_ROWS;
    uint32_t* qcount = &misc->qcount[buf];

    // ---- filter
    {
      const uint32_t row_base = warp * WARP_ROWS + g * PL;
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(rawb) + (row_base + a.plan.j0) * 8 + q;
      uint32_t sel[PL + R - 1];
#pragma unroll
      for (int i = 0; i < R - 1; ++i) sel[i] = make_sel(rp[i * 8]);
      auto push = [&](uint32_t acc, int i) {
        const uint32_t msk = acc & 0x80808080u;
        if (msk) {  // queue the quad: row, column quad, which of its 4 byte lanes passed
          const uint32_t idx = atomicAdd(qcount, 1u);
          if (idx < QCAP) queue[idx] = ((row_base + i) << 7) | (q << 4) | (((msk >> 7) * 0x10204080u) >> 28);
        }
      };
      uint32_t prev = 0;
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        sel[i + R - 1] = make_sel(rp[(i + R - 1) * 8]);
        uint32_t acc = bias4;
#pragma unroll
        for (int k = 0; k < R; ++k) acc += prmt_raw(tlo[k], thi[k], sel[i + k]);
        // one test per PAIR of quads ((a | b) & mask is a single LOP3); the rare hit re-tests each
        if (i & 1) {
          if ((prev | acc) & 0x80808080u) { push(prev, i - 1); push(acc, i); }
        } else if (i == PL - 1) {
          if (acc & 0x80808080u) push(acc, i);
        }
        prev = acc;
      }
    }
    __syncthreads();  // (B) queue complete

    // ---- exact recheck of the candidates (or of the whole tile on overflow)
    const uint32_t nq = *qcount;
    if (tid == 0) misc->qcount[buf ^ 1] = 0;  // nobody touches the other counter until the next barrier A
    if (nq <= QCAP) {
      for (uint32_t k = tid; k < nq * 4; k += PT) {
        const uint32_t code = queue[k >> 2], b = k & 3;
        if (!((code >> b) & 1u)) continue;
        const uint32_t row = code >> 7, col = ((code >> 4) & 7u) * 4 + LANE_COL[b];
        if (row >= rows_valid) continue;
        float s = exact_score(rawb, tab, m, row, col);
        if (s >= thr) record<THR>(a, s, (uint32_t)(tile_row0 + row - a.row_begin), col);
      }
      if (tid == 0) n_cand += nq;
    } else {
      for (uint32_t p = tid; p < rows_valid * 32; p += PT) {
        uint32_t row = p >> 5, col = p & 31;
        float s = exact_score(rawb, tab, m, row, col);
        if (s >= thr) record<THR>(a, s, (uint32_t)(tile_row0 + row - a.row_begin), col);
      }
      if (tid == 0) n_over += 1;
    }
  }
#ifdef LM_PRUNE_DEBUG
  if (tid == 0 && blockIdx.x < 2048) {
    unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    g_dbg[blockIdx.x][0] = clock64() - dbg_t0;
    g_dbg[blockIdx.x][1] = n_cand;
    g_dbg[blockIdx.x][2] = smid | ((unsigned long long)it << 32);
    g_dbg[blockIdx.x][3] = dbg_g0 | 0;  // start time (ns)
    (void)g1;
  }
#endif
  if (tid == 0 && a.stats) {
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

#ifdef LM_PRUNE_DEBUG
extern "C" void lm_debug_dump(unsigned long long* out, int n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 4 * n);
}
#endif

cudaError_t launch_prune(bool threshold_mode, const PruneArgs& a, int sm_count, cudaStream_t stream) {
  return threshold_mode ? launch_r<true>(a, sm_count, stream) : launch_r<false>(a, sm_count, stream);
}

}  // namespace lm
