// prune.cu -- exact filter + recheck scan (DNA), see prune.cuh.
//
// Kernel.  One CTA = 8 consumer (filter) warps + 1 producer warp, persistent.
//
//  producer claims tiles (8 x 4 x PL = 992 striped rows x 32 columns = 31 744 positions) one at a
//           time from a global counter -- so the whole grid sweeps the sequence as one dense
//           window --, stages each tile (+ halo) from HBM with ONE bulk-TMA copy into a ring of
//           NSTAGE shared-memory buffers (full[s] / empty[s] mbarriers), and publishes with each
//           tile the threshold it is to be filtered at (the running maximum, read from L2).
//           There is no CTA-wide barrier in steady state.
//  filter   consumer warp w owns rows [w * 4 PL, (w + 1) * 4 PL) of the tile; lane = (column quad
//           q = lane & 7, row group g = lane >> 3); a thread walks PL consecutive rows of its 4
//           columns.  One LDS.32 per row yields 4 symbols (4 adjacent columns = 4 independent
//           positions); one IDP.2A packs them into a PRMT selector that is reused by every
//           look-up touching that row:   acc = bias + sum_k PRMT(table[k], selector[i + k])
//           (byte lanes, no carries across lanes by construction; the adds are issued as IMADs
//           so that they run on the FMA pipe beside the PRMTs on the ALU pipe).  A position
//           passes iff bit 7 of its byte is set, i.e. iff its integer upper bound reaches the
//           current threshold.  In PAIRED mode a look-up covers TWO motif rows: the 8-entry PRMT
//           table is indexed by the symbol of the first row and one class bit of the second
//           row's symbol (entry = w1(s) + max over the class of w2), a valid upper bound that
//           costs one extra IMAD + LOP3 per sequence row.
//  recheck  passing positions go to the warp's shared-memory queue (ballot-compacted, the queue
//           length lives in a warp-uniform register); after the filter pass the warp rescans
//           them EXACTLY (sequential f32 sum over all M rows in the reference's order, from the
//           tile still in shared memory) and records those that really reach the threshold
//           (atomicMax on the per-column winner keys / hit list append).
//  fallback a warp whose queue overflows (low-complexity sequence, very low threshold) rescans
//           its slice of the tile exactly, position by position.
//
// The bound:  UB(p) = c0 + delta * sum_q(p) >= real-valued score(p), and the
// fp32 sequential sum differs from the real sum by at most eps, so
//   fp32score(p) >= T  ==>  sum_q(p) >= ceil((T - eps - c0) / delta - 1e-6)   (1e-6: rounding of the division)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "prune.cuh"

namespace lm {

namespace {

#ifndef LM_PRUNE_PW
#define LM_PRUNE_PW 8
#endif
constexpr int PW = LM_PRUNE_PW;         // consumer (filter) warps per CTA
constexpr int PT = (PW + 1) * 32;       // + one producer warp
#ifndef LM_PRUNE_PSLEEP
#define LM_PRUNE_PSLEEP 50   // ns between the producer's polls of an `empty` barrier
#endif
#ifndef LM_PRUNE_CSLEEP
#define LM_PRUNE_CSLEEP 20   // ns between a consumer warp's polls of a `full` barrier
#endif
#ifndef LM_PRUNE_EXP
#define LM_PRUNE_EXP 0       // experiments (timing only, wrong results): 1 = stream the tiles without filtering,
#endif                       // 2 = tiles cycle through an L2-resident window, 4 = never leave the fast path
#ifndef LM_PRUNE_SPT
#define LM_PRUNE_SPT 8       // slices (one warp's worth of rows) per staged tile
#endif
#ifndef LM_PRUNE_NSTAGE
#define LM_PRUNE_NSTAGE (16 / LM_PRUNE_SPT)   // 2 x 32 KB per CTA (4 x 16 KB and 8 x 8 KB measured: no faster,
#endif                                        // the kernel is bound by its integer pipes, not by the ring)
constexpr int NSTAGE = LM_PRUNE_NSTAGE; // tile buffers in flight per CTA
#ifndef LM_PRUNE_PL
#define LM_PRUNE_PL 31
#endif
#ifndef LM_PRUNE_MINBLOCKS
#define LM_PRUNE_MINBLOCKS 3
#endif
constexpr int PL = LM_PRUNE_PL;        // rows per thread per tile (odd => conflict-free LDS)
constexpr int WARP_ROWS = 4 * PL;
constexpr int SPT = LM_PRUNE_SPT;          // slices per staged tile
constexpr int PTILE_ROWS = SPT * WARP_ROWS;
constexpr int CLAIM = PW / SPT > 0 ? PW / SPT : 1;  // consecutive tiles a producer claims per trip to the global counter
static_assert(NSTAGE * SPT > PW + SPT - 1 || NSTAGE * SPT >= 2 * PW, "ring too shallow for the slices in flight");
constexpr int QCAP = 128;               // queue entries (quads) per warp and tile; more => exact fallback
#ifndef LM_PRUNE_ILV
#define LM_PRUNE_ILV 12
#endif
constexpr int BATCH_ILV = LM_PRUNE_ILV;  // motifs whose work items are interleaved in a many-motif launch
static_assert(PL % 2 == 1, "PL must be odd (bank-conflict-free row groups)");

// What a consumer warp needs to know about the tile in a stage.  Everything that depends on the
// MOTIF travels with the tile, so one launch can sweep many motifs over the sequence (work item =
// (motif, tile)): the filter tables, the filter window, the recheck table (a per-stage copy in
// shared memory) and the motif's result block.
struct TileDesc {
  unsigned long long tile;      // tile index within the motif's sweep; >= num_tiles: end marker
  unsigned long long* block;    // the motif's result block (keys[32] | first cell | hit count | statistics)
  lm_hit* hits;                 // threshold mode
  unsigned long long cap;
  uint32_t bias4;               // 128 - (quantised threshold), replicated in the 4 byte lanes
  float thr;                    // the threshold the tile is filtered / rechecked at
  uint32_t avail, valid;        // rows the copy delivers; rows of the tile inside the scored range
  uint32_t mq;                  // motif index within the launch
  int m_pad, j0;
  uint32_t pair_mul;
  uint32_t tabs[32];            // filter tables: entries 0..3 of look-up k in tabs[k], 4..7 in tabs[16 + k]
};
// pipeline control block: full[s] (TMA landed + descriptor valid) / empty[s] (all slices of the tile done)
struct Ctl {
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint32_t next_slice;      // slices (WARP_ROWS rows of a tile) handed out to the consumer warps so far
  uint32_t stage_mq[NSTAGE];  // motif whose recheck table each stage's table slot holds (producer only)
  TileDesc desc[NSTAGE];
};
struct PSmem {
  uint32_t table_off, table_bytes, ctl_off, queue_off, raw_off, raw_bytes, total;
};
__host__ __device__ inline PSmem psmem(int m_max, int tab_stride) {
  PSmem s;
  const uint32_t mp = (uint32_t)((m_max + 3) & ~3);
  const uint32_t tt = PTILE_ROWS + mp;
  s.table_off = 0;
  s.table_bytes = (mp * tab_stride * 4 + 127) & ~127u;  // one recheck table per stage
  s.ctl_off = NSTAGE * s.table_bytes;
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
__device__ __forceinline__ uint32_t add_fma_pipe(uint32_t p, uint32_t one, uint32_t acc) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(p), "r"(one), "r"(acc));
  return d;
}

// PRMT selector of a quad (4 symbols, one per byte, each <= 7): a single IDP.2A (FMA pipe)
//   dp2a.lo(x, {1, 16}) = (s0 + 256 s1) * 1 + (s2 + 256 s3) * 16 = s0 | s2 << 4 | s1 << 8 | s3 << 12
// packs the four symbols as nibbles in the order (0, 2, 1, 3).  Byte lane b of every look-up
// therefore belongs to column LANE_COL[b] of the quad.
__device__ __forceinline__ uint32_t make_sel(uint32_t x) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0x00001001u), "r"(0u));
  return d;
}
__device__ __constant__ const uint8_t LANE_COL[4] = {0, 2, 1, 3};

// a cell that reaches the threshold: hit record (threshold mode) or per-column winner (argmax mode)
template <bool THR>
__device__ __forceinline__ void record(const PruneArgs& a, const TileDesc& d, float s, uint32_t rel_row, uint32_t col) {
  if (THR) {
    unsigned long long slot = atomicAdd(d.block + 33, 1ull);
    if (slot < d.cap) {
      lm_hit h;
      h.row = rel_row + a.row_add;
      h.col = col;
      h.position = (uint64_t)col * (a.pos_rows ? a.pos_rows : a.row_end - a.row_begin) + h.row;
      h.score = s;
      d.hits[slot] = h;
    }
  } else {
    unsigned long long key = ((unsigned long long)f32_to_key(s) << 32) | rel_row;
    atomicMax(&d.block[col], key);
  }
}

// exact sequential f32 score of the window starting at (row_in_tile, col).  The table is padded
// to a multiple of 4 rows with +0.0 (x + 0.0 == x: the chain starts at +0.0 and never yields -0.0).
__device__ __forceinline__ float exact_score(const uint8_t* rawb, const float* tab, int m_pad,
                                             uint32_t row_in_tile, uint32_t col) {
  // The candidate's warp waits for this chain: keep its latency short.  All symbol loads of a block
  // of 8 motif rows are issued before the first look-up, all look-ups before the first add; only the
  // adds are sequential (they must be).
  float s = 0.0f;
  const uint8_t* p = rawb + row_in_tile * 32 + col;
  int j = 0;
  for (; j + 8 <= m_pad; j += 8) {
    uint32_t sy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sy[k] = p[32 * k];
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = tab[8 * k + sy[k]];
#pragma unroll
    for (int k = 0; k < 8; ++k) s = __fadd_rn(s, w[k]);
    p += 256;
    tab += 64;
  }
  if (j < m_pad) {  // m_pad is a multiple of 4
    const uint32_t s0 = p[0], s1 = p[32], s2 = p[64], s3 = p[96];
    const float w0 = tab[s0], w1 = tab[8 + s1], w2 = tab[16 + s2], w3 = tab[24 + s3];
    s = __fadd_rn(s, w0);
    s = __fadd_rn(s, w1);
    s = __fadd_rn(s, w2);
    s = __fadd_rn(s, w3);
  }
  return s;
}
// same from the residues in global memory (protein: the tile holds class codes); rows past the
// readable extent are the default symbol
// (not inlined: it is cold code, and the consumer loop must stay small -- beyond ~40 KB of SASS the
// warps start to wait for instruction fetches: "no_instruction" rose from 0.4 to 1.0 in ncu)
__device__ __noinline__ float exact_score_global(const uint8_t* seq_exact, uint64_t avail_rows, uint32_t pad_sym,
                                                 const float* tab, int tab_stride, int m_pad, uint64_t row,
                                                 uint32_t col) {
  float s = 0.0f;
  for (int j = 0; j < m_pad; ++j) {
    const uint64_t r = row + j;
    const uint32_t sym = r < avail_rows ? seq_exact[r * 32 + col] : pad_sym;
    s = __fadd_rn(s, tab[j * tab_stride + sym]);
  }
  return s;
}

// Recheck of a LONG candidate queue (real genomes at a modest maximum, low thresholds: tens of
// candidates per slice).  The in-line loop of the kernel gives each lane one FLAG of an entry --
// usually one flag in eight is set, so seven lanes in eight idle through the exact sum.  Here the
// set flags of 32 entries are enumerated with a warp prefix sum, so that every lane rescores a
// distinct candidate.  Compiled only into the DENSE instances of the kernel: the consumer loop of the
// sparse instance (one or two candidates per slice: the headline case) must stay small (instruction
// cache), and a non-inlined call costs it registers (measured: -11 %).  The host picks the instance
// from the candidate density the plan predicts -- or, for a matrix it has scanned before, observed.
struct RecheckCtx {
  const uint8_t* rawb;
  const float* tab;
  const uint32_t* queue;
  unsigned long long* block;
  lm_hit* hits;
  unsigned long long cap;
  uint64_t tile_row0, row_begin, row_add, pos_rows;
  float thr;
  int m_pad;
  uint32_t nq, tile_valid;
};
template <bool THR>
__device__ __forceinline__ void recheck_many(const RecheckCtx& c) {
  const int lane = threadIdx.x & 31;
  for (uint32_t base = 0; base < c.nq; base += 32) {
    const uint32_t code = base + lane < c.nq ? c.queue[base + lane] : 0u;
    const uint32_t cnt = __popc(code & 0xffu);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    for (uint32_t t = 0; t < total; t += 32) {
      const uint32_t want = t + lane;
      uint32_t e = 0;  // entry holding candidate `want`: first lane whose inclusive count exceeds it
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, incl, (e + step - 1) & 31);
        if (v <= want) e += step;
      }
      e &= 31;
      const uint32_t code_e = __shfl_sync(0xffffffffu, code, e);
      const uint32_t incl_e = __shfl_sync(0xffffffffu, incl, e);
      if (want >= total) continue;
      const uint32_t flags = code_e & 0xffu;
      const uint32_t k = want - (incl_e - __popc(flags));  // k-th set flag of the entry
      const uint32_t b = __fns(flags, 0, k + 1);
      const uint32_t row = (code_e >> 11) + (b >> 2), col = ((code_e >> 8) & 7u) * 4 + LANE_COL[b & 3];
      if (row >= c.tile_valid) continue;
      const float sc = exact_score(c.rawb, c.tab, c.m_pad, row, col);
      if (sc >= c.thr) {
        const uint32_t rel_row = (uint32_t)(c.tile_row0 + row - c.row_begin);
        if (THR) {
          unsigned long long slot = atomicAdd(c.block + 33, 1ull);
          if (slot < c.cap) {
            lm_hit h;
            h.row = rel_row + c.row_add;
            h.col = col;
            h.position = (uint64_t)col * c.pos_rows + h.row;
            h.score = sc;
            c.hits[slot] = h;
          }
        } else {
          atomicMax(&c.block[col], ((unsigned long long)f32_to_key(sc) << 32) | rel_row);
        }
      }
    }
  }
}

// BATCH = false: one motif, whose filter parameters are kernel arguments (uniform registers) and whose
// tables are loaded into registers once -- the instance the single-motif scans (and the headline
// benchmark) run; BATCH = true: the parameters travel with the tiles.  The look-up loop is identical,
// but it needs every register ptxas has (72 at 3 CTAs/SM): two more live values and it either spills
// or gives up interleaving the selector loads with the look-ups (measured: -10 %).
template <int R, bool PAIR, bool THR, bool BATCH, bool DENSE>
__global__ void __launch_bounds__(PT, LM_PRUNE_MINBLOCKS) prune_kernel(const PruneArgs a, uint64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PSmem S = psmem(a.m_max, a.tab_stride);
  const uint32_t halo = (uint32_t)((a.m_max + 3) & ~3);
  const uint32_t tt = PTILE_ROWS + halo;  // rows staged per tile (tile + halo of the widest motif)
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + S.ctl_off);
  uint8_t* raw = smem + S.raw_off;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&ctl->full[s], 32);  // every lane of the producer warp arrives: each one releases its own writes
      mbar_init(&ctl->empty[s], SPT);
    }
    ctl->next_slice = 0;
    for (int s = 0; s < NSTAGE; ++s) ctl->stage_mq[s] = 0xffffffffu;
    mbar_fence_init();
  }
  __syncthreads();  // the only CTA-wide barrier: from here on warps meet through mbarriers only

  if (warp == PW) {
    // ===================== producer warp: claims work items (motif, tile) from a global counter -- so
    // the CTAs of the whole grid sweep the sequence as one dense window --, stages each tile (+ halo)
    // with ONE bulk-TMA copy and publishes with it everything the consumers need: the threshold the
    // tile is to be filtered at (the motif's running maximum, read from L2) and the motif's tables.
    const uint64_t total_items = num_tiles * a.n_batch;
    uint32_t thr_key = 0xffffffffu, bias4 = 0, pushed_key = 0, sk_last = 0;
    float thr = -INFINITY;
    // the motif the producer currently serves: lane l keeps word l of its filter tables
    uint32_t cur_mq = 0xffffffffu, my_tab = 0;
    struct {  // (scalars only: no indexed member, so it lives in registers)
      double c0, delta, eps;
      int m, j0, smax;
      uint32_t pair_mul;
      const float* table;
      unsigned long long* block;
      lm_hit* hits;
      unsigned long long cap;
      float thr;
    } M = {0, 1, 0, 0, 0, 0, 2, nullptr, nullptr, nullptr, 0, 0.f};
    // CTA 0 measures the SM clock the scan runs at: cycle counter vs global nanosecond timer
    long long clk0 = 0;
    unsigned long long ns0 = 0;
    if (blockIdx.x == 0 && lane == 0) {
      clk0 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }
    uint32_t it = 0, n_end = 0, first_end = 0;
    unsigned long long claim_base = 0;
    for (;; ++it) {
      const int s = it % NSTAGE;
      // claim the next items and fetch the running maximum first: both round trips overlap the wait
      // (CLAIM consecutive tiles per trip: the tiles are small, the L2 round trip is not)
      unsigned long long item = claim_base + it % CLAIM;
      uint32_t sk = sk_last;
      const bool fresh = it % CLAIM == 0;
      if (fresh) {
        if (lane == 0) item = atomicAdd(&a.stats[2], (unsigned long long)CLAIM);
        item = __shfl_sync(0xffffffffu, item, 0);
        claim_base = item;
      }
      // Many motifs: the items of BATCH_ILV motifs are interleaved (tile-major within the block of motifs), so a
      // motif's sweep starts with grid / BATCH_ILV CTAs instead of the whole grid.  Every tile in flight when a
      // motif starts is filtered at the seed of its running maximum: with motif-major order that was the first
      // ~900 of a 3 150-tile sweep (config 2), and the seed's quality showed in the run time.  BATCH_ILV divides the
      // usual grid (444), so a CTA tends to stay with one motif and its tables stay in place.
      uint32_t mq = 0;
      unsigned long long tile = item;
      if (a.n_batch > 1) {
        const unsigned long long per_block = (unsigned long long)BATCH_ILV * num_tiles;
        const unsigned long long blk = item / per_block, r = item - blk * per_block;
        const uint32_t left = a.n_batch - (uint32_t)blk * BATCH_ILV;  // motifs in this block (the last one may be short)
        const uint32_t gb = left < (uint32_t)BATCH_ILV ? left : (uint32_t)BATCH_ILV;
        mq = (uint32_t)blk * BATCH_ILV + (uint32_t)(r % gb);
        tile = r / gb;
      }
      const bool live = item < total_items;
      if (live && mq != cur_mq) {  // (warp-uniform) next motif: its plan, tables and result block
        const BatchMotif* src = a.batch ? &a.batch[mq] : &a.one;
        M.c0 = src->c0; M.delta = src->delta; M.eps = src->eps;
        M.m = src->m; M.j0 = src->j0; M.smax = src->smax; M.pair_mul = src->pair_mul;
        M.table = src->table; M.block = src->block; M.hits = src->hits; M.cap = src->cap; M.thr = src->thr;
        my_tab = a.batch ? a.batch[mq].tab[lane] : a.one.tab[lane];
        cur_mq = mq;
        thr_key = 0xffffffffu;
        sk_last = 0;
      }
      if (!THR && live && (fresh || sk_last == 0)) {
        sk = (uint32_t)(__ldcg(&M.block[lane]) >> 32);
        // row-sharded scan: the maximum the OTHER GPUs have found so far (pushed over NVLink)
        if (a.link.world > 1 && lane == 0) sk = max(sk, peer_floor_key(a.link));
        uint32_t own = lane == 0 && a.link.world > 1 ? (uint32_t)(__ldcg(&M.block[0]) >> 32) : sk;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sk = max(sk, __shfl_xor_sync(0xffffffffu, sk, o));
          own = max(own, __shfl_xor_sync(0xffffffffu, own, o));
        }
        // CTA 0 tells the other GPUs whenever this GPU's own maximum has risen above the shared one
        if (a.link.world > 1 && blockIdx.x == 0 && own > pushed_key) {
          if (own > peer_floor_key(a.link) || pushed_key == 0) peer_push_floor(a.link, own, lane);
          pushed_key = own;
        }
        sk_last = sk;
      }
      // ---- current threshold -> byte bias (recomputed only when the threshold moved)
      if (live && (THR ? thr_key == 0xffffffffu : sk != thr_key)) {
        const float t = THR ? M.thr : (sk ? key_to_f32(sk) : -INFINITY);
        // sum_q >= qthr is implied by fp32score >= t (see header); 1e-6 guards the division
        double need = ((double)t - M.eps - M.c0) / M.delta;
        int bias;
        if (!(need > 1.0)) bias = 128;                               // everything passes
        else if (need > (double)M.smax + 2.0) bias = 0;              // nothing can pass
        else {
          int qthr = (int)ceil(need - 1e-6);
          qthr = qthr < 1 ? 1 : (qthr > 128 ? 128 : qthr);
          bias = 128 - qthr;
        }
        bias4 = (uint32_t)bias * 0x01010101u;
        thr = t;
        thr_key = THR ? 0u : sk;
      }
      // everything the reload needs is computed BEFORE the wait: between "stage free" and "copy
      // issued" only a few shared-memory stores, the expect_tx and the bulk copy remain
#if LM_PRUNE_EXP == 2
      const uint64_t r0 = a.row_first + (tile & 1023) * PTILE_ROWS;
#else
      const uint64_t r0 = a.row_first + tile * PTILE_ROWS;
#endif
      const uint64_t avail = a.avail_rows > r0 ? a.avail_rows - r0 : 0;
      const uint32_t nbytes = (avail < tt ? (uint32_t)avail : tt) * 32;
      const uint8_t* src = a.seq + r0 * 32;
      uint8_t* dst = raw + (size_t)s * S.raw_bytes;
      const uint64_t left = a.row_end > r0 ? a.row_end - r0 : 0;
      const uint32_t valid = left < PTILE_ROWS ? (uint32_t)left : PTILE_ROWS;
      if (it >= NSTAGE) mbar_wait_backoff(&ctl->empty[s], ((it / NSTAGE) - 1) & 1, LM_PRUNE_PSLEEP);
      TileDesc& d = ctl->desc[s];
      if (live) {
        uint32_t in_stage = 0;
        if (lane == 0) in_stage = ctl->stage_mq[s];
        in_stage = __shfl_sync(0xffffffffu, in_stage, 0);
        if (in_stage != mq) {  // (rare) another motif moves into this stage: its recheck table and filter tables
          float* tab = reinterpret_cast<float*>(smem + S.table_off + (size_t)s * S.table_bytes);
          const int words = ((M.m + 3) & ~3) * a.tab_stride;
          for (int i = lane; i < words; i += 32) tab[i] = __ldg(&M.table[i]);  // rows >= m are +0.0
          d.tabs[lane] = my_tab;
          if (lane == 0) ctl->stage_mq[s] = mq;
        }
        if (lane == 0) {
          d.tile = tile;
          d.block = M.block;
          d.hits = M.hits;
          d.cap = M.cap;
          d.bias4 = bias4;
          d.thr = thr;
          d.avail = nbytes / 32;
          d.valid = valid;
          d.mq = mq;
          d.m_pad = (M.m + 3) & ~3;
          d.j0 = M.j0;
          d.pair_mul = M.pair_mul;
        }
        __syncwarp();
        if (lane == 0) {
          // (generic-proxy writes into the stage -- the consumers' padding of rows past the end of the
          // sequence -- are fenced by the writers before they release the stage)
          mbar_expect_tx(&ctl->full[s], nbytes);
          if (nbytes) tma_load_1d(dst, src, nbytes, &ctl->full[s]);
        } else {
          mbar_arrive(&ctl->full[s]);  // (release: this lane's table words are visible to the consumers)
        }
      } else {
        // past the end: keep handing out end markers until every consumer warp has met one (the slices
        // of a tile go to whichever warps are free, so up to PW further slices may have been taken)
        if (lane == 0) d.tile = ~0ull;
        mbar_arrive(&ctl->full[s]);
        if (n_end == 0) first_end = it;
        if (++n_end > (uint32_t)((PW + SPT - 1) / SPT)) break;
      }
      __syncwarp();
    }
    // wait until the consumers are done with the tiles still in the other stages: after that this
    // CTA's winners, hits and statistics are final
    for (uint32_t it2 = it + 1; it2 < first_end + NSTAGE; ++it2) {
      if (it2 < NSTAGE) continue;
      const int s2 = it2 % NSTAGE;
      mbar_wait_backoff(&ctl->empty[s2], ((it2 / NSTAGE) - 1) & 1, LM_PRUNE_PSLEEP);
    }
    if (blockIdx.x == 0 && lane == 0) {
      unsigned long long ns1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
      a.stats[3] = (unsigned long long)(clock64() - clk0);
      a.stats[4] = ns1 - ns0;
    }
    if (a.link.world > 1) {
      // row-sharded scan: the LAST producer warp of the grid to finish (every CTA's winners, hits and
      // statistics are final by then) hands this GPU's result block to every GPU and raises the flags
      unsigned long long done = 0;
      if (lane == 0) {
        __threadfence();
        done = atomicAdd(&a.stats[5], 1ull);
      }
      done = __shfl_sync(0xffffffffu, done, 0);
      if (done == (unsigned long long)gridDim.x - 1) {
        __threadfence();
        peer_publish(a.link, a.one.block, lane);
      }
    }
    return;
  }

  // ===================== consumer warps: filter slices of the staged tiles, recheck their candidates
  const int q = lane & 7, g = lane >> 3;
  uint32_t* const queue = reinterpret_cast<uint32_t*>(smem + S.queue_off) + warp * QCAP;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  // filter tables in per-thread registers (PRMT cannot take its table operand from a uniform register,
  // and values that ptxas knows to be uniform would be re-materialised with a UR->R move before every
  // look-up: hence the volatile shared-memory reads).  Reloaded when the motif changes.
  uint32_t tlo[R], thi[R];
  const uint32_t one = a.one_mul;
  if (!BATCH) {
    if (lane < 16) { queue[lane] = a.one.tab[lane]; queue[16 + lane] = a.one.tab[16 + lane]; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
      tlo[k] = *reinterpret_cast<volatile uint32_t*>(&queue[k]);
      thi[k] = *reinterpret_cast<volatile uint32_t*>(&queue[16 + k]);
    }
    __syncwarp();
  }

  for (;;) {
    // The SPT slices of a tile (WARP_ROWS rows each) are not tied to warps: a free consumer warp takes
    // the next slice of the CTA's tile sequence.  A warp that had candidates to recheck simply takes
    // its next slice a little later while the others carry on.  At most PW slices are in flight,
    // spanning fewer than NSTAGE consecutive tiles, so the parity waits below cannot alias.
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(&ctl->next_slice, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    const uint32_t it = idx / SPT;
    const int s = it % NSTAGE;
    mbar_wait_backoff(&ctl->full[s], (it / NSTAGE) & 1, LM_PRUNE_CSLEEP);
    const TileDesc& d = ctl->desc[s];
    if (d.tile >= num_tiles) break;
    if (BATCH) {
      // the tables travel with the tile and are re-read for every slice (2 R shared-memory loads per
      // ~750 instructions): a "did the motif change?" test would cost a live register across the
      // look-up loop, and with it ptxas stops interleaving the selector loads with the look-ups
#pragma unroll
      for (int k = 0; k < R; ++k) {
        tlo[k] = *reinterpret_cast<const volatile uint32_t*>(&d.tabs[k]);
        thi[k] = *reinterpret_cast<const volatile uint32_t*>(&d.tabs[16 + k]);
      }
    }
    // (only what the filter itself needs is loaded here; the recheck re-reads the rest of the descriptor
    // when it runs -- fewer live registers across the look-up loop, which ptxas otherwise schedules worse)
    const uint32_t bias4 = d.bias4;
    const uint32_t wrow0 = (idx % SPT) * WARP_ROWS;                  // first row of the slice in the tile
    uint8_t* rawb = raw + (size_t)s * S.raw_bytes;
    {
      // rows past the readable extent (end of the sequence only): default symbol.  Every warp fills
      // the rows IT reads; overlapping halo rows get the same value from both neighbours.
      const uint32_t avail = d.avail;
      const uint32_t wend = wrow0 + WARP_ROWS + halo;
      if (avail < wend) {
        uint32_t w0 = (avail > wrow0 ? avail : wrow0) * 8, w1 = wend * 8;
        uint32_t fill = a.pad_fill * 0x01010101u;
        for (uint32_t w = w0 + lane; w < w1; w += 32) reinterpret_cast<uint32_t*>(rawb)[w] = fill;
        fence_proxy_async();  // these generic-proxy writes precede the next bulk copy into this stage
        __syncwarp();
      }
    }
    const uint32_t tile_valid = d.valid;  // > 0 for every claimed tile
    const uint32_t rows_valid = tile_valid > wrow0 ? min(tile_valid - wrow0, (uint32_t)WARP_ROWS) : 0;

#if LM_PRUNE_EXP == 1
    if (rows_valid && bias4 == 0x12345678u) {
#else
    if (rows_valid) {
#endif
      // ---- filter
      const uint32_t row_base = wrow0 + g * PL;
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(rawb) + (row_base + (BATCH ? d.j0 : a.one.j0)) * 8 + q;
      // selector window.  Single mode: look-up k of position i uses row i + k.  Paired mode:
      // look-up k uses the COMBINED selector of rows (i + 2k, i + 2k + 1): symbol of the first
      // row | one class bit of the second row << 2 (8-entry table = 4 symbols x 2 classes).
      constexpr int W = PAIR ? 2 * R : R;        // motif rows spanned by the filter
      constexpr int STEP = PAIR ? 2 : 1;
      constexpr int NSEL = PAIR ? PL + W - 2 : PL + W - 1;
      uint32_t sel[NSEL];
      uint32_t raw_prev = 0;
      const uint32_t pair_mul = a.one.pair_mul;  // (launch-uniform: batches are grouped by it)
      auto combine = [&](uint32_t lo_row, uint32_t hi_row) {  // IMAD (FMA pipe) + LOP3
        return lo_row | ((hi_row * pair_mul) & 0x4444u);
      };
      if (PAIR) {
        raw_prev = make_sel(rp[0]);
#pragma unroll
        for (int i = 0; i < W - 2; ++i) {
          const uint32_t nx = make_sel(rp[(i + 1) * 8]);
          sel[i] = combine(raw_prev, nx);
          raw_prev = nx;
        }
      } else {
#pragma unroll
        for (int i = 0; i < W - 1; ++i) sel[i] = make_sel(rp[i * 8]);
      }
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
      // The look-up loop runs table-outer / row-inner over the thread's whole run of PL rows, so that
      // consecutive PRMTs share their two table registers (operand reuse: PRMT is limited by
      // register-file ports, not only by the ALU pipe) and independent accumulator chains hide the
      // IMAD latency.
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        if (PAIR) {
          const uint32_t nx = make_sel(rp[(i + W - 1) * 8]);
          sel[i + W - 2] = combine(raw_prev, nx);
          raw_prev = nx;
        } else {
          sel[i + W - 1] = make_sel(rp[(i + W - 1) * 8]);
        }
      }
      uint32_t acc[PL];
#pragma unroll
      for (int i = 0; i < PL; ++i) acc[i] = bias4;
      // The adds are IMADs (`p * 1 + acc`, FMA pipe) beside the PRMTs on the ALU pipe.  NP3 PAIRS of look-ups are
      // instead summed by one 3-input add (IADD3, ALU pipe): two issue slots fewer per pair, one more ALU-pipe
      // instruction.  Measured on the headline instance (paired R = 6): one pair +1.4 %, two -1 %, three -4 %
      // (the ALU pipe becomes the limiter); one pair in every many-motif instance: config 2 -2.3 %; the other
      // single-motif instances measured within +-1 % either way and keep NP3 = 0.
#ifdef LM_PRUNE_ADD3
      constexpr int NP3 = (LM_PRUNE_ADD3 < R / 2 ? LM_PRUNE_ADD3 : R / 2);
#else
      constexpr int NP3 = (BATCH || (PAIR && R == 6 && !DENSE)) ? 1 : 0;
#endif
#pragma unroll
      for (int kp = 0; kp < NP3; ++kp) {
#pragma unroll
        for (int i = 0; i < PL; ++i)
          acc[i] = acc[i] + prmt_raw(tlo[2 * kp], thi[2 * kp], sel[i + STEP * 2 * kp]) +
                   prmt_raw(tlo[2 * kp + 1], thi[2 * kp + 1], sel[i + STEP * (2 * kp + 1)]);
      }
#pragma unroll
      for (int k = 2 * NP3; k < R; ++k) {
#pragma unroll
        for (int i = 0; i < PL; ++i) acc[i] = add_fma_pipe(prmt_raw(tlo[k], thi[k], sel[i + STEP * k]), one, acc[i]);
      }
      // one test for the whole run; the rare hit narrows it down through sub-blocks of 8 rows (their
      // ORs are by-products of the reduction) before it looks at row pairs
      constexpr int NSUB = (PL + 7) / 8;
      uint32_t sub[NSUB];
#pragma unroll
      for (int u = 0; u < NSUB; ++u) sub[u] = 0;
#pragma unroll
      for (int i = 0; i < PL; ++i) sub[i >> 3] |= acc[i];
      uint32_t any = 0;
#pragma unroll
      for (int u = 0; u < NSUB; ++u) any |= sub[u];
#if LM_PRUNE_EXP == 4
      if (__any_sync(0xffffffffu, (any & 0x80808080u) != 0) && bias4 == 0x12345678u) {
#else
      if (__any_sync(0xffffffffu, (any & 0x80808080u) != 0)) {
#endif
#pragma unroll
        for (int u = 0; u < NSUB; ++u) {
          if (!__any_sync(0xffffffffu, (sub[u] & 0x80808080u) != 0)) continue;
#pragma unroll
          for (int i = 8 * u; i < 8 * u + 8 && i < PL; i += 2) {
            if (i + 1 < PL) push2(acc[i], acc[i + 1], i);
            else push2(acc[i], 0u, i);
          }
        }
      }
      __syncwarp();  // queue complete

      if (nq) {
        // ---- exact recheck (sequential f32 sum in the reference's order, from the tile still in
        // shared memory) of the queued candidates -- or, when the queue overflowed (low-complexity
        // sequence, very low threshold), of every position of the slice.  One loop for both cases:
        // this is cold code and the consumer loop must stay small.
        const volatile TileDesc& dv = ctl->desc[s];
        // (lane 0 reads the descriptor and is the lane that releases the stage below)
        float thr = 0.f;
        int m_pad = 0;
        unsigned long long dtile = 0;
        if (lane == 0) { thr = dv.thr; m_pad = dv.m_pad; dtile = dv.tile; }
        thr = __shfl_sync(0xffffffffu, thr, 0);
        m_pad = __shfl_sync(0xffffffffu, m_pad, 0);
        dtile = __shfl_sync(0xffffffffu, dtile, 0);
        const float* tab = reinterpret_cast<const float*>(smem + S.table_off + (size_t)s * S.table_bytes);
        const uint64_t tile_row0 = a.row_first + dtile * PTILE_ROWS;
        const bool overflow = nq > QCAP;
        if (DENSE && !overflow && nq > 4 && !a.seq_exact) {
          RecheckCtx c;
          c.rawb = rawb; c.tab = tab; c.queue = queue; c.block = dv.block; c.hits = dv.hits; c.cap = dv.cap;
          c.tile_row0 = tile_row0; c.row_begin = a.row_begin; c.row_add = a.row_add;
          c.pos_rows = a.pos_rows ? a.pos_rows : a.row_end - a.row_begin;
          c.thr = thr; c.m_pad = m_pad; c.nq = nq; c.tile_valid = tile_valid;
          recheck_many<THR>(c);
          if (lane == 0) atomicAdd(d.block + 34, (unsigned long long)nq);
          nq = 0;
        }
        const uint32_t n_items = overflow ? rows_valid * 32 : nq * 8;
        for (uint32_t k = lane; k < n_items; k += 32) {
          uint32_t row, col;
          if (overflow) {
            row = wrow0 + (k >> 5);
            col = k & 31;
          } else {
            const uint32_t code = queue[k >> 3], b = k & 7;
            if (!((code >> b) & 1u)) continue;
            row = (code >> 11) + (b >> 2);
            col = ((code >> 8) & 7u) * 4 + LANE_COL[b & 3];
            if (row >= tile_valid) continue;
          }
          const float sc = a.seq_exact ? exact_score_global(a.seq_exact, a.avail_rows, a.pad_sym, tab, a.tab_stride, m_pad,
                                                            tile_row0 + row, col)
                                       : exact_score(rawb, tab, m_pad, row, col);
          if (sc >= thr) record<THR>(a, d, sc, (uint32_t)(tile_row0 + row - a.row_begin), col);
        }
        if (lane == 0 && (overflow || nq)) atomicAdd(d.block + (overflow ? 35 : 34), overflow ? 1ull : (unsigned long long)nq);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->empty[s]);  // one more slice of the stage is done
  }
}

template <bool THR>
cudaError_t launch_r(const PruneArgs& a_in, int sm_count, cudaStream_t stream) {
  if (a_in.row_end <= a_in.row_first) return cudaSuccess;
  PruneArgs a = a_in;
  a.one.pair_mul = a.plan.pair_mul;  // uniform over the launch (many-motif launches are grouped by it)
  if (!a.batch) {  // single motif: its parameters travel in the kernel arguments
    a.n_batch = 1;
    a.m_max = a.plan.m;
    BatchMotif& M = a.one;
    for (int k = 0; k < 16; ++k) { M.tab[k] = a.plan.tab_lo[k]; M.tab[16 + k] = a.plan.tab_hi[k]; }
    M.c0 = a.plan.c0; M.delta = a.plan.delta; M.eps = a.plan.eps;
    M.m = a.plan.m; M.j0 = a.plan.j0; M.smax = a.plan.smax; M.pair_mul = a.plan.pair_mul;
    M.table = a.table;
    M.block = a.colkeys;
    M.hits = a.hits; M.cap = a.cap; M.thr = a.thr;
  }
  const uint64_t num_tiles = (a.row_end - a.row_first + PTILE_ROWS - 1) / PTILE_ROWS;
  const PSmem S = psmem(a.m_max, a.tab_stride);
  void (*kern)(const PruneArgs, uint64_t) = nullptr;
  const bool batch = a.batch != nullptr;
#ifdef LM_PRUNE_ONLY  // experiment builds (tools/build_variants.py): the headline instance only
  if constexpr (!THR) {
    if (a.plan.pair && a.plan.r == 6 && !batch && !a.dense) kern = prune_kernel<6, true, false, false, false>;
  }
  if (!kern) return cudaErrorInvalidValue;
#else
#define LM_PICK(RR, PP)                                                                                      \
  kern = batch ? prune_kernel<RR, PP, THR, true, false>                                                    \
               : (a.dense ? prune_kernel<RR, PP, THR, false, true> : prune_kernel<RR, PP, THR, false, false>); \
  break;
  if (a.plan.pair) {
    switch (a.plan.r) {
      case 3: LM_PICK(3, true)
      case 4: LM_PICK(4, true)
      case 5: LM_PICK(5, true)
      case 6: LM_PICK(6, true)
      case 8: LM_PICK(8, true)
      default: return cudaErrorInvalidValue;
    }
  } else {
    switch (a.plan.r) {
      case 4: LM_PICK(4, false)
      case 6: LM_PICK(6, false)
      case 8: LM_PICK(8, false)
      case 10: LM_PICK(10, false)
      case 12: LM_PICK(12, false)
      case 16: LM_PICK(16, false)
      default: return cudaErrorInvalidValue;
    }
  }
#undef LM_PICK
#endif
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT, S.total);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = std::min<uint64_t>((uint64_t)sm_count * per_sm, num_tiles * a.n_batch);
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

namespace {

// One PRMT look-up of the filter: up to 8 table entries with their real-valued contents and
// their probabilities under a uniform ACGT background.
struct Unit {
  double v[8];
  double prob[8];
  int n;
};

// Modelled cost (issue slots per quad row = 128 positions of one warp instruction stream) of a
// filter with `r` look-ups spanning `w` motif rows whose per-position pass rate is `pass`.
// Calibrated on B200 (profiles/README.md).
double filter_cost(int r, int w, bool pair, double pass) {
  const double halo = (double)(PL + w - 1) / PL;
  const double core = 2.0 * r + (pair ? 5.0 : 2.0) * halo + 0.5;
  // slow path of the per-thread-tile test: taken when any of the 32 * PL quads of a warp tile has a
  // candidate; it re-tests the PL / 2 row pairs (~5 slots each) and appends the hits (~15 slots each)
  const double p_quad = std::min(1.0, 4.0 * pass);
  const double p_slow = 1.0 - std::pow(1.0 - p_quad, 32.0 * PL);
  const double push = (p_slow * 2.5 * PL + 15.0 * 32.0 * PL * p_quad) / PL;
  // exact recheck, per warp tile (WARP_ROWS * 32 positions): one pass of ~130 slots per 32 flags
  const double lambda = pass * WARP_ROWS * 32.0;
  const double recheck = (130.0 * (1.0 - std::exp(-lambda)) + 130.0 * lambda * 8.0 / 32.0) / PL;
  // queue overflow => the warp rescans its slice exactly (~350 slots per quad row)
  const double overflow = lambda > QCAP / 4.0 ? 350.0 : 0.0;
  // weights fitted to the measured sweep of forced configurations (profiles/README.md, r01 v5):
  // the candidate paths are partly hidden behind the other warps' filter work
  return core + 0.4 * push + 0.6 * recheck + overflow;
}
constexpr double EXACT_COST = 350.0;  // exact-everywhere kernel (width-20 DNA), same unit

// Quantise R look-up units into byte tables: clamp every unit from below at (its maximum - clamp),
// pick the common step delta so that the largest attainable sum is <= 127, round every entry UP.
// Returns false when no valid tables exist.  On success `pass` is the probability that a random
// window reaches the planning threshold t_plan, by exact convolution of the units' distributions.
struct Quantised {
  int q[16][8];
  int smax = 0;
  double c0 = 0, delta = 1, pass = 1;
};
bool quantise_units(Unit* u, int R, double clamp, double restmax, double t_plan, double eps, Quantised& out) {
  double base[16], range_sum = 0, base_sum = 0;
  for (int k = 0; k < R; ++k) {
    double mx = -INFINITY, mn = INFINITY;
    for (int e = 0; e < u[k].n; ++e) mx = std::max(mx, u[k].v[e]);
    if (!std::isfinite(mx)) return false;
    for (int e = 0; e < u[k].n; ++e) {
      u[k].v[e] = std::max(u[k].v[e], mx - clamp);
      mn = std::min(mn, u[k].v[e]);
    }
    base[k] = mn;
    base_sum += mn;
    range_sum += mx - mn;
  }
  if (!(range_sum > 0)) return false;
  const double delta = range_sum / (127.0 - R);
  int smax = 0;
  for (int k = 0; k < R; ++k) {
    int mxq = 0;
    for (int e = 0; e < u[k].n; ++e) {
      int v = (int)std::ceil((u[k].v[e] - base[k]) / delta - 1e-9);
      if (v < 0) v = 0;
      while (base[k] + v * delta < u[k].v[e]) v++;  // the bound must hold despite the slack
      out.q[k][e] = v;
      mxq = std::max(mxq, v);
    }
    for (int e = u[k].n; e < 8; ++e) out.q[k][e] = 0;
    smax += mxq;
  }
  if (smax > 127) return false;
  out.smax = smax;
  out.delta = delta;
  out.c0 = base_sum + restmax;
  const double need = (t_plan - eps - out.c0) / delta;
  const int qthr = need > 1.0 ? (int)std::ceil(need - 1e-6) : 0;
  if (qthr <= 0) out.pass = 1.0;
  else if (qthr > smax) out.pass = 0.0;
  else {
    double d[128] = {0}, nd[128];
    d[0] = 1.0;
    for (int k = 0; k < R; ++k) {
      std::fill(nd, nd + 128, 0.0);
      for (int e = 0; e < u[k].n; ++e)
        if (u[k].prob[e] > 0)
          for (int i = 0; i + out.q[k][e] < 128; ++i) nd[i + out.q[k][e]] += u[k].prob[e] * d[i];
      std::copy(nd, nd + 128, d);
    }
    out.pass = 0;
    for (int i = qthr; i < 128; ++i) out.pass += d[i];
  }
  return true;
}

void store_tables(PrunePlan& plan, const Quantised& Q, int R) {
  for (int k = 0; k < 16; ++k) plan.tab_lo[k] = plan.tab_hi[k] = 0;
  for (int k = 0; k < R; ++k) {
    plan.tab_lo[k] = (uint32_t)Q.q[k][0] | ((uint32_t)Q.q[k][1] << 8) | ((uint32_t)Q.q[k][2] << 16) | ((uint32_t)Q.q[k][3] << 24);
    plan.tab_hi[k] = (uint32_t)Q.q[k][4] | ((uint32_t)Q.q[k][5] << 8) | ((uint32_t)Q.q[k][6] << 16) | ((uint32_t)Q.q[k][7] << 24);
  }
  plan.smax = Q.smax;
  plan.c0 = Q.c0;
  plan.delta = Q.delta;
  plan.est_pass_rate = Q.pass;
}

}  // namespace

PrunePlan plan_prune(const float* pssm, size_t m, size_t stride, double t_plan, bool t_is_estimate,
                     double n_positions, int force_r, int force_mode) {
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
  static const int RS_SINGLE[] = {4, 6, 8, 10, 12, 16};
  static const int RS_PAIR[] = {3, 4, 5, 6, 8};
  double best_cost = 1e300;

  auto consider = [&](bool pair, int R, int j0, uint32_t mul, double C) {
    const int W = pair ? 2 * R : R;  // rows [j0, j0 + W) are covered; rows >= m do not exist
    Unit u[16];
    double restmax = max_total;
    for (int k = 0; k < R; ++k) {
      Unit& U = u[k];
      if (!pair) {
        const int row = j0 + k;
        restmax -= r.rowmax[row];
        U.n = 5;  // A, C, T, G, N (N: probability 0, but its entry must still bound its weight)
        for (int a = 0; a < 5; ++a) { U.v[a] = r.w[row * 5 + a]; U.prob[a] = a < 4 ? 0.25 : 0.0; }
      } else {
        const int rf = j0 + 2 * k, rh = rf + 1;  // full row, half (class-bit) row
        restmax -= r.rowmax[rf];
        // half row: best weight within each class of the second symbol.  class(s) = bit 2 of
        // s * mul: mul = 2 -> {T, G}, mul = 4 -> {C, G}; N (4) is always class 0.
        double hmax[2] = {0.0, 0.0}, hall = 0.0;
        if (rh < (int)m) {
          restmax -= r.rowmax[rh];
          hmax[0] = hmax[1] = -INFINITY;
          for (int a = 0; a < 5; ++a) {
            const int cls = ((a * mul) >> 2) & 1;
            hmax[cls] = std::max(hmax[cls], r.w[rh * 5 + a]);
          }
          hall = std::max(hmax[0], hmax[1]);
        }
        U.n = 8;  // entry = s | class << 2
        for (int b = 0; b < 2; ++b)
          for (int a = 0; a < 4; ++a) { U.v[a + 4 * b] = r.w[rf * 5 + a] + hmax[b]; U.prob[a + 4 * b] = 0.125; }
        // N in the full row encodes as nibble 4 | class << 2 = entry (A, class 1) whatever the
        // second symbol is: that entry must also bound  w(N) + any second-row weight
        U.v[4] = std::max(U.v[4], r.w[rf * 5 + 4] + hall);
      }
    }
    Quantised Q;
    if (!quantise_units(u, R, C * (pair ? 1.5 : 1.0), restmax, t_plan, eps, Q)) return;
    const double cost = filter_cost(R, std::min<int>(W, (int)m - j0), pair, Q.pass);
    if (cost < best_cost) {
      best_cost = cost;
      best.m = (int)m;
      best.r = R;
      best.pair = pair ? 1 : 0;
      best.pair_mul = mul;
      best.j0 = j0;
      best.eps = eps;
      best.est_cost = cost;
      store_tables(best, Q, R);
    }
  };

  // Search.  Evaluating a configuration costs a 128-bin convolution per look-up, so the window
  // start j0 is pre-selected with a cheap proxy of its selectivity -- the expected deficit below the
  // row maximum of the rows it covers (half rows count half), clamped -- and only the three best
  // windows of every (flavour, R) are evaluated (0.3 ms per motif instead of 4 ms: the planner is on
  // the critical path of many-motif scans).
  std::vector<double> info(m);
  for (size_t j = 0; j < m; ++j) {
    double e = 0;
    for (int a = 0; a < 4; ++a) e += 0.25 * std::min(8.0, r.rowmax[j] - std::max(r.w[j * 5 + a], r.rowmax[j] - 8.0));
    info[j] = e;
  }
  static const double CS3[] = {6.0, 12.0, CLAMP_FLOOR};
  auto best_windows = [&](bool pair, int R, int out[3]) {
    const int W = pair ? 2 * R : R, last = pair ? (int)m - (2 * R - 1) : (int)m - R;
    double sc[3] = {-1, -1, -1};
    out[0] = out[1] = out[2] = -1;
    for (int j0 = 0; j0 <= last; ++j0) {
      double v = 0;
      for (int k = 0; k < W && j0 + k < (int)m; ++k) v += (pair && (k & 1)) ? 0.5 * info[j0 + k] : info[j0 + k];
      for (int q = 0; q < 3; ++q)
        if (v > sc[q]) {
          for (int z = 2; z > q; --z) { sc[z] = sc[z - 1]; out[z] = out[z - 1]; }
          sc[q] = v; out[q] = j0;
          break;
        }
    }
  };
  if (force_mode != 2) {
    for (int R : RS_SINGLE) {
      if (R > (int)m || (force_r && R != force_r)) continue;
      int win[3];
      best_windows(false, R, win);
      for (int j0 : win)
        if (j0 >= 0)
          for (double C : CS3) consider(false, R, j0, 0, C);
    }
  }
  if (force_mode != 1) {
    for (int R : RS_PAIR) {
      if (2 * R - 1 > (int)m || (force_r && R != force_r)) continue;  // the last half row may be missing
      int win[3];
      best_windows(true, R, win);
      for (int j0 : win)
        if (j0 >= 0)
          for (uint32_t mul : {2u, 4u})
            for (double C : CS3) consider(true, R, j0, mul, C);
    }
  }
  best.ok = best_cost < EXACT_COST && best.r > 0;
  return best;
}

// ------------------------------------------------------------------ protein

namespace {
constexpr int PK = 21, PX = 20;  // protein alphabet size, code of the default symbol X

bool load_rows_protein(const float* pssm, size_t m, size_t stride, std::vector<double>& w, std::vector<double>& rowmax) {
  w.resize(m * PK);
  rowmax.resize(m);
  for (size_t j = 0; j < m; ++j) {
    double mx = -INFINITY;
    for (int a = 0; a < PK; ++a) {
      const float x = pssm[j * stride + a];
      if (std::isnan(x) || x == INFINITY) return false;
      w[j * PK + a] = x;
      if (x > mx) mx = x;
    }
    if (!std::isfinite(mx)) return false;
    rowmax[j] = mx;
  }
  return true;
}
}  // namespace

double estimate_max_score_protein(const float* pssm, size_t m, size_t stride, double n_positions) {
  std::vector<double> w, rowmax;
  if (!load_rows_protein(pssm, m, stride, w, rowmax)) return -INFINITY;
  // distribution of the score under a uniform 20-letter background on a 1/16 grid
  const double res = 1.0 / 16.0, floor_ = 24.0;
  std::vector<int> qw(m * PX);
  int span = 0;
  double maxsum = 0;
  for (size_t j = 0; j < m; ++j) {
    double mx = -INFINITY;
    for (int a = 0; a < PX; ++a) mx = std::max(mx, w[j * PK + a]);
    if (!std::isfinite(mx)) return -INFINITY;
    maxsum += mx;
    int mn = 0;
    for (int a = 0; a < PX; ++a) {
      qw[j * PX + a] = (int)std::floor((std::max(w[j * PK + a], mx - floor_) - mx) / res);
      mn = std::min(mn, qw[j * PX + a]);
    }
    span += -mn;
  }
  std::vector<double> d(span + 1, 0.0), nd(span + 1);
  d[0] = 1.0;
  for (size_t j = 0; j < m; ++j) {
    std::fill(nd.begin(), nd.end(), 0.0);
    for (int a = 0; a < PX; ++a) {
      const int sft = -qw[j * PX + a];
      for (int i = 0; i + sft <= span; ++i) nd[i + sft] += d[i] / PX;
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

ProteinPlan plan_prune_protein(const float* pssm, size_t m, size_t stride, double t_plan, bool t_is_estimate,
                               double n_positions, int force_r) {
  ProteinPlan best;
  std::memset(best.class_of, 0, sizeof(best.class_of));
  std::vector<double> w, rowmax;
  if (m < 4 || m > 64 || !load_rows_protein(pssm, m, stride, w, rowmax)) return best;
  (void)n_positions;
  if (t_is_estimate) t_plan -= 1.0;
  double abs_sum = 0, max_total = 0;
  for (size_t j = 0; j < m; ++j) {
    double am = 0;
    for (int a = 0; a < PK; ++a)
      if (std::isfinite(w[j * PK + a])) am = std::max(am, std::fabs(w[j * PK + a]));
    abs_sum += am;
    max_total += rowmax[j];
  }
  const double eps = 2.0 * (double)m * std::ldexp(1.0, -24) * abs_sum + 1e-30;
  std::vector<double> info(m);
  for (size_t j = 0; j < m; ++j) {
    double e = 0;
    for (int a = 0; a < PX; ++a) e += std::min(8.0, rowmax[j] - std::max(w[j * PK + a], rowmax[j] - 8.0)) / PX;
    info[j] = e;
  }
  static const int RS[] = {4, 6, 8, 10, 12, 16};
  static const double CS3[] = {6.0, 12.0, CLAMP_FLOOR};
  // the class-map pass moves 2 bytes per residue through HBM: ~43 issue-slot equivalents per 128 positions
  const double PREPASS_COST = 43.0;
  const double exact_cost = 17.5 * (double)((m + 3) & ~(size_t)3);  // exact kernel: ~17.5 per motif row
  double best_cost = 1e300;
  for (int R : RS) {
    if (R > (int)m || (force_r && R != force_r)) continue;
    // the three most informative windows of R rows
    int win[3] = {-1, -1, -1};
    double sc[3] = {-1, -1, -1};
    for (int j0 = 0; j0 + R <= (int)m; ++j0) {
      double v = 0;
      for (int k = 0; k < R; ++k) v += info[j0 + k];
      for (int q = 0; q < 3; ++q)
        if (v > sc[q]) {
          for (int z = 2; z > q; --z) { sc[z] = sc[z - 1]; win[z] = win[z - 1]; }
          sc[q] = v; win[q] = j0;
          break;
        }
    }
    for (int j0 : win) {
      if (j0 < 0) continue;
      // cluster the 21 residues into 8 classes on the window's (clamped) columns: greedy merging of
      // the pair of classes whose union adds the least over-estimate, weighted by class size
      std::vector<std::vector<int>> cls(PK);
      for (int a = 0; a < PK; ++a) cls[a] = {a};
      auto wc = [&](int k, int a) { return std::max(w[(j0 + k) * PK + a], rowmax[j0 + k] - 12.0); };
      auto cmax = [&](const std::vector<int>& c, int k) {
        double mx = -INFINITY;
        for (int a : c) mx = std::max(mx, wc(k, a));
        return mx;
      };
      auto weight = [&](const std::vector<int>& c) {
        int n = 0;
        for (int a : c) n += a != PX;
        return (double)n;
      };
      while (cls.size() > 8) {
        double bestc = INFINITY;
        size_t bi = 0, bj = 1;
        for (size_t i = 0; i < cls.size(); ++i)
          for (size_t j = i + 1; j < cls.size(); ++j) {
            double cost = 0;
            for (int k = 0; k < R; ++k) {
              const double mi = cmax(cls[i], k), mj = cmax(cls[j], k), mu = std::max(mi, mj);
              cost += weight(cls[i]) * (mu - mi) + weight(cls[j]) * (mu - mj);
            }
            if (cost < bestc) { bestc = cost; bi = i; bj = j; }
          }
        cls[bi].insert(cls[bi].end(), cls[bj].begin(), cls[bj].end());
        cls.erase(cls.begin() + bj);
      }
      double restmax = max_total;
      for (int k = 0; k < R; ++k) restmax -= rowmax[j0 + k];
      for (double C : CS3) {
        Unit u[16];
        for (int k = 0; k < R; ++k) {
          u[k].n = (int)cls.size();
          for (size_t e = 0; e < cls.size(); ++e) {
            double mx = -INFINITY;
            for (int a : cls[e]) mx = std::max(mx, w[(j0 + k) * PK + a]);  // true weights: the bound must hold
            u[k].v[e] = mx;
            u[k].prob[e] = weight(cls[e]) / PX;
          }
        }
        Quantised Q;
        if (!quantise_units(u, R, C, restmax, t_plan, eps, Q)) continue;
        const double cost = filter_cost(R, R, false, Q.pass) + PREPASS_COST;
        if (cost < best_cost) {
          best_cost = cost;
          best.plan.m = (int)m;
          best.plan.r = R;
          best.plan.pair = 0;
          best.plan.pair_mul = 2;
          best.plan.j0 = j0;
          best.plan.eps = eps;
          best.plan.est_cost = cost;
          store_tables(best.plan, Q, R);
          for (size_t e = 0; e < cls.size(); ++e)
            for (int a : cls[e]) best.class_of[a] = (uint8_t)e;
        }
      }
    }
  }
  best.plan.ok = best.plan.r > 0 && best_cost < exact_cost;
  return best;
}

int prune_slice_positions() { return WARP_ROWS * 32; }

// Load every instance of the filter kernel now (CUDA loads kernels lazily, at their first launch, and a load
// may synchronise with the kernels running on the device: see lmi::preload_kernels in capi.cu).
namespace {
template <int RR, bool PP>
cudaError_t preload_instances() {
  cudaFuncAttributes at;
  cudaError_t e = cudaSuccess;
#ifdef LM_PRUNE_ONLY
  if (RR == 6 && PP) e = cudaFuncGetAttributes(&at, prune_kernel<6, true, false, false, false>);
#else
#define LM_LOAD(THR, BATCH, DENSE) \
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, prune_kernel<RR, PP, THR, BATCH, DENSE>);
  LM_LOAD(false, true, false) LM_LOAD(false, false, true) LM_LOAD(false, false, false)
  LM_LOAD(true, true, false) LM_LOAD(true, false, true) LM_LOAD(true, false, false)
#undef LM_LOAD
#endif
  return e;
}
}  // namespace

cudaError_t preload_prune_kernels() {
  cudaError_t e = preload_instances<3, true>();
  if (e == cudaSuccess) e = preload_instances<4, true>();
  if (e == cudaSuccess) e = preload_instances<5, true>();
  if (e == cudaSuccess) e = preload_instances<6, true>();
  if (e == cudaSuccess) e = preload_instances<8, true>();
  if (e == cudaSuccess) e = preload_instances<4, false>();
  if (e == cudaSuccess) e = preload_instances<6, false>();
  if (e == cudaSuccess) e = preload_instances<8, false>();
  if (e == cudaSuccess) e = preload_instances<10, false>();
  if (e == cudaSuccess) e = preload_instances<12, false>();
  if (e == cudaSuccess) e = preload_instances<16, false>();
  return e;
}

cudaError_t launch_prune(bool threshold_mode, const PruneArgs& a, int sm_count, cudaStream_t stream) {
  return threshold_mode ? launch_r<true>(a, sm_count, stream) : launch_r<false>(a, sm_count, stream);
}

}  // namespace lm
