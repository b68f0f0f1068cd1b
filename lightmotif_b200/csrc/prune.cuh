// prune.cuh -- exact filter + recheck ("branch and bound") scan, DNA only.
//
// For the FUSED outputs (argmax, threshold hits) the per-position f32 scores
// never leave the chip, so only positions that can matter need the exact
// sequential f32 sum.  A cheap integer UPPER BOUND of the score is computed for
// every position (4 positions per 32-bit register, PRMT byte look-ups); only
// positions whose bound reaches the current threshold are rescored exactly
// (same sequential order as the reference => bit-identical results).  This is
// the reference's own Scanner idea (lightmotif/src/scan.rs:160-250: u8
// DiscreteMatrix prefilter, f32 recheck of candidates) rebuilt for the GPU.
#pragma once

#include "common.cuh"

namespace lm {

constexpr int PRUNE_MAX_M = 256;

// Host-built plan for one DNA PSSM.
struct PrunePlan {
  int m = 0;          // motif rows
  int r = 0;          // PRMT look-ups per position.  Single mode (pair == 0): look-up k covers motif row
                      // j0 + k, r in {4,6,8,10,12,16}.  Paired mode: look-up k covers row j0 + 2k exactly
                      // and row j0 + 2k + 1 through one class bit of its symbol, r in {3,4,5,6,8}.
  int pair = 0;
  uint32_t pair_mul = 2;  // paired mode: class(s) = bit 2 of s * pair_mul (2: {T,G}, 4: {C,G})
  int j0 = 0;
  uint32_t tab_lo[16];  // per look-up: table entries 0..3 (single: q[A,C,T,G]; paired: class 0)
  uint32_t tab_hi[16];  // per look-up: table entries 4..7 (single: q[N],0,0,0;   paired: class 1)
  int smax = 0;         // max attainable sum of q over the filter rows (<= 127)
  // UB(p) = c0 + delta * sum_q(p)  >=  real-valued score(p);  fp32 score <= real + eps
  double c0 = 0, delta = 1, eps = 0;
  double est_pass_rate = 1.0;  // under a uniform background at the planning threshold
  double est_cost = 0.0;       // modelled issue slots per quad row (128 positions)
  bool ok = false;
};

// Plan the filter for `pssm` ([m][stride] f32, DNA: columns A,C,T,G,N).
// `t_plan` is the threshold the filter is expected to run at (the user's
// threshold, or an estimate of the final maximum for argmax scans over
// `n_positions` cells).  Returns plan.ok == false when pruning is not
// expected to pay off (the caller then uses the exact-everywhere kernel).
// `force_r` != 0 forces the number of look-ups, `force_mode` 1 = single rows only, 2 = paired only.
PrunePlan plan_prune(const float* pssm, size_t m, size_t stride, double t_plan, bool t_is_estimate,
                     double n_positions, int force_r = 0, int force_mode = 0);

// Score quantile helper used for argmax planning: the score t such that the
// expected number of cells with score >= t among n uniform-random positions is ~1.
double estimate_max_score(const float* pssm, size_t m, size_t stride, double n_positions);

// Protein: the 21-letter alphabet does not fit a PRMT table, so the filter runs on a CLASS-CODED copy
// of the sequence: the residues are partitioned into 8 classes (chosen per matrix by clustering its
// columns), a look-up entry is the best weight within the class -- still an upper bound.
struct ProteinPlan {
  PrunePlan plan;        // single-row look-ups over class codes 0..7 (tab_lo = classes 0..3, tab_hi = 4..7)
  uint8_t class_of[32];  // residue code (0..20) -> class
};
ProteinPlan plan_prune_protein(const float* pssm, size_t m, size_t stride, double t_plan, bool t_is_estimate,
                               double n_positions, int force_r = 0);
double estimate_max_score_protein(const float* pssm, size_t m, size_t stride, double n_positions);

// One motif of a launch: what the kernel needs besides the sequence.  A launch may sweep MANY motifs
// (the lightmotif-cli job loop: one resident sequence, a motif file) as long as their filters share
// the template instance (look-ups per position, single / paired): work items are (motif, tile).
struct BatchMotif {
  uint32_t tab[32];            // filter tables: tab_lo[16] | tab_hi[16] of the plan
  double c0 = 0, delta = 1, eps = 0;
  int m = 0, j0 = 0, smax = 0;
  uint32_t pair_mul = 2;
  const float* table = nullptr;          // device: recheck table [m rounded up to 4][tab_stride] f32
  // device: the motif's 512-byte result block -- u64 keys[32] (per-column winners, zero-initialised),
  // first cell @256 B, hit count @264, statistics @272 (candidates, overflowing slices)
  unsigned long long* block = nullptr;
  lm_hit* hits = nullptr;                // threshold mode
  unsigned long long cap = 0;
  float thr = 0.f;
  uint32_t pad_ = 0;
};

struct PruneArgs {
  const uint8_t* seq = nullptr;  // [avail_rows][32]: what the FILTER reads (DNA symbols / protein class codes)
  // protein: the residues themselves, same shape, for the exact recheck (nullptr: recheck from the tile)
  const uint8_t* seq_exact = nullptr;
  int tab_stride = 8;            // floats per row of `table` (8 DNA, 32 protein)
  uint64_t avail_rows = 0;
  uint64_t row_begin = 0;        // first row of the scored range (positions are relative to it)
  uint64_t row_first = 0;        // first row THIS launch scans (>= row_begin; rows before were bootstrapped)
  uint64_t row_end = 0;
  uint8_t pad_sym = 4;           // default symbol of the alphabet (exact recheck past the readable extent)
  uint8_t pad_fill = 4;          // what the FILTER sees past the readable extent (DNA: N; protein: class of X)
  uint32_t one_mul = 1;          // opaque multiplier: `acc = p * one_mul + acc` keeps the byte-lane adds on the FMA pipe (IMAD)
  const float* table = nullptr;  // device: [m rounded up to 4][tab_stride] f32, pad rows +0.0
  PrunePlan plan;                // single motif: the whole plan; many motifs: r / pair select the kernel instance
  // single-motif launches: use the kernel instance whose recheck enumerates candidates across the
  // warp (worth it from ~5 candidates per slice of 3968 positions; costs the sparse case ~10 %)
  bool dense = false;
  // many motifs in one launch: device array of n_batch motifs, m_max = widest motif.  When `batch` is
  // null the launcher builds `one` from plan / table / colkeys / thr / hits / cap below.
  const BatchMotif* batch = nullptr;
  uint32_t n_batch = 1;
  int m_max = 0;
  BatchMotif one;
  // argmax mode: colkeys[32] holds the running per-column winners (read for the
  // dynamic threshold, updated with atomicMax)
  unsigned long long* colkeys = nullptr;
  // threshold mode
  float thr = 0.f;
  lm_hit* hits = nullptr;
  unsigned long long cap = 0;
  unsigned long long* count = nullptr;
  // launch-level scratch (zero-initialised): [2] the work-item counter, [3] / [4] SM cycles / ns of
  // CTA 0, [5] counter of finished producer warps (the last one publishes to the peers).  Single-motif
  // launches pass the statistics words of the motif's own block (block + 34), whose [0] / [1] are
  // the motif's candidate / overflow counts.
  unsigned long long* stats = nullptr;
  // hit coordinates (threshold mode): row = row relative to row_begin + row_add,
  // position = col * pos_rows + row  (pos_rows == 0: row_end - row_begin)
  uint64_t row_add = 0, pos_rows = 0;
  // row-sharded scans (link.world > 1): shared running maximum + publication of the result block
  PeerLink link;
};

cudaError_t launch_prune(bool threshold_mode, const PruneArgs& a, int sm_count, cudaStream_t stream);
// positions one consumer warp filters at a time (the granularity of the recheck queue)
int prune_slice_positions();
cudaError_t preload_prune_kernels();  // load every kernel instance now instead of at its first launch

}  // namespace lm
