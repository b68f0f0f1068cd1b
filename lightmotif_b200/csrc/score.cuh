// score.cuh -- launch interface of the exact scoring kernels (score.cu).
#pragma once

#include "common.cuh"

namespace lm {

enum ScoreMode : int {
  MODE_STORE = 1,      // materialise StripedScores ([rows][stride])
  MODE_ARGMAX = 2,     // fused per-column winners
  MODE_THRESHOLD = 4,  // fused hit list
};

// Look-up table layout used by the kernels: [m_pad][KP] 32-bit words, KP = 16
// (DNA) or 32 (protein) so that a group of 4 motif rows is 256 / 512 bytes and
// its shared-memory address has a zero low byte (see score.cu).  m_pad is m
// rounded up to a multiple of 4; the extra rows are +0.0 (f32) / 0 (u8),
// which leaves a sequential sum bit-identical.
__host__ __device__ inline int table_kp(int alphabet) { return alphabet == LM_DNA ? 16 : 32; }
inline int table_mpad(size_t m) { return (int)((m + 3) / 4 * 4); }

struct ScoreArgs {
  // sequence: [avail_rows][32] bytes on the device
  const uint8_t* seq = nullptr;
  uint64_t avail_rows = 0;  // rows that may be read from `seq`
  uint64_t row_begin = 0, row_end = 0;
  uint8_t pad_sym = 0;
  // motifs: tables[table_off[q] ...] holds motif q as [m_pad[q]][KP] words
  const uint32_t* tables = nullptr;
  const uint64_t* table_off = nullptr;  // device, [n_motifs] (nullptr: single motif at 0)
  const int* m_pads = nullptr;          // device, [n_motifs] (nullptr: use m_pad)
  int m_pad = 0;                        // single motif / max over the batch
  int n_motifs = 1;
  uint32_t motif_chunks = 1;            // set by the launcher (see score.cu)
  // MODE_STORE
  void* out = nullptr;
  uint64_t out_stride = 0;  // elements
  // MODE_ARGMAX: motif q's 32 packed keys (score key << 32 | row, zero-initialised) start at
  // colkeys[q * colkeys_stride]; first_cell[q * first_cell_stride] = raw 32-bit score of cell (row_begin, 0)
  unsigned long long* colkeys = nullptr;
  uint32_t* first_cell = nullptr;
  uint64_t colkeys_stride = 32, first_cell_stride = 1;
  // MODE_THRESHOLD
  uint32_t thr_bits = 0;  // f32 bits or u8 value
  lm_hit* hits = nullptr;
  unsigned long long cap = 0;
  unsigned long long* count = nullptr;
  // hit coordinates: row = row relative to row_begin + row_add, position = col * pos_rows + row
  // (pos_rows == 0: row_end - row_begin)
  uint64_t row_add = 0, pos_rows = 0;
};

// Enqueue the scoring kernel on `stream`.  is_u8 selects the saturating u8
// pipeline (tables hold u32-expanded DiscreteMatrix cells).  Returns a CUDA
// error code.
cudaError_t preload_score_kernels();
cudaError_t launch_score(int alphabet, int mode, bool is_u8, const ScoreArgs& a, int sm_count,
                         cudaStream_t stream);

}  // namespace lm
