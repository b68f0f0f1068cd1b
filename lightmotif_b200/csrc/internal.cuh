// internal.cuh -- state behind the opaque handles of the C ABI and the helpers shared by the
// translation units that implement it (capi.cu: single-GPU calls, comm.cu: multi-GPU calls).
#pragma once

#include <cstddef>
#include <cstdio>
#include <vector>

#include "common.cuh"
#include "prune.cuh"
#include "score.cuh"

// ------------------------------------------------------------------ context

struct lm_batch_cache;  // capi.cu: the last motif set of the batch calls, planned and staged
void free_batch_cache(lm_batch_cache* b);

struct lm_uploader;  // upload.cu: pinned ring, streams and events of the packed upload
void free_uploader(lm_uploader* u);

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct lm_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  float last_ms = 0.f;
  uint64_t launches = 0;
  char err[256] = {0};
  DevBuf d_in, d_out, d_table, d_table2, d_small, d_hits;  // device scratch, grown on demand
  // protein filter: class-coded copy of the last sequence scanned (one-entry cache keyed by the
  // sequence buffer and the residue -> class map)
  DevBuf d_class;
  uint64_t class_seq_id = 0;  // lm_seq::id the copy was made from (0 = none)
  size_t class_rows = 0;
  uint8_t class_lut[32] = {0};
  std::vector<uint8_t> pplan_key;
  lm::ProteinPlan pplan_cache;
  int opt_prune = 1;                      // LM_OPT_PRUNE: 0 never, 1 auto, 2 whenever a plan exists
  int opt_filter_rows = 0;                // LM_OPT_FILTER_ROWS: 0 = planner decides
  int opt_filter_mode = 0;                // LM_OPT_FILTER_MODE: 0 = planner decides, 1 = single rows, 2 = paired rows
  int opt_class_cache = 1;                // LM_OPT_CLASS_CACHE: reuse the class-coded copy of a protein sequence
  int opt_reserve_sms = 0;                // LM_OPT_RESERVE_SMS: SMs left free for concurrent (NCCL) kernels
  size_t opt_boot_rows = 262144;          // LM_OPT_BOOT_ROWS: rows scored exactly first to seed the running maximum
  uint64_t stats[4] = {0, 0, 0, 0};       // last scoring call: candidates, overflow tiles, pruned?, filter rows
  float sm_mhz = 0.f;                     // SM clock the last filtered scan ran at (cycle counter / global timer)
  // the last matrix uploaded to d_table / d_table2 and the last filter plan: a
  // caller scanning many sequences (or benchmarking) with one PSSM pays for the
  // host-side set-up once
  std::vector<uint8_t> table_key, table8_key, plan_key;
  lm::PrunePlan plan_cache;
  // candidates per slice observed by the last filtered single-motif scan, and the plan it ran with:
  // the next scan with that plan picks the kernel instance accordingly (prune.cuh PruneArgs::dense)
  std::vector<uint8_t> density_key;
  double density = -1.0;
  // argmax scans plan their filter for the maximum a uniform-random sequence of that length would
  // reach; a sequence whose real maximum is far lower (a real genome without a strong site) floods
  // that filter with candidates.  When a scan observes that, the next scan with the same matrix is
  // planned for the maximum actually found (once per matrix).
  std::vector<uint8_t> replan_key;   // plan key the observation belongs to
  double replan_t = 0.0;
  uint64_t replan_seq = 0;           // lm_seq::id the observation was made on
  int replan_state = 0;              // 0 none, 1 pending (use replan_t at the next plan), 2 done
  // in-flight fused argmax calls (launch / finish split, LM_SLOTS slots): result blocks
  // live at d_small + slot * 1024 and are copied to pinned host memory
  struct Slot {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;
    void* h = nullptr;  // pinned SmallBlock
    int state = 0;      // 0 idle, 1 in flight, 2 in flight but empty (LM_EMPTY)
    size_t n_rows = 0;
    uint64_t seq_id = 0;
    bool pruned = false;
    int r = 0;
  } slots[LM_SLOTS];
  lm_batch_cache* batch = nullptr;
  bool batch_staged = false;                      // d_table still holds that set's tables
  void* h_pin = nullptr;                          // pinned host staging
  size_t h_pin_bytes = 0;
  // large striped DNA matrices from host memory (lm_seq_from_striped): two-ended upload, upload.cu
  lm_uploader* uploader = nullptr;
  int opt_upload = 1;                             // LM_OPT_UPLOAD: 0 plain copy, 1 adaptive, 2 packed chunks only, 3 / 4 two-ended (gaps only / always)
  int opt_upload_threads = 0;                     // LM_OPT_UPLOAD_THREADS: host threads that pack (0 = all but one)
  long long opt_upload_min_bytes = 64ll << 20;    // LM_OPT_UPLOAD_MIN_BYTES: smaller matrices take the plain copy
  uint64_t upload_stats[4] = {0, 0, 0, 0};        // last upload: bytes over PCIe, packed chunks, plain chunks, threads
  // LM_OPT_UPLOAD = 1: bytes/s the two-ended upload and the plain copy achieved on this context's last large
  // uploads (0 = not tried yet); the faster one is used, the other re-tried every 32nd upload
  double upload_rate_packed = 0.0, upload_rate_plain = 0.0;
  uint32_t upload_calls = 0;
  bool upload_ring_created = false;               // the last two-ended upload also allocated the ring: its time is not a rate
};

struct lm_seq {
  int alphabet = 0;
  int device = 0;
  uint8_t* data = nullptr;  // [alloc_rows][32]
  bool owned = true;
  size_t rows = 0, wrap = 0, length = 0, alloc_rows = 0;
  // row slab of a longer sequence (lm_seq_synthetic with a sub-range, lm_comm_shard_*): this handle
  // holds rows [row_offset, row_offset + rows) of a striped matrix of `rows_global` rows; its halo rows
  // are the NEXT rows of that matrix, not wrap rows.  rows_global == rows for a whole sequence.
  uint64_t row_offset = 0, rows_global = 0;
  uint64_t id = 0;  // unique per handle (cache keys must not rely on the buffer address)
  bool is_slab() const { return rows_global != rows; }
};


// small block layout: [0,256) column keys, [256] first cell, [264] hit count, [272..320) filter stats
struct ColKeys {
  unsigned long long keys[32];
  uint32_t first_cell;
};
struct SmallBlock {
  ColKeys ck;                    // 260 bytes (+4 pad)
  unsigned long long count;      // @264
  unsigned long long stats[6];   // @272: candidates, overflow tiles, tile counter, SM cycles, ns (CTA 0's producer), finished warps
};
static_assert(offsetof(SmallBlock, count) == 264 && offsetof(SmallBlock, stats) == 272, "small block layout");

namespace lm {
// hand `small` (a SmallBlock in this rank's HBM) to every rank of a sharded scan: one tiny kernel,
// used when the scan kernel could not do it itself (exact-only scans, empty ranges); comm.cu
cudaError_t launch_publish(const PeerLink& link, const unsigned long long* small, cudaStream_t stream);
}  // namespace lm

// helpers of capi.cu shared with comm.cu
namespace lmi {

int fail(lm_ctx* ctx, cudaError_t e, const char* what);
int bind(lm_ctx* ctx);
int ensure(lm_ctx* ctx, DevBuf& b, size_t bytes);
uint8_t pad_symbol(int alphabet);
void decode_colwins(const unsigned long long* keys, bool is_u8, lm_colwin* out);
int apply_rule(const lm_colwin* cw, float first_cell, int rule, bool is_u8, uint64_t total_rows, uint64_t row_base,
               lm_hit* best);
float bits_to_score(uint32_t bits, bool is_u8);
void sort_hits(lm_hit* hits, size_t n);
int seq_alloc(lm_ctx* ctx, int alphabet, size_t rows, size_t wrap, size_t length, lm_seq** out);

// Enqueue a fused score+argmax scan of rows [row_begin, row_end) of `seq` into launch slot `slot`
// (lm_seq_score_argmax_f32_launch).  `link` (nullable) couples the scan to the other ranks of a
// row-sharded scan: shared running maximum + publication of the result block (prune.cuh PeerLink).
int argmax_launch(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride, size_t row_begin,
                  size_t row_end, int slot, const lm::PeerLink* link);
// Enqueue a fused score+threshold scan; hits (row-major order NOT applied) go to `d_hits` (device,
// capacity `cap`), their number to the result block of `slot`.  Rows / positions of the hits are
// reported in the coordinates of the global matrix when `seq` is a slab.
// `small_off` = byte offset of the 512-byte result block inside ctx->d_small.
int threshold_launch(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride, size_t row_begin,
                     size_t row_end, float t, lm_hit* d_hits, size_t cap, size_t small_off, const lm::PeerLink* link,
                     lm::PrunePlan* plan_out);
// upload.cu
bool upload_wants_packed(lm_ctx* ctx, int alphabet, size_t stride, size_t bytes, const void* src);
void upload_note_rate(lm_ctx* ctx, bool packed, size_t bytes, double seconds);
int upload_striped_dna(lm_ctx* ctx, uint8_t* dst, const uint8_t* src, size_t bytes);
int check_score_args(size_t m, size_t rows, size_t wrap, size_t length, size_t row_begin, size_t row_end);

}  // namespace lmi

#define LM_TRY(expr)                                          \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return lmi::fail(ctx, _e, #expr);  \
  } while (0)

#define LM_CHECK(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s < 0) return _s;      \
  } while (0)
