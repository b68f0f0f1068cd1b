// capi.cu -- the extern "C" boundary (include/lightmotif_b200.h): contexts,
// device-resident sequence handles, host<->device plumbing and the tiny host
// finishing steps (tie rules over 32 per-column winners, hit ordering).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "internal.cuh"
#include "prune.cuh"
#include "reduce.cuh"
#include "score.cuh"
#include "seqops.cuh"

using namespace lm;
using namespace lmi;

#ifndef LM_BATCH_BOOT
#define LM_BATCH_BOOT 4096  // rows scored exactly to seed the running maxima of a many-motif argmax call
#endif

namespace {
constexpr size_t SLACK_ROWS = 64;      // readable rows kept behind rows + wrap
constexpr size_t SMALL_BYTES = 1 << 20;  // keys / counters block
// d_small: launch slot s owns [s * 1024, s * 1024 + 512); the synchronous calls use the block below
constexpr size_t SYNC_SMALL_OFF = (size_t)LM_SLOTS * 1024;

// Some kernel instances keep a few bytes per thread on the stack.  The driver reserves a device's local
// memory lazily, at the first launch that needs it -- and that launch synchronises with every kernel running
// on the device.  A row-sharded scan keeps a kernel spinning until the OTHER ranks have published
// (gather_wait_kernel): when two ranks share a device (tests, lm_comm_init with a device listed twice) the
// late rank's first launch would wait for the kernel that waits for it.  So the reservation is made here, once
// per context, before anything else runs.
__global__ void reserve_local_memory_kernel(int* out, int n) {
  int buf[64];  // (indexed dynamically: lives on the stack)
  for (int i = 0; i < 64; ++i) buf[(i * n) & 63] = i;
  if (out) out[0] = buf[(n + threadIdx.x) & 63];
}
}  // namespace

namespace lmi {

int fail(lm_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();  // clear sticky-free errors
  return e == cudaErrorMemoryAllocation ? LM_ERR_OOM : LM_ERR_CUDA;
}

int ensure(lm_ctx* ctx, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return LM_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  size_t want = std::max<size_t>(bytes, 4096);
  LM_TRY(cudaMalloc(&b.p, want));
  b.bytes = want;
  return LM_OK;
}

int ensure_pinned(lm_ctx* ctx, size_t bytes) {
  if (ctx->h_pin_bytes >= bytes && ctx->h_pin) return LM_OK;
  if (ctx->h_pin) { cudaFreeHost(ctx->h_pin); ctx->h_pin = nullptr; ctx->h_pin_bytes = 0; }
  size_t want = std::max<size_t>(bytes, 1 << 16);
  LM_TRY(cudaHostAlloc(&ctx->h_pin, want, cudaHostAllocDefault));
  ctx->h_pin_bytes = want;
  return LM_OK;
}

int bind(lm_ctx* ctx) {
  if (!ctx) return LM_ERR_INVALID_ARG;
  LM_TRY(cudaSetDevice(ctx->device));
  return LM_OK;
}

void timer_begin(lm_ctx* ctx) { cudaEventRecord(ctx->ev0, ctx->stream); }
void timer_end(lm_ctx* ctx) {
  cudaEventRecord(ctx->ev1, ctx->stream);
  ctx->timed = true;
}

// copy `rows` rows of `width` bytes; a single flat copy when both sides are dense
cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                      cudaMemcpyKind kind, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, stream);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, stream);
}

bool valid_alphabet(int a) { return a == LM_DNA || a == LM_PROTEIN; }
uint8_t pad_symbol(int a) { return (uint8_t)(alphabet_k(a) - 1); }

// Build the kernel look-up table [m_pad][KP] (32-bit words) from a PSSM.
// Columns >= K are zero-filled and never indexed; rows >= m are +0.0 / 0.
void build_table_f32(int alphabet, const float* pssm, size_t m, size_t stride, uint32_t* dst) {
  const int kp = table_kp(alphabet), k = alphabet_k(alphabet), mp = table_mpad(m);
  std::memset(dst, 0, (size_t)mp * kp * 4);
  for (size_t j = 0; j < m; ++j)
    for (int s = 0; s < k; ++s) std::memcpy(&dst[j * kp + s], &pssm[j * stride + s], 4);
}
void build_table_u8(int alphabet, const uint8_t* dm, size_t m, size_t stride, uint32_t* dst) {
  const int kp = table_kp(alphabet), k = alphabet_k(alphabet), mp = table_mpad(m);
  std::memset(dst, 0, (size_t)mp * kp * 4);
  for (size_t j = 0; j < m; ++j)
    for (int s = 0; s < k; ++s) dst[j * kp + s] = dm[j * stride + s];
}

void decode_colwins(const unsigned long long* keys, bool is_u8, lm_colwin* out) {
  for (int c = 0; c < 32; ++c) {
    unsigned long long k = keys[c];
    out[c].valid = k != 0;
    out[c].row = (uint32_t)(k & 0xffffffffull);
    uint32_t sk = (uint32_t)(k >> 32);
    out[c].score = k == 0 ? 0.f : (is_u8 ? (float)(sk - 1u) : key_to_f32(sk));
  }
}

// Apply a reference back-end's tie rule to 32 per-column winners (each the
// column maximum with the LAST row attaining it).  first_cell = score of
// cell (0, 0) of the scored range.  Returns LM_EMPTY when nothing is valid.
int apply_rule(const lm_colwin* cw, float first_cell, int rule, bool is_u8, uint64_t total_rows,
               uint64_t row_base, lm_hit* best) {
  int bc = -1;
  uint64_t br = 0;
  float bs = 0.f;
  bool any = false;
  for (int c = 0; c < 32; ++c) any = any || cw[c].valid;
  if (!any) return LM_EMPTY;
  if (rule == LM_RULE_GENERIC) {
    // pli/mod.rs:140-152: row-major scan with `>=`: highest score, then the
    // largest row, then the largest column
    for (int c = 0; c < 32; ++c) {
      if (!cw[c].valid) continue;
      if (bc < 0 || cw[c].score > bs || (cw[c].score == bs && cw[c].row >= br)) {
        bc = c; br = cw[c].row; bs = cw[c].score;
      }
    }
  } else if (rule == LM_RULE_AVX2 && !is_u8) {
    // avx2.rs:411-421: start from cell (0,0), strict `>` over columns 0..31
    bc = 0; br = row_base; bs = first_cell;
    for (int c = 0; c < 32; ++c)
      if (cw[c].valid && cw[c].score > bs) { bc = c; br = cw[c].row; bs = cw[c].score; }
  } else if (rule == LM_RULE_SSE2 && !is_u8) {
    // sse2.rs:246-258: start from (0,0), `>=` over columns => last column wins
    bc = 0; br = row_base; bs = first_cell;
    for (int c = 0; c < 32; ++c)
      if (cw[c].valid && cw[c].score >= bs) { bc = c; br = cw[c].row; bs = cw[c].score; }
  } else {
    // u8: avx2.rs:525-528 Iterator::max_by_key => last column among equal maxima;
    // SSE2 has no u8 argmax and falls back to the generic scan
    if (rule == LM_RULE_SSE2) return apply_rule(cw, first_cell, LM_RULE_GENERIC, is_u8, total_rows, row_base, best);
    for (int c = 0; c < 32; ++c)
      if (cw[c].valid && (bc < 0 || cw[c].score >= bs)) { bc = c; br = cw[c].row; bs = cw[c].score; }
  }
  best->row = br;
  best->col = (uint32_t)bc;
  best->score = bs;
  best->position = (uint64_t)bc * total_rows + br;
  return LM_OK;
}

float bits_to_score(uint32_t bits, bool is_u8) {
  if (is_u8) return (float)bits;
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

void sort_hits(lm_hit* hits, size_t n) {
  std::sort(hits, hits + n, [](const lm_hit& a, const lm_hit& b) {
    return a.row != b.row ? a.row < b.row : a.col < b.col;
  });
}

// shared argument validation for scoring calls; returns LM_OK, LM_EMPTY
// (nothing to score: the reference resizes to (0, 0)) or an error
int check_score_args(size_t m, size_t rows, size_t wrap, size_t length, size_t row_begin,
                     size_t row_end) {
  if (m == 0) return LM_ERR_INVALID_ARG;
  if (wrap + 1 < m) return LM_ERR_WRAP;             // avx2.rs:832-837
  if (length < m || row_end <= row_begin) return LM_EMPTY;  // pli/mod.rs:85-88
  if (row_end > rows) return LM_ERR_INVALID_ARG;
  if (rows + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;  // 32-bit rows in packed keys
  return LM_OK;
}

// canonical byte key of a matrix: header + the K valid cells of every row
std::vector<uint8_t> matrix_key(int tag, int alphabet, const void* pssm, size_t m, size_t stride, size_t esz) {
  const size_t k = (size_t)alphabet_k(alphabet);
  std::vector<uint8_t> key(16 + m * k * esz);
  int64_t hdr[2] = {((int64_t)tag << 32) | (int64_t)alphabet, (int64_t)m};
  std::memcpy(key.data(), hdr, 16);
  for (size_t j = 0; j < m; ++j)
    std::memcpy(key.data() + 16 + j * k * esz, (const uint8_t*)pssm + j * stride * esz, k * esz);
  return key;
}

int upload_table(lm_ctx* ctx, int alphabet, const void* pssm, size_t m, size_t stride, bool is_u8) {
  std::vector<uint8_t> key = matrix_key(is_u8 ? 2 : 1, alphabet, pssm, m, stride, is_u8 ? 1 : 4);
  if (key == ctx->table_key && ctx->d_table.p) return LM_OK;  // already resident
  ctx->table_key.clear();
  ctx->batch_staged = false;  // d_table is about to hold this matrix instead of a batch's tables
  const size_t bytes = (size_t)table_mpad(m) * table_kp(alphabet) * 4;
  LM_CHECK(ensure_pinned(ctx, bytes));
  LM_CHECK(ensure(ctx, ctx->d_table, bytes));
  if (is_u8) build_table_u8(alphabet, (const uint8_t*)pssm, m, stride, (uint32_t*)ctx->h_pin);
  else build_table_f32(alphabet, (const float*)pssm, m, stride, (uint32_t*)ctx->h_pin);
  LM_TRY(cudaMemcpyAsync(ctx->d_table.p, ctx->h_pin, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // the pinned buffer is reused by the next call: make sure the copy has left it
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->table_key.swap(key);
  return LM_OK;
}

int limits_ok(int alphabet, size_t m) {
  // shared-memory budget of score.cu (table + tile buffers <= 200 KB)
  size_t mp = table_mpad(m);
  size_t total = mp * table_kp(alphabet) * 4 + (128 + mp) * (32 * 3 + 4) + 4096 + mp * 128;  // (+ pair table)
  return total <= 200 * 1024 ? LM_OK : LM_ERR_UNSUPPORTED;
}

// core: score rows [row_begin, row_end) of device matrix `seq` with the table
// already uploaded to ctx->d_table
int run_score(lm_ctx* ctx, int alphabet, bool is_u8, int mode, const uint8_t* seq, size_t avail_rows,
              size_t m, size_t row_begin, size_t row_end, void* d_out, size_t out_stride,
              float thr_f32, uint8_t thr_u8, lm_hit* d_hits, size_t cap, bool standalone = true,
              size_t small_off = 0) {
  ScoreArgs a;
  a.seq = seq;
  a.avail_rows = avail_rows;
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.pad_sym = pad_symbol(alphabet);
  a.tables = (const uint32_t*)ctx->d_table.p;
  a.m_pad = table_mpad(m);
  a.n_motifs = 1;
  a.out = d_out;
  a.out_stride = out_stride;
  uint8_t* small = (uint8_t*)ctx->d_small.p + small_off;
  a.colkeys = (unsigned long long*)small;
  a.first_cell = (uint32_t*)(small + 256);
  a.count = (unsigned long long*)(small + 264);
  if (is_u8) a.thr_bits = thr_u8; else std::memcpy(&a.thr_bits, &thr_f32, 4);
  a.hits = d_hits;
  a.cap = cap;
  if (standalone) {
    LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
    ctx->stats[0] = ctx->stats[1] = ctx->stats[2] = ctx->stats[3] = 0;
    timer_begin(ctx);
  }
  LM_TRY(launch_score(alphabet, mode, is_u8, a, std::max(1, ctx->sm_count - ctx->opt_reserve_sms), ctx->stream));
  if (standalone) timer_end(ctx);
  ctx->launches += 1;
  return LM_OK;
}

// f32 table [m_pad][8] (A,C,T,G,N,0,0,0) for the exact recheck of the pruned scan
int upload_table8(lm_ctx* ctx, const float* pssm, size_t m, size_t stride) {
  std::vector<uint8_t> key = matrix_key(3, LM_DNA, pssm, m, stride, 4);
  if (key == ctx->table8_key && ctx->d_table2.p) return LM_OK;
  ctx->table8_key.clear();  // set again only once the new table is on the device
  const size_t m_pad = (m + 3) & ~(size_t)3;  // rows >= m are +0.0 (the recheck loop is unrolled by 4)
  const size_t bytes = m_pad * 8 * 4;
  LM_CHECK(ensure_pinned(ctx, bytes));
  LM_CHECK(ensure(ctx, ctx->d_table2, bytes));
  float* h = (float*)ctx->h_pin;
  std::memset(h, 0, bytes);
  for (size_t j = 0; j < m; ++j)
    for (int a = 0; a < 8; ++a) h[j * 8 + a] = a < 5 ? pssm[j * stride + a] : 0.f;
  LM_TRY(cudaMemcpyAsync(ctx->d_table2.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->table8_key.swap(key);
  return LM_OK;
}


// plan_prune with a one-entry cache (same matrix, same threshold situation)
PrunePlan cached_plan(lm_ctx* ctx, const float* pssm, size_t m, size_t stride, bool argmax_mode, double t,
                      double npos, uint64_t seq_id) {
  std::vector<uint8_t> key = matrix_key(argmax_mode ? 4 : 5, LM_DNA, pssm, m, stride, 4);
  if (ctx->replan_state && seq_id != ctx->replan_seq) {  // the observation was about another sequence
    if (ctx->replan_state == 2 && key == ctx->plan_key) ctx->plan_key.clear();
    ctx->replan_state = 0;
  }
  double extra[4] = {argmax_mode ? 0.0 : t, npos, (double)ctx->opt_filter_rows, (double)ctx->opt_filter_mode};
  key.insert(key.end(), (const uint8_t*)extra, (const uint8_t*)extra + sizeof(extra));
  const bool replan = argmax_mode && ctx->replan_state == 1 && key == ctx->replan_key;
  if (key == ctx->plan_key && !replan) return ctx->plan_cache;
  if (!replan && key != ctx->replan_key) ctx->replan_state = 0;  // another matrix: forget the observation
  const double t_plan = replan ? ctx->replan_t : estimate_max_score(pssm, m, stride, npos);
  PrunePlan plan = argmax_mode
      ? plan_prune(pssm, m, stride, t_plan, true, npos, ctx->opt_filter_rows, ctx->opt_filter_mode)
      : plan_prune(pssm, m, stride, t, false, npos, ctx->opt_filter_rows, ctx->opt_filter_mode);
  if (replan) {
    ctx->replan_state = 2;
    ctx->density = -1.0;  // the observed density belonged to the old plan
  }
  ctx->plan_key.swap(key);
  ctx->plan_cache = plan;
  return plan;
}

// Which instance of the filter kernel a single-motif scan runs: the one whose recheck enumerates
// candidates across the warp pays off from ~5 queue entries per slice.  Predicted from the plan's pass
// rate under a uniform background -- or, when the previous scan ran with the same plan, from the
// density it actually met (a real genome at a modest maximum is far denser than the model says).
bool want_dense(lm_ctx* ctx, const PrunePlan& plan) {
  if (ctx->density >= 0.0 && ctx->density_key == ctx->plan_key) return ctx->density > 4.0;
  return plan.est_pass_rate * prune_slice_positions() > 4.0;
}
void note_density(lm_ctx* ctx, uint64_t candidates, uint64_t rows) {
  ctx->density_key = ctx->plan_key;
  ctx->density = rows ? (double)candidates * prune_slice_positions() / ((double)rows * 32.0) : -1.0;
}

// protein: plan with a one-entry cache, recheck table [m_pad][32], class-coded copy of the sequence
ProteinPlan cached_protein_plan(lm_ctx* ctx, const float* pssm, size_t m, size_t stride, bool argmax_mode, double t,
                                double npos) {
  std::vector<uint8_t> key = matrix_key(argmax_mode ? 6 : 7, LM_PROTEIN, pssm, m, stride, 4);
  double extra[3] = {argmax_mode ? 0.0 : t, npos, (double)ctx->opt_filter_rows};
  key.insert(key.end(), (const uint8_t*)extra, (const uint8_t*)extra + sizeof(extra));
  if (key == ctx->pplan_key) return ctx->pplan_cache;
  ProteinPlan plan = argmax_mode ? plan_prune_protein(pssm, m, stride, estimate_max_score_protein(pssm, m, stride, npos),
                                                      true, npos, ctx->opt_filter_rows)
                                 : plan_prune_protein(pssm, m, stride, t, false, npos, ctx->opt_filter_rows);
  ctx->pplan_key.swap(key);
  ctx->pplan_cache = plan;
  return plan;
}

int upload_table32(lm_ctx* ctx, const float* pssm, size_t m, size_t stride) {
  std::vector<uint8_t> key = matrix_key(8, LM_PROTEIN, pssm, m, stride, 4);
  if (key == ctx->table8_key && ctx->d_table2.p) return LM_OK;
  ctx->table8_key.clear();
  const size_t m_pad = (m + 3) & ~(size_t)3;
  const size_t bytes = m_pad * 32 * 4;
  LM_CHECK(ensure_pinned(ctx, bytes));
  LM_CHECK(ensure(ctx, ctx->d_table2, bytes));
  float* h = (float*)ctx->h_pin;
  std::memset(h, 0, bytes);
  for (size_t j = 0; j < m; ++j)
    for (int a = 0; a < 21; ++a) h[j * 32 + a] = pssm[j * stride + a];
  LM_TRY(cudaMemcpyAsync(ctx->d_table2.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->table8_key.swap(key);
  return LM_OK;
}

int ensure_class_copy(lm_ctx* ctx, const lm_seq* seq, const uint8_t* class_of) {
  if (ctx->opt_class_cache && seq->owned && ctx->class_seq_id == seq->id && ctx->class_rows == seq->alloc_rows &&
      ctx->d_class.p && std::memcmp(ctx->class_lut, class_of, 32) == 0)
    return LM_OK;  // same handle (ids are never reused), same rows, same classes: the copy is still valid
  LM_CHECK(ensure(ctx, ctx->d_class, seq->alloc_rows * 32));
  LM_TRY(launch_classmap(seq->data, seq->alloc_rows * 32, (uint8_t*)ctx->d_class.p, class_of, ctx->sm_count,
                         ctx->stream));
  ctx->launches += 1;
  ctx->class_seq_id = seq->id;
  ctx->class_rows = seq->alloc_rows;
  std::memcpy(ctx->class_lut, class_of, 32);
  return LM_OK;
}

int fetch_stats(lm_ctx* ctx, const PrunePlan& plan, size_t small_off = 0) {
  unsigned long long st[2] = {0, 0};
  LM_TRY(cudaMemcpyAsync(st, (uint8_t*)ctx->d_small.p + small_off + 272, 16, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->stats[0] = st[0];
  ctx->stats[1] = st[1];
  ctx->stats[2] = 1;
  ctx->stats[3] = (uint64_t)plan.r | ((uint64_t)plan.pair << 8);
  return LM_OK;
}

int fetch_colkeys(lm_ctx* ctx, ColKeys* out, const PrunePlan* plan = nullptr) {
  SmallBlock sb;
  LM_TRY(cudaMemcpyAsync(&sb, ctx->d_small.p, sizeof(SmallBlock), cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  *out = sb.ck;
  if (plan) {
    ctx->stats[0] = sb.stats[0];
    ctx->stats[1] = sb.stats[1];
    ctx->stats[2] = 1;
    ctx->stats[3] = (uint64_t)plan->r | ((uint64_t)plan->pair << 8);
  }
  return LM_OK;
}

int seq_alloc(lm_ctx* ctx, int alphabet, size_t rows, size_t wrap, size_t length, lm_seq** out) {
  lm_seq* s = new (std::nothrow) lm_seq();
  if (!s) return LM_ERR_OOM;
  s->alphabet = alphabet;
  s->device = ctx->device;
  s->rows = rows;
  s->wrap = wrap;
  s->length = length;
  s->alloc_rows = rows + wrap + SLACK_ROWS;
  s->rows_global = rows;
  static std::atomic<uint64_t> next_id{1};
  s->id = next_id.fetch_add(1);
  // stream-ordered allocation from the device's default pool (kept warm: see lm_ctx_create): a caller
  // that uploads, scans and frees sequence after sequence does not pay the OS for every cycle
  cudaError_t e = cudaMallocAsync((void**)&s->data, s->alloc_rows * 32, ctx->stream);
  if (e != cudaSuccess) { delete s; return fail(ctx, e, "cudaMalloc(sequence)"); }
  *out = s;
  return LM_OK;
}

}  // namespace

// --------------------------------------------------------------- lifecycle

extern "C" {

int lm_abi_version(void) { return LM_ABI_VERSION; }

const char* lm_strerror(int status) {
  switch (status) {
    case LM_OK: return "ok";
    case LM_EMPTY: return "empty (None)";
    case LM_ERR_INVALID_ARG: return "invalid argument";
    case LM_ERR_INVALID_SYMBOL: return "invalid symbol in sequence";
    case LM_ERR_WRAP: return "not enough wrapping rows for motif";
    case LM_ERR_CUDA: return "CUDA error";
    case LM_ERR_OOM: return "out of memory";
    case LM_ERR_UNSUPPORTED: return "unsupported backend or size";
    case LM_ERR_OVERFLOW: return "output buffer too small";
    default: return "unknown status";
  }
}

const char* lm_ctx_last_error(const lm_ctx* ctx) { return ctx ? ctx->err : ""; }

int lm_ctx_create_on_stream(int device, void* cuda_stream, lm_ctx** out) {
  if (!out) return LM_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return LM_ERR_UNSUPPORTED;  // err.rs:34 UnsupportedBackend
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LM_ERR_UNSUPPORTED;
  if (prop.major < 10) return LM_ERR_UNSUPPORTED;  // kernels are built for sm_100a only
  lm_ctx* ctx = new (std::nothrow) lm_ctx();
  if (!ctx) return LM_ERR_OOM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  auto bail = [&](cudaError_t e, const char* w) { int s = fail(ctx, e, w); lm_ctx_destroy(ctx); return s; };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bail(e, "cudaSetDevice");
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail(e, "cudaStreamCreate");
    ctx->own_stream = true;
  }
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
  {  // keep freed sequence buffers in the device's default memory pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  int s = ensure(ctx, ctx->d_small, SMALL_BYTES);
  if (s < 0) { lm_ctx_destroy(ctx); return s; }
  // (the per-motif tables never regrow in steady state: a cudaFree synchronises the device -- see the note on
  // reserve_local_memory_kernel -- and the largest motif the kernels accept needs < 128 KB)
  if ((s = ensure(ctx, ctx->d_table, 128 << 10)) < 0 || (s = ensure(ctx, ctx->d_table2, 128 << 10)) < 0 ||
      (s = ensure_pinned(ctx, 128 << 10)) < 0) {
    lm_ctx_destroy(ctx);
    return s;
  }
  reserve_local_memory_kernel<<<1, 32, 0, ctx->stream>>>((int*)ctx->d_small.p, 1);
  if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "reserve_local_memory_kernel");
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "reserve_local_memory_kernel");
  for (auto& sl : ctx->slots) {
    if ((e = cudaEventCreate(&sl.ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&sl.ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaHostAlloc(&sl.h, 512, cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc");
  }
  *out = ctx;
  return LM_OK;
}

int lm_ctx_create(int device, lm_ctx** out) { return lm_ctx_create_on_stream(device, nullptr, out); }

void lm_ctx_destroy(lm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (DevBuf* b : {&ctx->d_in, &ctx->d_out, &ctx->d_table, &ctx->d_table2, &ctx->d_small, &ctx->d_hits, &ctx->d_class})
    if (b->p) cudaFree(b->p);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  free_uploader(ctx->uploader);
  free_batch_cache(ctx->batch);
  for (auto& sl : ctx->slots) {
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    if (sl.ev1) cudaEventDestroy(sl.ev1);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.h) cudaFreeHost(sl.h);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void* lm_ctx_stream(lm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int lm_ctx_synchronize(lm_ctx* ctx) {
  LM_CHECK(bind(ctx));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  return LM_OK;
}

uint64_t lm_ctx_kernel_launches(const lm_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lm_ctx_set_option(lm_ctx* ctx, int option, int value) {
  if (!ctx) return LM_ERR_INVALID_ARG;
  if (option == LM_OPT_PRUNE && value >= 0 && value <= 2) {
    ctx->opt_prune = value;
    return LM_OK;
  }
  if (option == LM_OPT_RESERVE_SMS && value >= 0 && value < 64) {
    ctx->opt_reserve_sms = value;
    return LM_OK;
  }
  if (option == LM_OPT_FILTER_ROWS && value >= 0 && value <= 16) {
    ctx->opt_filter_rows = value;
    return LM_OK;
  }
  if (option == LM_OPT_FILTER_MODE && value >= 0 && value <= 2) {
    ctx->opt_filter_mode = value;
    return LM_OK;
  }
  if (option == LM_OPT_CLASS_CACHE && (value == 0 || value == 1)) {
    ctx->opt_class_cache = value;
    return LM_OK;
  }
  if (option == LM_OPT_UPLOAD && value >= 0 && value <= 4) {
    ctx->opt_upload = value;
    ctx->upload_rate_packed = ctx->upload_rate_plain = 0.0;  // (measure again)
    ctx->upload_calls = 0;
    return LM_OK;
  }
  if (option == LM_OPT_UPLOAD_THREADS && value >= 0) {
    ctx->opt_upload_threads = value;
    return LM_OK;
  }
  if (option == LM_OPT_UPLOAD_MIN_BYTES && value >= 0) {
    ctx->opt_upload_min_bytes = value;
    return LM_OK;
  }
  if (option == LM_OPT_BOOT_ROWS && value >= 1) {
    ctx->opt_boot_rows = (size_t)value;
    return LM_OK;
  }
  return LM_ERR_INVALID_ARG;
}

int lm_ctx_last_upload(const lm_ctx* ctx, uint64_t out[4]) {
  if (!ctx || !out) return LM_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) out[i] = ctx->upload_stats[i];
  return LM_OK;
}

int lm_ctx_last_stats(const lm_ctx* ctx, uint64_t out[4]) {
  if (!ctx || !out) return LM_ERR_INVALID_ARG;
  for (int i = 0; i < 4; ++i) out[i] = ctx->stats[i];
  return LM_OK;
}

float lm_ctx_last_sm_clock_mhz(const lm_ctx* ctx) { return ctx ? ctx->sm_mhz : 0.f; }

float lm_ctx_last_kernel_ms(const lm_ctx* ctx) {
  if (!ctx) return 0.f;
  if (!ctx->timed) return ctx->last_ms;  // set by the last *_finish
  lm_ctx* c = const_cast<lm_ctx*>(ctx);
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return 0.f;
  c->last_ms = ms;
  return ms;
}

// ----------------------------------------------- host-buffer trait mirrors

static int encode_impl(lm_ctx* ctx, int alphabet, const uint8_t* ascii, size_t n, uint8_t* dst,
                       size_t* bad_index, bool lossy) {
  LM_CHECK(bind(ctx));
  if (!valid_alphabet(alphabet) || (n && (!ascii || !dst))) return LM_ERR_INVALID_ARG;
  if (n == 0) return LM_OK;
  LM_CHECK(ensure(ctx, ctx->d_in, n));
  LM_CHECK(ensure(ctx, ctx->d_out, n));
  unsigned long long* d_bad = (unsigned long long*)ctx->d_small.p;
  LM_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, ctx->stream));
  LM_TRY(cudaMemcpyAsync(ctx->d_in.p, ascii, n, cudaMemcpyHostToDevice, ctx->stream));
  timer_begin(ctx);
  LM_TRY(launch_encode((const uint8_t*)ctx->d_in.p, n, (uint8_t*)ctx->d_out.p, alphabet, d_bad,
                       ctx->sm_count, ctx->stream));
  timer_end(ctx);
  ctx->launches += 1;
  unsigned long long bad = 0;
  LM_TRY(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaMemcpyAsync(dst, ctx->d_out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  if (!lossy && bad != ~0ull) {
    if (bad_index) *bad_index = (size_t)bad;
    return LM_ERR_INVALID_SYMBOL;
  }
  return LM_OK;
}

int lm_encode(lm_ctx* ctx, int alphabet, const uint8_t* ascii, size_t n, uint8_t* dst,
              size_t* bad_index) {
  return encode_impl(ctx, alphabet, ascii, n, dst, bad_index, false);
}

int lm_encode_lossy(lm_ctx* ctx, int alphabet, const uint8_t* ascii, size_t n, uint8_t* dst) {
  return encode_impl(ctx, alphabet, ascii, n, dst, nullptr, true);
}

int lm_stripe(lm_ctx* ctx, int alphabet, const uint8_t* sym, size_t n, uint8_t* dst,
              size_t dst_stride) {
  LM_CHECK(bind(ctx));
  if (!valid_alphabet(alphabet) || (n && (!sym || !dst)) || dst_stride < 32) return LM_ERR_INVALID_ARG;
  if (n == 0) return LM_OK;
  const size_t rows = (n + 31) / 32;
  LM_CHECK(ensure(ctx, ctx->d_in, n + 16));  // the stripe kernel reads whole 4-byte words
  LM_CHECK(ensure(ctx, ctx->d_out, rows * 32));
  LM_TRY(cudaMemcpyAsync(ctx->d_in.p, sym, n, cudaMemcpyHostToDevice, ctx->stream));
  LM_TRY(cudaMemsetAsync((uint8_t*)ctx->d_in.p + n, 0, 16, ctx->stream));  // bytes of the last word past n
  timer_begin(ctx);
  LM_TRY(launch_stripe((const uint8_t*)ctx->d_in.p, false, n, rows, 0, rows, (uint8_t*)ctx->d_out.p,
                       alphabet, nullptr, ctx->sm_count, ctx->stream));
  timer_end(ctx);
  ctx->launches += 1;
  LM_TRY(copy_rows(dst, dst_stride, ctx->d_out.p, 32, 32, rows, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  return LM_OK;
}

int lm_configure_wrap(lm_ctx* ctx, int alphabet, uint8_t* data, size_t rows, size_t stride,
                      size_t wrap) {
  LM_CHECK(bind(ctx));
  if (!valid_alphabet(alphabet) || !data || stride < 32) return LM_ERR_INVALID_ARG;
  if (wrap == 0) return LM_OK;
  // the wrap rows only depend on the first min(rows, wrap) base rows; the
  // device works on a compact copy [base | wrap] of those
  const size_t base = std::min(rows, wrap);
  LM_CHECK(ensure(ctx, ctx->d_in, (rows == 0 ? 1 : rows + wrap) * 32));
  uint8_t* d = (uint8_t*)ctx->d_in.p;
  if (rows == 0) {
    // DenseMatrix::resize default-fills new rows (dense.rs) and the loop at
    // seq.rs:373-378 then copies defaults: every wrap row is the default symbol
    std::vector<uint8_t> fill(32, pad_symbol(alphabet));
    for (size_t i = 0; i < wrap; ++i) std::memcpy(data + i * stride, fill.data(), 32);
    return LM_OK;
  }
  LM_TRY(copy_rows(d, 32, data, stride, 32, base, cudaMemcpyHostToDevice, ctx->stream));
  timer_begin(ctx);
  LM_TRY(launch_wrap(d, rows, rows, rows + wrap, alphabet, ctx->stream));
  timer_end(ctx);
  ctx->launches += 1;
  LM_TRY(copy_rows(data + rows * stride, stride, d + rows * 32, 32, 32, wrap,
                           cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  return LM_OK;
}

static int score_rows_host(lm_ctx* ctx, int alphabet, bool is_u8, const void* pssm, size_t m,
                           size_t pssm_stride, const uint8_t* seq, size_t rows, size_t wrap,
                           size_t seq_stride, size_t length, size_t row_begin, size_t row_end,
                           void* out, size_t out_stride, size_t* out_rows) {
  LM_CHECK(bind(ctx));
  if (out_rows) *out_rows = 0;
  if (!valid_alphabet(alphabet) || !pssm || !seq || seq_stride < 32 || out_stride < 32 ||
      pssm_stride < (size_t)alphabet_k(alphabet))
    return LM_ERR_INVALID_ARG;
  int st = check_score_args(m, rows, wrap, length, row_begin, row_end);
  if (st == LM_EMPTY) return LM_OK;  // scores.resize(0, 0)
  if (st < 0) return st;
  if (!out) return LM_ERR_INVALID_ARG;
  LM_CHECK(limits_ok(alphabet, m));
  const size_t n = row_end - row_begin;
  const size_t slab = std::min(rows + wrap, row_end + m - 1) - row_begin;  // rows uploaded
  const size_t esz = is_u8 ? 1 : 4;
  LM_CHECK(ensure(ctx, ctx->d_in, slab * 32));
  LM_CHECK(ensure(ctx, ctx->d_out, n * 32 * esz));
  LM_CHECK(upload_table(ctx, alphabet, pssm, m, pssm_stride, is_u8));
  LM_TRY(copy_rows(ctx->d_in.p, 32, seq + row_begin * seq_stride, seq_stride, 32, slab,
                           cudaMemcpyHostToDevice, ctx->stream));
  LM_CHECK(run_score(ctx, alphabet, is_u8, MODE_STORE, (const uint8_t*)ctx->d_in.p, slab, m, 0, n,
                     ctx->d_out.p, 32, 0.f, 0, nullptr, 0));
  LM_TRY(copy_rows(out, out_stride * esz, ctx->d_out.p, 32 * esz, 32 * esz, n,
                           cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  if (out_rows) *out_rows = n;
  return LM_OK;
}

int lm_score_f32_rows(lm_ctx* ctx, int alphabet, const float* pssm, size_t m, size_t pssm_stride,
                      const uint8_t* seq, size_t rows, size_t wrap, size_t seq_stride, size_t length,
                      size_t row_begin, size_t row_end, float* out, size_t out_stride,
                      size_t* out_rows) {
  return score_rows_host(ctx, alphabet, false, pssm, m, pssm_stride, seq, rows, wrap, seq_stride,
                         length, row_begin, row_end, out, out_stride, out_rows);
}

int lm_score_u8_rows(lm_ctx* ctx, int alphabet, const uint8_t* dm, size_t m, size_t dm_stride,
                     const uint8_t* seq, size_t rows, size_t wrap, size_t seq_stride, size_t length,
                     size_t row_begin, size_t row_end, uint8_t* out, size_t out_stride,
                     size_t* out_rows) {
  return score_rows_host(ctx, alphabet, true, dm, m, dm_stride, seq, rows, wrap, seq_stride, length,
                         row_begin, row_end, out, out_stride, out_rows);
}

static int reduce_host(lm_ctx* ctx, const void* scores, bool is_u8, size_t rows, size_t stride,
                       ColKeys* ck) {
  LM_CHECK(bind(ctx));
  if (!scores || stride < 32) return LM_ERR_INVALID_ARG;
  if (rows >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  const size_t esz = is_u8 ? 1 : 4;
  LM_CHECK(ensure(ctx, ctx->d_in, rows * 32 * esz));
  LM_TRY(copy_rows(ctx->d_in.p, 32 * esz, scores, stride * esz, 32 * esz, rows,
                           cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* small = (uint8_t*)ctx->d_small.p;
  LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
  timer_begin(ctx);
  LM_TRY(launch_colmax(ctx->d_in.p, is_u8, rows, 32, (unsigned long long*)small,
                       (uint32_t*)(small + 256), ctx->sm_count, ctx->stream));
  timer_end(ctx);
  ctx->launches += 1;
  return fetch_colkeys(ctx, ck);
}

static int argmax_host(lm_ctx* ctx, const void* scores, bool is_u8, size_t rows, size_t stride,
                       int rule, lm_hit* best) {
  if (!best) return LM_ERR_INVALID_ARG;
  if (rows == 0) return LM_EMPTY;  // pli/mod.rs:136-138
  ColKeys ck;
  LM_CHECK(reduce_host(ctx, scores, is_u8, rows, stride, &ck));
  lm_colwin cw[32];
  decode_colwins(ck.keys, is_u8, cw);
  return apply_rule(cw, bits_to_score(ck.first_cell, is_u8), rule, is_u8, rows, 0, best);
}

int lm_argmax_f32(lm_ctx* ctx, const float* scores, size_t rows, size_t stride, int rule, lm_hit* best) {
  return argmax_host(ctx, scores, false, rows, stride, rule, best);
}
int lm_argmax_u8(lm_ctx* ctx, const uint8_t* scores, size_t rows, size_t stride, int rule, lm_hit* best) {
  return argmax_host(ctx, scores, true, rows, stride, rule, best);
}

int lm_max_f32(lm_ctx* ctx, const float* scores, size_t rows, size_t stride, float* best) {
  lm_hit h;
  int s = argmax_host(ctx, scores, false, rows, stride, LM_RULE_GENERIC, &h);
  if (s == LM_OK && best) *best = h.score;
  return s;
}
int lm_max_f32_rule(lm_ctx* ctx, const float* scores, size_t rows, size_t stride, int rule, float* best) {
  float m = 0.f;
  int s = lm_max_f32(ctx, scores, rows, stride, &m);
  // avx2.rs:438-441: the AVX2 accumulators start at 0.0, so that back-end never returns less
  if (s == LM_OK && rule == LM_RULE_AVX2 && !(m > 0.0f)) m = 0.0f;
  if (s == LM_OK && best) *best = m;
  return s;
}
int lm_max_u8(lm_ctx* ctx, const uint8_t* scores, size_t rows, size_t stride, uint8_t* best) {
  lm_hit h;
  int s = argmax_host(ctx, scores, true, rows, stride, LM_RULE_GENERIC, &h);
  if (s == LM_OK && best) *best = (uint8_t)h.score;
  return s;
}

static int fetch_hits(lm_ctx* ctx, lm_hit* hits, size_t cap, size_t* count, size_t small_off = 0) {
  unsigned long long total = 0;
  LM_TRY(cudaMemcpyAsync(&total, (uint8_t*)ctx->d_small.p + small_off + 264, 8, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  size_t n = (size_t)std::min<unsigned long long>(total, cap);
  if (n) {
    LM_TRY(cudaMemcpyAsync(hits, ctx->d_hits.p, n * sizeof(lm_hit), cudaMemcpyDeviceToHost, ctx->stream));
    LM_TRY(cudaStreamSynchronize(ctx->stream));
    sort_hits(hits, n);  // the reference's row-major order (pli/mod.rs:212-219)
  }
  if (count) *count = (size_t)total;
  return total > cap ? LM_ERR_OVERFLOW : LM_OK;
}

static int threshold_host(lm_ctx* ctx, const void* scores, bool is_u8, size_t rows, size_t stride,
                          uint32_t thr_bits, lm_hit* hits, size_t cap, size_t* count) {
  LM_CHECK(bind(ctx));
  if (count) *count = 0;
  if (rows == 0) return LM_OK;
  if (!scores || stride < 32 || (cap && !hits)) return LM_ERR_INVALID_ARG;
  const size_t esz = is_u8 ? 1 : 4;
  LM_CHECK(ensure(ctx, ctx->d_in, rows * 32 * esz));
  LM_CHECK(ensure(ctx, ctx->d_hits, std::max<size_t>(cap, 1) * sizeof(lm_hit)));
  LM_TRY(copy_rows(ctx->d_in.p, 32 * esz, scores, stride * esz, 32 * esz, rows,
                           cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* small = (uint8_t*)ctx->d_small.p;
  LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
  timer_begin(ctx);
  LM_TRY(launch_threshold(ctx->d_in.p, is_u8, rows, 32, thr_bits, (lm_hit*)ctx->d_hits.p, cap,
                          (unsigned long long*)(small + 264), ctx->sm_count, ctx->stream));
  timer_end(ctx);
  ctx->launches += 1;
  return fetch_hits(ctx, hits, cap, count);
}

int lm_threshold_f32(lm_ctx* ctx, const float* scores, size_t rows, size_t stride, float t,
                     lm_hit* hits, size_t cap, size_t* count) {
  uint32_t bits;
  std::memcpy(&bits, &t, 4);
  return threshold_host(ctx, scores, false, rows, stride, bits, hits, cap, count);
}
int lm_threshold_u8(lm_ctx* ctx, const uint8_t* scores, size_t rows, size_t stride, uint8_t t,
                    lm_hit* hits, size_t cap, size_t* count) {
  return threshold_host(ctx, scores, true, rows, stride, t, hits, cap, count);
}

// ------------------------------------------------ device-resident sequences

static int seq_from_flat(lm_ctx* ctx, int alphabet, const uint8_t* src, size_t n, bool from_ascii,
                         bool lossy, size_t wrap, lm_seq** out, size_t* bad_index) {
  LM_CHECK(bind(ctx));
  if (!out || !valid_alphabet(alphabet) || (n && !src)) return LM_ERR_INVALID_ARG;
  *out = nullptr;
  const size_t rows = (n + 31) / 32;
  if (rows + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  lm_seq* s = nullptr;
  LM_CHECK(seq_alloc(ctx, alphabet, rows, wrap, n, &s));
  int st = LM_OK;
  do {
    if ((st = ensure(ctx, ctx->d_in, n + 16)) < 0) break;  // the stripe kernel reads whole 4-byte words
    unsigned long long* d_bad = (unsigned long long*)ctx->d_small.p;
    cudaError_t e = cudaMemsetAsync(d_bad, 0xFF, 8, ctx->stream);
    if (e == cudaSuccess && n) {
      if (!from_ascii && upload_wants_packed(ctx, alphabet, 32, n, src)) {  // encoded DNA symbols: partly 2-bit packed (upload.cu)
        if ((st = upload_striped_dna(ctx, (uint8_t*)ctx->d_in.p, src, n)) < 0) break;
      } else {
        e = cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->stream);
      }
    }
    if (e == cudaSuccess) e = cudaMemsetAsync((uint8_t*)ctx->d_in.p + n, 0, 16, ctx->stream);  // last word past n
    if (e != cudaSuccess) { st = fail(ctx, e, "upload"); break; }
    timer_begin(ctx);
    // every allocated row gets its closed-form value, so growing `wrap` later is free
    e = launch_stripe((const uint8_t*)ctx->d_in.p, from_ascii, n, rows, 0, s->alloc_rows, s->data, alphabet,
                      d_bad, ctx->sm_count, ctx->stream);
    timer_end(ctx);
    ctx->launches += 1;
    if (e != cudaSuccess) { st = fail(ctx, e, "stripe kernel"); break; }
    unsigned long long bad = ~0ull;
    e = cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { st = fail(ctx, e, "stripe sync"); break; }
    if (from_ascii && !lossy && bad != ~0ull) {
      if (bad_index) *bad_index = (size_t)bad;
      st = LM_ERR_INVALID_SYMBOL;
    }
  } while (0);
  if (st < 0) { lm_seq_free(ctx, s); return st; }
  *out = s;
  return LM_OK;
}

int lm_seq_from_ascii(lm_ctx* ctx, int alphabet, const uint8_t* ascii, size_t n, int lossy, size_t wrap,
                      lm_seq** out, size_t* bad_index) {
  return seq_from_flat(ctx, alphabet, ascii, n, true, lossy != 0, wrap, out, bad_index);
}

int lm_seq_from_symbols(lm_ctx* ctx, int alphabet, const uint8_t* sym, size_t n, size_t wrap, lm_seq** out) {
  return seq_from_flat(ctx, alphabet, sym, n, false, true, wrap, out, nullptr);
}

int lm_seq_from_striped(lm_ctx* ctx, int alphabet, const uint8_t* data, size_t rows, size_t wrap,
                        size_t stride, size_t length, lm_seq** out) {
  LM_CHECK(bind(ctx));
  if (!out || !valid_alphabet(alphabet) || ((rows + wrap) && !data) || stride < 32)
    return LM_ERR_INVALID_ARG;  // `length` may exceed rows * 32: a row slab of a longer sequence
  *out = nullptr;
  if (rows + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  lm_seq* s = nullptr;
  LM_CHECK(seq_alloc(ctx, alphabet, rows, wrap, length, &s));
  cudaError_t e = cudaSuccess;
  const size_t nbytes = (rows + wrap) * 32;
  const bool large = alphabet == LM_DNA && stride == 32 && nbytes >= (size_t)ctx->opt_upload_min_bytes;
  const bool packed = upload_wants_packed(ctx, alphabet, stride, nbytes, data);
  const auto t0 = std::chrono::steady_clock::now();
  if (packed) {
    // large DNA matrix: plain copy from the front, 2-bit packed chunks from the back (upload.cu)
    int st = upload_striped_dna(ctx, s->data, data, nbytes);
    if (st < 0) { lm_seq_free(ctx, s); return st; }
  } else if (rows + wrap) {
    e = copy_rows(s->data, 32, data, stride, 32, rows + wrap, cudaMemcpyHostToDevice, ctx->stream);
    ctx->upload_stats[0] = (rows + wrap) * 32;
    ctx->upload_stats[1] = ctx->upload_stats[2] = ctx->upload_stats[3] = 0;
  }
  if (e == cudaSuccess)
    e = cudaMemsetAsync(s->data + (rows + wrap) * 32, pad_symbol(alphabet), SLACK_ROWS * 32, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { int st = fail(ctx, e, "upload striped"); lm_seq_free(ctx, s); return st; }
  if (large && ctx->opt_upload == 1)
    upload_note_rate(ctx, packed, nbytes, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  *out = s;
  return LM_OK;
}

int lm_seq_from_striped_slab(lm_ctx* ctx, int alphabet, const uint8_t* data, size_t rows, size_t halo, size_t stride,
                             size_t length, uint64_t row_offset, uint64_t rows_global, lm_seq** out) {
  if (row_offset + rows > rows_global || (uint64_t)(length + 31) / 32 != rows_global) return LM_ERR_INVALID_ARG;
  int st = lm_seq_from_striped(ctx, alphabet, data, rows, halo, stride, length, out);
  if (st < 0) return st;
  (*out)->row_offset = row_offset;
  (*out)->rows_global = rows_global;
  return LM_OK;
}

int lm_seq_from_device(lm_ctx* ctx, int alphabet, void* device_ptr, size_t rows, size_t wrap,
                       size_t length, lm_seq** out) {
  LM_CHECK(bind(ctx));
  if (!out || !valid_alphabet(alphabet) || !device_ptr || ((uintptr_t)device_ptr & 15))
    return LM_ERR_INVALID_ARG;
  if (rows + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  lm_seq* s = new (std::nothrow) lm_seq();
  if (!s) return LM_ERR_OOM;
  s->alphabet = alphabet;
  s->device = ctx->device;
  s->data = (uint8_t*)device_ptr;
  s->owned = false;
  s->rows = rows;
  s->wrap = wrap;
  s->length = length;
  s->alloc_rows = rows + wrap;
  *out = s;
  return LM_OK;
}

int lm_seq_synthetic(lm_ctx* ctx, int alphabet, uint64_t seed, uint64_t total_length, uint64_t row_begin,
                     uint64_t n_rows, size_t wrap, lm_seq** out) {
  LM_CHECK(bind(ctx));
  if (!out || !valid_alphabet(alphabet)) return LM_ERR_INVALID_ARG;
  *out = nullptr;
  const uint64_t rows_global = (total_length + 31) / 32;
  if (row_begin + n_rows > rows_global) return LM_ERR_INVALID_ARG;
  if (rows_global + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  lm_seq* s = nullptr;
  LM_CHECK(seq_alloc(ctx, alphabet, n_rows, wrap, total_length, &s));
  s->row_offset = row_begin;
  s->rows_global = rows_global;
  timer_begin(ctx);
  cudaError_t e = launch_synth(s->data, s->alloc_rows, rows_global, row_begin, total_length, seed, alphabet,
                               ctx->sm_count, ctx->stream);
  timer_end(ctx);
  ctx->launches += 1;
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { int st = fail(ctx, e, "synth kernel"); lm_seq_free(ctx, s); return st; }
  *out = s;
  return LM_OK;
}

int lm_seq_configure_wrap(lm_ctx* ctx, lm_seq* seq, size_t wrap) {
  LM_CHECK(bind(ctx));
  if (!seq) return LM_ERR_INVALID_ARG;
  if (wrap <= seq->wrap) return LM_OK;  // seq.rs:370: only grows
  if (!seq->owned) return LM_ERR_UNSUPPORTED;
  // a row slab's halo rows are the NEXT rows of the global matrix, which this handle does not hold:
  // the local closed form would overwrite them with wrong symbols
  if (seq->is_slab() || seq->length > seq->rows * 32) return LM_ERR_UNSUPPORTED;
  if (seq->rows + wrap >= (1ull << 32)) return LM_ERR_UNSUPPORTED;
  if (seq->rows + wrap + SLACK_ROWS > seq->alloc_rows) {
    size_t new_alloc = seq->rows + wrap + SLACK_ROWS;
    uint8_t* nd = nullptr;
    LM_TRY(cudaMallocAsync((void**)&nd, new_alloc * 32, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(nd, seq->data, (seq->rows + seq->wrap) * 32, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(nd + (seq->rows + seq->wrap) * 32, pad_symbol(seq->alphabet),
                                              (new_alloc - seq->rows - seq->wrap) * 32, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(nd, ctx->stream); return fail(ctx, e, "grow sequence"); }
    cudaFreeAsync(seq->data, ctx->stream);
    seq->data = nd;
    seq->alloc_rows = new_alloc;
  }
  if (seq->rows == 0) {
    LM_TRY(cudaMemsetAsync(seq->data, pad_symbol(seq->alphabet), wrap * 32, ctx->stream));
  } else {
    timer_begin(ctx);
    LM_TRY(launch_wrap(seq->data, seq->rows, seq->rows + seq->wrap, seq->rows + wrap, seq->alphabet, ctx->stream));
    timer_end(ctx);
    ctx->launches += 1;
  }
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  seq->wrap = wrap;
  if (ctx->class_seq_id == seq->id) ctx->class_seq_id = 0;  // wrap rows changed
  return LM_OK;
}

int lm_seq_info(const lm_seq* seq, int* alphabet, size_t* rows, size_t* wrap, size_t* length) {
  if (!seq) return LM_ERR_INVALID_ARG;
  if (alphabet) *alphabet = seq->alphabet;
  if (rows) *rows = seq->rows;
  if (wrap) *wrap = seq->wrap;
  if (length) *length = seq->length;
  return LM_OK;
}

void* lm_seq_device_ptr(const lm_seq* seq) { return seq ? seq->data : nullptr; }

int lm_seq_download(lm_ctx* ctx, const lm_seq* seq, uint8_t* dst, size_t dst_stride) {
  LM_CHECK(bind(ctx));
  if (!seq || dst_stride < 32) return LM_ERR_INVALID_ARG;
  size_t n = seq->rows + seq->wrap;
  if (n == 0) return LM_OK;
  if (!dst) return LM_ERR_INVALID_ARG;
  LM_TRY(copy_rows(dst, dst_stride, seq->data, 32, 32, n, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  return LM_OK;
}

void lm_seq_free(lm_ctx* ctx, lm_seq* seq) {
  if (!seq) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->class_seq_id == seq->id) ctx->class_seq_id = 0;  // the class-coded copy dies with its sequence
  }
  if (seq->owned && seq->data) {
    if (ctx) cudaFreeAsync(seq->data, ctx->stream);
    else cudaFree(seq->data);
  }
  delete seq;
}

static int seq_score(lm_ctx* ctx, const lm_seq* seq, bool is_u8, const void* pssm, size_t m, size_t stride,
                     size_t row_begin, size_t row_end, void* out, size_t out_stride, int out_is_device,
                     size_t* out_rows) {
  LM_CHECK(bind(ctx));
  if (out_rows) *out_rows = 0;
  if (!seq || !pssm || out_stride < 32 || stride < (size_t)alphabet_k(seq->alphabet)) return LM_ERR_INVALID_ARG;
  if (seq->device != ctx->device) return LM_ERR_INVALID_ARG;
  int st = check_score_args(m, seq->rows, seq->wrap, seq->length, row_begin, row_end);
  if (st == LM_EMPTY) return LM_OK;
  if (st < 0) return st;
  if (!out) return LM_ERR_INVALID_ARG;
  LM_CHECK(limits_ok(seq->alphabet, m));
  const size_t n = row_end - row_begin;
  const size_t esz = is_u8 ? 1 : 4;
  LM_CHECK(upload_table(ctx, seq->alphabet, pssm, m, stride, is_u8));
  void* d_out = out;
  size_t d_stride = out_stride;
  if (!out_is_device) {
    LM_CHECK(ensure(ctx, ctx->d_out, n * 32 * esz));
    d_out = ctx->d_out.p;
    d_stride = 32;
  }
  LM_CHECK(run_score(ctx, seq->alphabet, is_u8, MODE_STORE, seq->data, seq->alloc_rows, m, row_begin, row_end,
                     d_out, d_stride, 0.f, 0, nullptr, 0));
  if (!out_is_device)
    LM_TRY(copy_rows(out, out_stride * esz, d_out, 32 * esz, 32 * esz, n, cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  if (out_rows) *out_rows = n;
  return LM_OK;
}

int lm_seq_score_f32(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride,
                     size_t row_begin, size_t row_end, float* out, size_t out_stride, int out_is_device,
                     size_t* out_rows) {
  return seq_score(ctx, seq, false, pssm, m, pssm_stride, row_begin, row_end, out, out_stride, out_is_device, out_rows);
}

int lm_seq_score_u8(lm_ctx* ctx, const lm_seq* seq, const uint8_t* dm, size_t m, size_t dm_stride,
                    size_t row_begin, size_t row_end, uint8_t* out, size_t out_stride, int out_is_device,
                    size_t* out_rows) {
  return seq_score(ctx, seq, true, dm, m, dm_stride, row_begin, row_end, out, out_stride, out_is_device, out_rows);
}

}  // extern "C"

namespace lmi {

int argmax_launch(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride, size_t row_begin,
                  size_t row_end, int slot, const PeerLink* link) {
  LM_CHECK(bind(ctx));
  if (!seq || !pssm || slot < 0 || slot >= LM_SLOTS || pssm_stride < (size_t)alphabet_k(seq->alphabet))
    return LM_ERR_INVALID_ARG;
  if (seq->device != ctx->device) return LM_ERR_INVALID_ARG;
  lm_ctx::Slot& sl = ctx->slots[slot];
  if (sl.state != 0) return LM_ERR_INVALID_ARG;  // previous launch on this slot not finished
  int st = check_score_args(m, seq->rows, seq->wrap, seq->length, row_begin, row_end);
  if (st < 0) return st;
  const size_t off = (size_t)slot * 1024;
  uint8_t* small = (uint8_t*)ctx->d_small.p + off;
  const bool linked = link && link->world > 1;
  if (st == LM_EMPTY) {
    sl.state = 2;
    if (linked) {  // the other ranks still wait for this rank's (empty) block
      LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
      LM_TRY(launch_publish(*link, (const unsigned long long*)small, ctx->stream));
      ctx->launches += 1;
    }
    return LM_OK;
  }
  LM_CHECK(limits_ok(seq->alphabet, m));
  LM_CHECK(upload_table(ctx, seq->alphabet, pssm, m, pssm_stride, false));
  const size_t n_rows = row_end - row_begin;
  PrunePlan plan;
  ProteinPlan pplan;
  const bool protein = seq->alphabet == LM_PROTEIN;
  if (ctx->opt_prune && n_rows > 2 * ctx->opt_boot_rows) {
    // a row-sharded scan is planned for the whole matrix: every rank filters at the global maximum
    const double npos = (double)(linked && seq->rows_global ? seq->rows_global : n_rows) * 32.0;
    if (protein) {
      pplan = cached_protein_plan(ctx, pssm, m, pssm_stride, true, 0.0, npos);
      plan = pplan.plan;
    } else {
      plan = cached_plan(ctx, pssm, m, pssm_stride, true, 0.0, npos, seq->id);
    }
    if (ctx->opt_prune == 2 && plan.r > 0) plan.ok = true;
  }
  if (plan.ok) LM_CHECK(protein ? upload_table32(ctx, pssm, m, pssm_stride) : upload_table8(ctx, pssm, m, pssm_stride));
  LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
  LM_TRY(cudaEventRecord(sl.ev0, ctx->stream));
  if (plan.ok) {
    // 1. exact scan of the first rows seeds the per-column winners (= the running maximum)
    // 2. filter + exact recheck over the rest, raising the maximum as it goes
    LM_CHECK(run_score(ctx, seq->alphabet, false, MODE_ARGMAX, seq->data, seq->alloc_rows, m, row_begin,
                       row_begin + ctx->opt_boot_rows, nullptr, 0, 0.f, 0, nullptr, 0, false, off));
    PruneArgs pa;
    pa.seq = seq->data;
    if (protein) {  // the filter reads class codes, the recheck the residues
      LM_CHECK(ensure_class_copy(ctx, seq, pplan.class_of));
      pa.seq = (const uint8_t*)ctx->d_class.p;
      pa.seq_exact = seq->data;
      pa.tab_stride = 32;
      pa.pad_fill = pplan.class_of[20];
    }
    pa.avail_rows = seq->alloc_rows;
    pa.row_begin = row_begin;
    pa.row_first = row_begin + ctx->opt_boot_rows;
    pa.row_end = row_end;
    pa.pad_sym = pad_symbol(seq->alphabet);
    pa.table = (const float*)ctx->d_table2.p;
    pa.plan = plan;
    pa.colkeys = (unsigned long long*)small;
    pa.stats = (unsigned long long*)(small + 272);
    pa.dense = !protein && want_dense(ctx, plan);
    if (linked) pa.link = *link;  // the scan kernel itself shares the maximum and publishes the block
    LM_TRY(launch_prune(false, pa, std::max(1, ctx->sm_count - ctx->opt_reserve_sms), ctx->stream));
    ctx->launches += 1;
  } else {
    LM_CHECK(run_score(ctx, seq->alphabet, false, MODE_ARGMAX, seq->data, seq->alloc_rows, m, row_begin, row_end,
                       nullptr, 0, 0.f, 0, nullptr, 0, false, off));
    if (linked) {
      LM_TRY(launch_publish(*link, (const unsigned long long*)small, ctx->stream));
      ctx->launches += 1;
    }
  }
  LM_TRY(cudaEventRecord(sl.ev1, ctx->stream));
  LM_TRY(cudaMemcpyAsync(sl.h, small, sizeof(SmallBlock), cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaEventRecord(sl.done, ctx->stream));
  sl.state = 1;
  sl.n_rows = n_rows;
  sl.seq_id = seq->id;
  sl.pruned = plan.ok;
  sl.r = plan.r | (plan.pair << 8);
  return LM_OK;
}

}  // namespace lmi

extern "C" {

int lm_seq_score_argmax_f32_launch(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m,
                                   size_t pssm_stride, size_t row_begin, size_t row_end, int slot) {
  return argmax_launch(ctx, seq, pssm, m, pssm_stride, row_begin, row_end, slot, nullptr);
}

int lm_seq_score_argmax_f32_finish(lm_ctx* ctx, int slot, int rule, lm_hit* best, lm_colwin* per_column,
                                   float* first_cell) {
  LM_CHECK(bind(ctx));
  if (!best || slot < 0 || slot >= LM_SLOTS) return LM_ERR_INVALID_ARG;
  lm_ctx::Slot& sl = ctx->slots[slot];
  if (per_column) std::memset(per_column, 0, 32 * sizeof(lm_colwin));
  if (sl.state == 0) return LM_ERR_INVALID_ARG;
  if (sl.state == 2) { sl.state = 0; return LM_EMPTY; }
  sl.state = 0;
  LM_TRY(cudaEventSynchronize(sl.done));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess) { ctx->last_ms = ms; ctx->timed = false; }
  const SmallBlock* sb = (const SmallBlock*)sl.h;
  if (sl.pruned) note_density(ctx, sb->stats[0] + 128 * sb->stats[1], sl.n_rows);  // (an overflowing slice had > 128 entries)
  const bool flooded = sl.pruned && ctx->density > 8.0 && ctx->replan_state == 0;
  ctx->stats[0] = sl.pruned ? sb->stats[0] : 0;
  ctx->stats[1] = sl.pruned ? sb->stats[1] : 0;
  ctx->stats[2] = sl.pruned ? 1 : 0;
  ctx->stats[3] = sl.pruned ? (uint64_t)sl.r : 0;
  ctx->sm_mhz = (sl.pruned && sb->stats[4]) ? (float)((double)sb->stats[3] * 1e3 / (double)sb->stats[4]) : 0.f;
  lm_colwin cw[32];
  decode_colwins(sb->ck.keys, false, cw);
  if (per_column) std::memcpy(per_column, cw, sizeof(cw));
  const float fc = bits_to_score(sb->ck.first_cell, false);
  if (first_cell) *first_cell = fc;
  const int rs = apply_rule(cw, fc, rule, false, sl.n_rows, 0, best);
  if (flooded && rs == LM_OK && std::isfinite(best->score)) {  // plan the next scan of this matrix for the real maximum
    ctx->replan_key = ctx->plan_key;
    ctx->replan_t = (double)best->score;
    ctx->replan_seq = sl.seq_id;
    ctx->replan_state = 1;
  }
  return rs;
}

void* lm_ctx_result_block(lm_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= LM_SLOTS || !ctx->d_small.p) return nullptr;
  return (uint8_t*)ctx->d_small.p + (size_t)slot * 1024;
}

void lm_decode_keys(const uint64_t* keys, lm_colwin* out) {
  if (keys && out) decode_colwins((const unsigned long long*)keys, false, out);
}

int lm_seq_score_argmax_f32(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride,
                            size_t row_begin, size_t row_end, int rule, lm_hit* best, lm_colwin* per_column,
                            float* first_cell) {
  if (!best) return LM_ERR_INVALID_ARG;
  if (ctx && ctx->slots[0].state != 0) return LM_ERR_INVALID_ARG;  // slot 0 is busy with an async call
  int st = lm_seq_score_argmax_f32_launch(ctx, seq, pssm, m, pssm_stride, row_begin, row_end, 0);
  if (st < 0) return st;
  return lm_seq_score_argmax_f32_finish(ctx, 0, rule, best, per_column, first_cell);
}

}  // extern "C"

namespace lmi {

int threshold_launch(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride, size_t row_begin,
                     size_t row_end, float t, lm_hit* d_hits, size_t cap, size_t small_off, const PeerLink* link,
                     PrunePlan* plan_out) {
  LM_CHECK(limits_ok(seq->alphabet, m));
  LM_CHECK(upload_table(ctx, seq->alphabet, pssm, m, pssm_stride, false));
  PrunePlan plan;
  ProteinPlan pplan;
  const bool protein = seq->alphabet == LM_PROTEIN;
  const bool linked = link && link->world > 1;
  if (ctx->opt_prune && row_end - row_begin > 4096) {
    const double npos = (double)(row_end - row_begin) * 32.0;
    if (protein) {
      pplan = cached_protein_plan(ctx, pssm, m, pssm_stride, false, (double)t, npos);
      plan = pplan.plan;
    } else {
      plan = cached_plan(ctx, pssm, m, pssm_stride, false, (double)t, npos, seq->id);
    }
    if (ctx->opt_prune == 2 && plan.r > 0) plan.ok = true;
  }
  if (plan_out) *plan_out = plan;
  // a slab reports its hits in the coordinates of the whole matrix
  const uint64_t row_add = seq->is_slab() ? seq->row_offset + row_begin : 0;
  const uint64_t pos_rows = seq->is_slab() ? seq->rows_global : 0;
  uint8_t* small = (uint8_t*)ctx->d_small.p + small_off;
  LM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
  ctx->stats[0] = ctx->stats[1] = ctx->stats[2] = ctx->stats[3] = 0;
  timer_begin(ctx);
  if (plan.ok) {
    LM_CHECK(protein ? upload_table32(ctx, pssm, m, pssm_stride) : upload_table8(ctx, pssm, m, pssm_stride));
    PruneArgs pa;
    pa.seq = seq->data;
    if (protein) {
      LM_CHECK(ensure_class_copy(ctx, seq, pplan.class_of));
      pa.seq = (const uint8_t*)ctx->d_class.p;
      pa.seq_exact = seq->data;
      pa.tab_stride = 32;
      pa.pad_fill = pplan.class_of[20];
    }
    pa.avail_rows = seq->alloc_rows;
    pa.row_begin = row_begin;
    pa.row_first = row_begin;
    pa.row_end = row_end;
    pa.pad_sym = pad_symbol(seq->alphabet);
    pa.table = (const float*)ctx->d_table2.p;
    pa.plan = plan;
    pa.thr = t;
    pa.hits = d_hits;
    pa.cap = cap;
    pa.colkeys = (unsigned long long*)small;
    pa.count = (unsigned long long*)(small + 264);
    pa.stats = (unsigned long long*)(small + 272);
    pa.row_add = row_add;
    pa.pos_rows = pos_rows;
    pa.dense = !protein && want_dense(ctx, plan);
    if (linked) pa.link = *link;
    LM_TRY(launch_prune(true, pa, std::max(1, ctx->sm_count - ctx->opt_reserve_sms), ctx->stream));
  } else {
    ScoreArgs a;
    a.seq = seq->data;
    a.avail_rows = seq->alloc_rows;
    a.row_begin = row_begin;
    a.row_end = row_end;
    a.pad_sym = pad_symbol(seq->alphabet);
    a.tables = (const uint32_t*)ctx->d_table.p;
    a.m_pad = table_mpad(m);
    a.n_motifs = 1;
    a.colkeys = (unsigned long long*)small;
    a.first_cell = (uint32_t*)(small + 256);
    a.count = (unsigned long long*)(small + 264);
    std::memcpy(&a.thr_bits, &t, 4);
    a.hits = d_hits;
    a.cap = cap;
    a.row_add = row_add;
    a.pos_rows = pos_rows;
    LM_TRY(launch_score(seq->alphabet, MODE_THRESHOLD, false, a, std::max(1, ctx->sm_count - ctx->opt_reserve_sms),
                        ctx->stream));
    if (linked) {
      LM_TRY(launch_publish(*link, (const unsigned long long*)small, ctx->stream));
      ctx->launches += 1;
    }
  }
  timer_end(ctx);
  ctx->launches += 1;
  return LM_OK;
}

}  // namespace lmi

extern "C" {

int lm_seq_score_threshold_f32(lm_ctx* ctx, const lm_seq* seq, const float* pssm, size_t m, size_t pssm_stride,
                               size_t row_begin, size_t row_end, float t, lm_hit* hits, size_t cap, size_t* count) {
  LM_CHECK(bind(ctx));
  if (count) *count = 0;
  if (!seq || !pssm || (cap && !hits) || pssm_stride < (size_t)alphabet_k(seq->alphabet)) return LM_ERR_INVALID_ARG;
  if (seq->device != ctx->device) return LM_ERR_INVALID_ARG;
  int st = check_score_args(m, seq->rows, seq->wrap, seq->length, row_begin, row_end);
  if (st == LM_EMPTY) return LM_OK;
  if (st < 0) return st;
  LM_CHECK(ensure(ctx, ctx->d_hits, std::max<size_t>(cap, 1) * sizeof(lm_hit)));
  PrunePlan plan;
  LM_CHECK(threshold_launch(ctx, seq, pssm, m, pssm_stride, row_begin, row_end, t, (lm_hit*)ctx->d_hits.p, cap,
                            SYNC_SMALL_OFF, nullptr, &plan));
  if (plan.ok) {
    LM_CHECK(fetch_stats(ctx, plan, SYNC_SMALL_OFF));
    note_density(ctx, ctx->stats[0] + 128 * ctx->stats[1], row_end - row_begin);
  }
  return fetch_hits(ctx, hits, cap, count, SYNC_SMALL_OFF);
}

}  // extern "C"

// ---- many motifs over one resident sequence (the lightmotif-cli job loop, main.rs:540-562).  Every
// motif is planned on the host (in parallel; the plans are cached per motif set); motifs whose
// filters share a kernel instance (look-ups per position, single / paired) are swept over the
// sequence by ONE launch (work item = (motif, tile), tables travel with the tiles), the running
// maxima of an argmax batch are seeded by ONE multi-motif exact launch over the first rows, motifs
// without a profitable filter run the exact kernel; the host synchronises ONCE at the end.
namespace {

struct BatchItem {
  size_t m = 0;
  bool pruned = false;
  PrunePlan plan;
  size_t tab16_off = 0, tab8_off = 0;  // byte offsets into the staged table buffer
  int group = -1;
};
struct BatchGroup {
  int r = 0, pair = 0, m_max = 0;
  uint32_t pair_mul = 2;
  std::vector<uint32_t> motifs;
  size_t desc_off = 0;  // byte offset of the group's BatchMotif array in the staged buffer
};

constexpr size_t BATCH_BLOCK = 512;  // bytes of device scratch per motif (SmallBlock layout)

// reverse complement of a DNA matrix: rows reversed, columns A<->T, C<->G (pwm/mod.rs:566-577;
// symbol order A, C, T, G, N: abc.rs:113-135)
void reverse_complement(const float* p, size_t m, size_t stride, float* out) {
  static const int comp[5] = {2, 3, 0, 1, 4};
  for (size_t j = 0; j < m; ++j)
    for (int c = 0; c < 5; ++c) out[j * stride + c] = p[(m - 1 - j) * stride + comp[c]];
}

}  // namespace

// host-side state of the batch calls (one per context): the last motif set, planned and staged
struct lm_batch_cache {
  std::vector<uint8_t> key;
  std::vector<float> mats;  // the matrices scanned (direct and / or reverse complement), DenseMatrix layout
  std::vector<uint64_t> offs;
  std::vector<uint32_t> widths;
  std::vector<BatchItem> items;
  std::vector<BatchGroup> groups;
  size_t table_bytes = 0, total_bytes = 0, toff_off = 0, mpad_off = 0, max_m = 0;
};

namespace {

// plan every motif (cached), lay out + upload tables, per-group descriptor arrays and the
// multi-motif kernel's offset arrays.  Descriptors carry pointers into d_out (result blocks) /
// d_hits, so they are rebuilt whenever those move or the thresholds change.
int batch_prepare(lm_ctx* ctx, const lm_seq* seq, const float* pssms, const uint32_t* widths, const uint64_t* offsets,
                  size_t n_in, size_t pssm_stride, int strands, bool argmax_mode, const float* thresholds, size_t cap,
                  lm_batch_cache** out) {
  if (!ctx->batch) ctx->batch = new (std::nothrow) lm_batch_cache();
  if (!ctx->batch) return LM_ERR_OOM;
  lm_batch_cache& B = *ctx->batch;
  const int alphabet = seq->alphabet;
  const int kp = table_kp(alphabet);
  const size_t k = (size_t)alphabet_k(alphabet);
  const size_t n_rows = seq->rows;
  const bool both = strands == 3, rev_only = strands == 2;
  const size_t n = n_in * (both ? 2 : 1);
  // ---- key: everything the plans depend on
  std::vector<uint8_t> key;
  {
    uint64_t hdr[8] = {(uint64_t)alphabet, n_in, (uint64_t)strands, argmax_mode ? 1u : 0u, (uint64_t)ctx->opt_prune,
                       (uint64_t)ctx->opt_filter_rows, (uint64_t)ctx->opt_filter_mode,
                       (uint64_t)(63 - __builtin_clzll((unsigned long long)std::max<size_t>(n_rows, 1)))};
    key.insert(key.end(), (const uint8_t*)hdr, (const uint8_t*)hdr + sizeof(hdr));
    key.insert(key.end(), (const uint8_t*)widths, (const uint8_t*)(widths + n_in));
    for (size_t q = 0; q < n_in; ++q)
      for (size_t j = 0; j < widths[q]; ++j) {
        const uint8_t* row = (const uint8_t*)(pssms + offsets[q] + j * pssm_stride);
        key.insert(key.end(), row, row + k * 4);
      }
    if (!argmax_mode) key.insert(key.end(), (const uint8_t*)thresholds, (const uint8_t*)(thresholds + n_in));
  }
  if (key != B.key) {
    B.key.clear();
    ctx->batch_staged = false;
    // ---- the matrices to scan: direct, reverse complement, or both (motif i -> entries 2i, 2i + 1)
    B.widths.assign(n, 0);
    B.offs.assign(n, 0);
    size_t total_rows = 0;
    for (size_t q = 0; q < n; ++q) {
      const size_t src = both ? q / 2 : q;
      if (widths[src] == 0) return LM_ERR_INVALID_ARG;
      B.widths[q] = widths[src];
      B.offs[q] = total_rows * pssm_stride;
      total_rows += widths[src];
    }
    B.mats.assign(total_rows * pssm_stride, 0.f);
    for (size_t q = 0; q < n; ++q) {
      const size_t src = both ? q / 2 : q;
      const bool rc = rev_only || (both && (q & 1));
      const float* p = pssms + offsets[src];
      float* dst = B.mats.data() + B.offs[q];
      if (rc) reverse_complement(p, widths[src], pssm_stride, dst);
      else
        for (size_t j = 0; j < widths[src]; ++j) std::memcpy(dst + j * pssm_stride, p + j * pssm_stride, k * 4);
    }
    // ---- plans (~0.3 ms of host work per motif: spread over the host cores)
    B.items.assign(n, BatchItem());
    const double npos = (double)n_rows * 32.0;
    for (size_t q = 0; q < n; ++q) B.items[q].m = B.widths[q];
    if (ctx->opt_prune && alphabet == LM_DNA && n_rows > 16384) {
      const int opt_prune = ctx->opt_prune, force_r = ctx->opt_filter_rows, force_mode = ctx->opt_filter_mode;
      const size_t length = seq->length;
      auto plan_range = [&](size_t q0, size_t q1) {
        for (size_t q = q0; q < q1; ++q) {
          BatchItem& it = B.items[q];
          if (length < it.m || it.m > 64) continue;
          const float* p = B.mats.data() + B.offs[q];
          const float thr = argmax_mode ? 0.f : thresholds[both ? q / 2 : q];
          it.plan = argmax_mode ? plan_prune(p, it.m, pssm_stride, estimate_max_score(p, it.m, pssm_stride, npos), true,
                                             npos, force_r, force_mode)
                                : plan_prune(p, it.m, pssm_stride, (double)thr, false, npos, force_r, force_mode);
          if (opt_prune == 2 && it.plan.r > 0) it.plan.ok = true;
          it.pruned = it.plan.ok;
        }
      };
      unsigned hw = std::thread::hardware_concurrency();
      const size_t n_threads = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)32, n / 8}));
      if (n_threads <= 1) {
        plan_range(0, n);
      } else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < n_threads; ++t) pool.emplace_back(plan_range, n * t / n_threads, n * (t + 1) / n_threads);
        for (auto& th : pool) th.join();
      }
    }
    // ---- groups of motifs that share a kernel instance; layout of the staged buffer
    B.groups.clear();
    size_t bytes = 0;
    B.max_m = 0;
    for (size_t q = 0; q < n; ++q) {
      BatchItem& it = B.items[q];
      B.max_m = std::max(B.max_m, it.m);
      it.tab16_off = bytes;
      bytes += (size_t)table_mpad(it.m) * kp * 4;
      it.tab8_off = bytes;
      if (it.pruned) {
        bytes += (size_t)table_mpad(it.m) * 8 * 4;
        int g = -1;
        for (size_t i = 0; i < B.groups.size(); ++i)
          if (B.groups[i].r == it.plan.r && B.groups[i].pair == it.plan.pair &&
              (!it.plan.pair || B.groups[i].pair_mul == it.plan.pair_mul))
            g = (int)i;
        if (g < 0) {
          B.groups.push_back(BatchGroup());
          g = (int)B.groups.size() - 1;
          B.groups[g].r = it.plan.r;
          B.groups[g].pair = it.plan.pair;
          B.groups[g].pair_mul = it.plan.pair_mul;
        }
        B.groups[g].motifs.push_back((uint32_t)q);
        B.groups[g].m_max = std::max(B.groups[g].m_max, (int)it.m);
        it.group = g;
      }
    }
    B.table_bytes = bytes = (bytes + 255) & ~(size_t)255;
    B.toff_off = bytes;
    bytes += n * 8;
    B.mpad_off = bytes;
    bytes = (bytes + n * 4 + 255) & ~(size_t)255;
    for (auto& g : B.groups) {
      g.desc_off = bytes;
      bytes += g.motifs.size() * sizeof(BatchMotif);
    }
    B.total_bytes = bytes;
    B.key.swap(key);
  }
  // ---- (re)build the staged image: tables once per motif set, descriptors every call
  LM_CHECK(ensure_pinned(ctx, B.total_bytes + 64));
  LM_CHECK(ensure(ctx, ctx->d_table, B.total_bytes + 64));
  LM_CHECK(ensure(ctx, ctx->d_out, n * BATCH_BLOCK + 1024));
  if (!argmax_mode) LM_CHECK(ensure(ctx, ctx->d_hits, std::max<size_t>(n * cap, 1) * sizeof(lm_hit)));
  ctx->table_key.clear();  // d_table now holds the batch, not a single matrix
  ctx->table8_key.clear();
  uint8_t* h = (uint8_t*)ctx->h_pin;
  uint8_t* dbase = (uint8_t*)ctx->d_table.p;
  const size_t first = ctx->batch_staged ? B.table_bytes : 0;  // the tables are still on the device
  if (!ctx->batch_staged) {
    for (size_t q = 0; q < n; ++q) {
      const BatchItem& it = B.items[q];
      const float* p = B.mats.data() + B.offs[q];
      build_table_f32(alphabet, p, it.m, pssm_stride, (uint32_t*)(h + it.tab16_off));
      if (it.pruned) {
        float* t8 = (float*)(h + it.tab8_off);
        std::memset(t8, 0, (size_t)table_mpad(it.m) * 8 * 4);
        for (size_t j = 0; j < it.m; ++j)
          for (int a = 0; a < 5; ++a) t8[j * 8 + a] = p[j * pssm_stride + a];
      }
    }
  }
  uint64_t* toff = (uint64_t*)(h + B.toff_off);
  int* mpads = (int*)(h + B.mpad_off);
  for (size_t q = 0; q < n; ++q) {
    toff[q] = B.items[q].tab16_off / 4;
    mpads[q] = table_mpad(B.items[q].m);
  }
  for (const auto& g : B.groups) {
    BatchMotif* bm = (BatchMotif*)(h + g.desc_off);
    for (size_t i = 0; i < g.motifs.size(); ++i) {
      const uint32_t q = g.motifs[i];
      const BatchItem& it = B.items[q];
      BatchMotif& M = bm[i];
      M = BatchMotif();
      for (int t = 0; t < 16; ++t) { M.tab[t] = it.plan.tab_lo[t]; M.tab[16 + t] = it.plan.tab_hi[t]; }
      M.c0 = it.plan.c0; M.delta = it.plan.delta; M.eps = it.plan.eps;
      M.m = it.plan.m; M.j0 = it.plan.j0; M.smax = it.plan.smax; M.pair_mul = it.plan.pair_mul;
      M.table = (const float*)(dbase + it.tab8_off);
      M.block = (unsigned long long*)((uint8_t*)ctx->d_out.p + (size_t)q * BATCH_BLOCK);
      if (!argmax_mode) {
        M.hits = (lm_hit*)ctx->d_hits.p + (size_t)q * cap;
        M.cap = cap;
        M.thr = thresholds[both ? q / 2 : q];
      }
    }
  }
  LM_TRY(cudaMemcpyAsync(dbase + first, h + first, B.total_bytes - first, cudaMemcpyHostToDevice, ctx->stream));
  // (the pinned image is rewritten by the next call: it must have left the host before we return --
  // every batch call ends with a stream synchronisation)
  ctx->batch_staged = true;
  *out = &B;
  return LM_OK;
}

// enqueue the scans of a prepared batch.  n_rows > 0, motifs longer than the sequence are skipped.
int batch_launch(lm_ctx* ctx, const lm_seq* seq, lm_batch_cache& B, bool argmax_mode, const float* thresholds, int strands,
                 size_t cap, uint64_t* n_pruned_out, uint64_t* launches_out) {
  const int alphabet = seq->alphabet;
  const size_t n = B.items.size(), n_rows = seq->rows;
  const int sms = std::max(1, ctx->sm_count - ctx->opt_reserve_sms);
  uint8_t* dbase = (uint8_t*)ctx->d_table.p;
  uint8_t* blocks = (uint8_t*)ctx->d_out.p;
  uint8_t* scratch = blocks + n * BATCH_BLOCK;  // launch-level counters of the grouped launches
  uint64_t launches = 0, n_pruned = 0;
  LM_TRY(cudaMemsetAsync(blocks, 0, n * BATCH_BLOCK + 1024, ctx->stream));
  // rows scored exactly first to seed the running maxima of the filtered argmax scans
  // (with motif-major work items the seed mattered: 8192 / 4096 / 2048 rows gave 30.1 / 32.0 / 34.2 ms per call of
  // config 2, because the whole grid started every motif at once; with interleaved motifs -- prune.cu BATCH_ILV --
  // 27.8 / 27.7 / 28.0 ms)
  const size_t boot = argmax_mode ? std::min<size_t>(n_rows, LM_BATCH_BOOT) : 0;
  bool any_pruned = false;
  for (const auto& it : B.items) any_pruned = any_pruned || (it.pruned && seq->length >= it.m);
  if (argmax_mode && any_pruned && boot) {
    // ONE multi-motif exact launch over rows [0, boot): tile resident in shared memory, loop over the motifs
    ScoreArgs a;
    a.seq = seq->data;
    a.avail_rows = seq->alloc_rows;
    a.row_begin = 0;
    a.row_end = boot;
    a.pad_sym = pad_symbol(alphabet);
    a.tables = (const uint32_t*)dbase;
    a.table_off = (const uint64_t*)(dbase + B.toff_off);
    a.m_pads = (const int*)(dbase + B.mpad_off);
    a.m_pad = table_mpad(B.max_m);
    a.n_motifs = (int)n;
    a.colkeys = (unsigned long long*)blocks;
    a.colkeys_stride = BATCH_BLOCK / 8;
    a.first_cell = (uint32_t*)(blocks + 256);
    a.first_cell_stride = BATCH_BLOCK / 4;
    LM_TRY(launch_score(alphabet, MODE_ARGMAX, false, a, sms, ctx->stream));
    ++launches;
  }
  // ONE launch per group of filtered motifs
  size_t gi = 0;
  for (const auto& g : B.groups) {
    PruneArgs pa;
    pa.seq = seq->data;
    pa.avail_rows = seq->alloc_rows;
    pa.row_begin = 0;
    pa.row_first = boot;
    pa.row_end = n_rows;
    pa.pad_sym = pad_symbol(alphabet);
    pa.plan.r = g.r;
    pa.plan.pair = g.pair;
    pa.plan.pair_mul = g.pair_mul;
    pa.batch = (const BatchMotif*)(dbase + g.desc_off);
    pa.n_batch = (uint32_t)g.motifs.size();
    pa.m_max = g.m_max;
    pa.stats = (unsigned long long*)(scratch + (gi % 16) * 64);
    if (n_rows > boot) {
      LM_TRY(launch_prune(!argmax_mode, pa, sms, ctx->stream));
      ++launches;
    }
    n_pruned += g.motifs.size();
    ++gi;
  }
  // motifs without a profitable filter: the exact kernel over every row
  for (size_t q = 0; q < n; ++q) {
    const BatchItem& it = B.items[q];
    if (it.pruned || seq->length < it.m) continue;
    uint8_t* block = blocks + q * BATCH_BLOCK;
    ScoreArgs a;
    a.seq = seq->data;
    a.avail_rows = seq->alloc_rows;
    a.row_begin = 0;
    a.row_end = n_rows;
    a.pad_sym = pad_symbol(alphabet);
    a.tables = (const uint32_t*)(dbase + it.tab16_off);
    a.m_pad = table_mpad(it.m);
    a.n_motifs = 1;
    a.colkeys = (unsigned long long*)block;
    a.first_cell = (uint32_t*)(block + 256);
    if (argmax_mode) {
      LM_TRY(launch_score(alphabet, MODE_ARGMAX, false, a, sms, ctx->stream));
      ++launches;
    } else {
      a.count = (unsigned long long*)(block + 264);
      const float thr = thresholds[strands == 3 ? q / 2 : q];
      std::memcpy(&a.thr_bits, &thr, 4);
      a.hits = (lm_hit*)ctx->d_hits.p + q * cap;
      a.cap = cap;
      LM_TRY(launch_score(alphabet, MODE_THRESHOLD, false, a, sms, ctx->stream));
      ++launches;
    }
  }
  *n_pruned_out = n_pruned;
  *launches_out = launches;
  return LM_OK;
}

}  // namespace

void free_batch_cache(lm_batch_cache* b) { delete b; }

extern "C" {

int lm_seq_score_argmax_f32_batch_rc(lm_ctx* ctx, const lm_seq* seq, const float* pssms, const uint32_t* widths,
                                     const uint64_t* offsets, size_t n_motifs, size_t pssm_stride, int strands, int rule,
                                     lm_hit* best, int* status) {
  LM_CHECK(bind(ctx));
  if (!seq || !pssms || !widths || !offsets || !best || pssm_stride < (size_t)alphabet_k(seq->alphabet) ||
      strands < 1 || strands > 3 || (strands != 1 && seq->alphabet != LM_DNA))
    return LM_ERR_INVALID_ARG;
  if (seq->device != ctx->device) return LM_ERR_INVALID_ARG;
  if (n_motifs == 0) return LM_OK;
  if (n_motifs > 65536) return LM_ERR_UNSUPPORTED;
  const size_t n = n_motifs * (strands == 3 ? 2 : 1);
  size_t max_m = 0;
  for (size_t q = 0; q < n_motifs; ++q) max_m = std::max<size_t>(max_m, widths[q]);
  int st = check_score_args(max_m ? max_m : 1, seq->rows, seq->wrap, std::max(seq->length, max_m), 0, seq->rows);
  if (st < 0) return st;
  LM_CHECK(limits_ok(seq->alphabet, max_m));
  for (size_t q = 0; q < n; ++q) {
    std::memset(&best[q], 0, sizeof(lm_hit));
    if (status) status[q] = LM_EMPTY;
  }
  if (st == LM_EMPTY) return LM_OK;  // no rows: every motif is "None"
  lm_batch_cache* B = nullptr;
  LM_CHECK(batch_prepare(ctx, seq, pssms, widths, offsets, n_motifs, pssm_stride, strands, true, nullptr, 0, &B));
  uint64_t n_pruned = 0, launches = 0;
  timer_begin(ctx);
  LM_CHECK(batch_launch(ctx, seq, *B, true, nullptr, strands, 0, &n_pruned, &launches));
  timer_end(ctx);
  ctx->launches += launches;
  std::vector<uint8_t> blocks(n * BATCH_BLOCK);
  LM_TRY(cudaMemcpyAsync(blocks.data(), ctx->d_out.p, blocks.size(), cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  uint64_t cand = 0, over = 0;
  for (size_t q = 0; q < n; ++q) {
    const SmallBlock* sb = (const SmallBlock*)(blocks.data() + q * BATCH_BLOCK);
    if (seq->length < B->items[q].m) continue;  // LM_EMPTY: the reference's scores.resize(0, 0) -> argmax None
    lm_colwin cw[32];
    decode_colwins(sb->ck.keys, false, cw);
    const int r = apply_rule(cw, bits_to_score(sb->ck.first_cell, false), rule, false, seq->rows, 0, &best[q]);
    if (status) status[q] = r;
    cand += sb->stats[0];
    over += sb->stats[1];
  }
  ctx->stats[0] = cand;
  ctx->stats[1] = over;
  ctx->stats[2] = n_pruned;  // number of motifs that ran the filter
  ctx->stats[3] = 0;
  return LM_OK;
}

int lm_seq_score_argmax_f32_batch(lm_ctx* ctx, const lm_seq* seq, const float* pssms, const uint32_t* widths,
                                  const uint64_t* offsets, size_t n_motifs, size_t pssm_stride, int rule,
                                  lm_hit* best) {
  return lm_seq_score_argmax_f32_batch_rc(ctx, seq, pssms, widths, offsets, n_motifs, pssm_stride, LM_STRAND_DIRECT, rule,
                                          best, nullptr);
}

int lm_seq_score_threshold_f32_batch_rc(lm_ctx* ctx, const lm_seq* seq, const float* pssms, const uint32_t* widths,
                                        const uint64_t* offsets, size_t n_motifs, size_t pssm_stride, int strands,
                                        const float* thresholds, lm_hit* hits, size_t cap, size_t* counts) {
  LM_CHECK(bind(ctx));
  if (!seq || !pssms || !widths || !offsets || !thresholds || !counts || (cap && !hits) ||
      pssm_stride < (size_t)alphabet_k(seq->alphabet) || strands < 1 || strands > 3 ||
      (strands != 1 && seq->alphabet != LM_DNA))
    return LM_ERR_INVALID_ARG;
  if (seq->device != ctx->device) return LM_ERR_INVALID_ARG;
  const size_t n = n_motifs * (strands == 3 ? 2 : 1);
  for (size_t q = 0; q < n; ++q) counts[q] = 0;
  if (n_motifs == 0) return LM_OK;
  if (n_motifs > 65536) return LM_ERR_UNSUPPORTED;
  size_t max_m = 0;
  for (size_t q = 0; q < n_motifs; ++q) max_m = std::max<size_t>(max_m, widths[q]);
  int st = check_score_args(max_m ? max_m : 1, seq->rows, seq->wrap, std::max(seq->length, max_m), 0, seq->rows);
  if (st < 0) return st;
  if (st == LM_EMPTY) return LM_OK;
  LM_CHECK(limits_ok(seq->alphabet, max_m));
  lm_batch_cache* B = nullptr;
  LM_CHECK(batch_prepare(ctx, seq, pssms, widths, offsets, n_motifs, pssm_stride, strands, false, thresholds, cap, &B));
  uint64_t n_pruned = 0, launches = 0;
  timer_begin(ctx);
  LM_CHECK(batch_launch(ctx, seq, *B, false, thresholds, strands, cap, &n_pruned, &launches));
  timer_end(ctx);
  ctx->launches += launches;
  std::vector<uint8_t> blocks(n * BATCH_BLOCK);
  LM_TRY(cudaMemcpyAsync(blocks.data(), ctx->d_out.p, blocks.size(), cudaMemcpyDeviceToHost, ctx->stream));
  LM_TRY(cudaStreamSynchronize(ctx->stream));
  bool overflow = false;
  uint64_t cand = 0, over = 0;
  size_t max_n = 0;
  for (size_t q = 0; q < n; ++q) {
    const SmallBlock* sb = (const SmallBlock*)(blocks.data() + q * BATCH_BLOCK);
    counts[q] = (size_t)sb->count;
    cand += sb->stats[0];
    over += sb->stats[1];
    overflow = overflow || counts[q] > cap;
    max_n = std::max(max_n, std::min<size_t>(counts[q], cap));
  }
  // all hit lists with ONE strided copy (rows = motifs, width = the longest list)
  if (max_n) {
    LM_TRY(cudaMemcpy2DAsync(hits, cap * sizeof(lm_hit), ctx->d_hits.p, cap * sizeof(lm_hit), max_n * sizeof(lm_hit), n,
                             cudaMemcpyDeviceToHost, ctx->stream));
    LM_TRY(cudaStreamSynchronize(ctx->stream));
  }
  for (size_t q = 0; q < n; ++q) sort_hits(hits + q * cap, std::min<size_t>(counts[q], cap));
  ctx->stats[0] = cand;
  ctx->stats[1] = over;
  ctx->stats[2] = n_pruned;
  ctx->stats[3] = 0;
  return overflow ? LM_ERR_OVERFLOW : LM_OK;
}

int lm_seq_score_threshold_f32_batch(lm_ctx* ctx, const lm_seq* seq, const float* pssms, const uint32_t* widths,
                                     const uint64_t* offsets, size_t n_motifs, size_t pssm_stride,
                                     const float* thresholds, lm_hit* hits, size_t cap, size_t* counts) {
  return lm_seq_score_threshold_f32_batch_rc(ctx, seq, pssms, widths, offsets, n_motifs, pssm_stride, LM_STRAND_DIRECT,
                                             thresholds, hits, cap, counts);
}

int lm_pssm_reverse_complement(const float* pssm, size_t m, size_t pssm_stride, float* out) {
  if (!pssm || !out || pssm_stride < 5 || pssm == out) return LM_ERR_INVALID_ARG;
  reverse_complement(pssm, m, pssm_stride, out);
  return LM_OK;
}

int lm_plan_filter(const float* pssm, size_t m, size_t pssm_stride, double threshold, int is_argmax,
                   double n_positions, int force_lookups, int force_mode, lm_filter_plan* out) {
  if (!pssm || !out || m == 0 || pssm_stride < 5) return LM_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  const double t = is_argmax ? estimate_max_score(pssm, m, pssm_stride, n_positions) : threshold;
  const PrunePlan p = plan_prune(pssm, m, pssm_stride, t, is_argmax != 0, n_positions, force_lookups, force_mode);
  out->ok = p.ok ? 1 : 0;
  out->r = p.r;
  out->pair = p.pair;
  out->pair_mul = p.pair_mul;
  out->j0 = p.j0;
  out->smax = p.smax;
  for (int k = 0; k < 16; ++k) { out->tab_lo[k] = p.tab_lo[k]; out->tab_hi[k] = p.tab_hi[k]; }
  out->c0 = p.c0;
  out->delta = p.delta;
  out->eps = p.eps;
  out->est_pass_rate = p.est_pass_rate;
  out->est_cost = p.est_cost;
  return p.r > 0 ? LM_OK : LM_EMPTY;
}

int lm_plan_filter_protein(const float* pssm, size_t m, size_t pssm_stride, double threshold, int is_argmax,
                           double n_positions, int force_lookups, lm_filter_plan* out, uint8_t class_of[21]) {
  if (!pssm || !out || !class_of || m == 0 || pssm_stride < 21) return LM_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  const double t = is_argmax ? estimate_max_score_protein(pssm, m, pssm_stride, n_positions) : threshold;
  const ProteinPlan pp = plan_prune_protein(pssm, m, pssm_stride, t, is_argmax != 0, n_positions, force_lookups);
  const PrunePlan& p = pp.plan;
  out->ok = p.ok ? 1 : 0;
  out->r = p.r;
  out->pair = 0;
  out->pair_mul = p.pair_mul;
  out->j0 = p.j0;
  out->smax = p.smax;
  for (int k = 0; k < 16; ++k) { out->tab_lo[k] = p.tab_lo[k]; out->tab_hi[k] = p.tab_hi[k]; }
  out->c0 = p.c0;
  out->delta = p.delta;
  out->eps = p.eps;
  out->est_pass_rate = p.est_pass_rate;
  out->est_cost = p.est_cost;
  for (int a = 0; a < 21; ++a) class_of[a] = pp.class_of[a];
  return p.r > 0 ? LM_OK : LM_EMPTY;
}

int lm_argmax_merge(const lm_colwin* parts, size_t n_parts, const uint64_t* row_offsets, float first_cell,
                    uint64_t total_rows, int rule, lm_hit* best) {
  if (!parts || !row_offsets || !best || n_parts == 0) return LM_ERR_INVALID_ARG;
  lm_colwin cw[32];
  for (int c = 0; c < 32; ++c) {
    cw[c].valid = 0; cw[c].row = 0; cw[c].score = 0.f;
    for (size_t p = 0; p < n_parts; ++p) {
      const lm_colwin& x = parts[p * 32 + c];
      if (!x.valid) continue;
      // later slabs hold larger rows: '>=' keeps the last row on ties
      if (!cw[c].valid || x.score >= cw[c].score) {
        cw[c].valid = 1;
        cw[c].score = x.score;
        cw[c].row = row_offsets[p] + x.row;
      }
    }
  }
  return apply_rule(cw, first_cell, rule, false, total_rows, row_offsets[0], best);
}

}  // extern "C"
