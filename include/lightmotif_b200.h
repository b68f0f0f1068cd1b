/*
 * lightmotif_b200.h -- C ABI of the B200-native PWM scanning engine.
 *
 * This is the drop-in boundary for the `lightmotif::pli` pipeline traits of
 * althonos/lightmotif (reference @ 9df9a159, v0.10.1).  Every entry point
 * cites the reference interface it replaces (paths relative to the lightmotif
 * repository).  A Rust back-end (`struct Cuda; impl Backend for Cuda`) binds
 * these symbols one-to-one -- see INTEGRATION.md and bindings/rust/.
 *
 * Conventions
 *  - every function returns an `int` status: LM_OK (0), LM_EMPTY (1, the
 *    reference's `None`), or a negative LM_ERR_* code; nothing unwinds across
 *    the ABI (the reference panics, e.g. avx2.rs:832-837; the Rust shim turns
 *    LM_ERR_WRAP back into that panic);
 *  - buffers are plain pointers + sizes; host pointers are borrowed for the
 *    duration of the call only; device memory is owned by `lm_ctx` / `lm_seq`
 *    handles with explicit free;
 *  - the striped layout is the reference's: `[rows + wrap][32]` u8, row
 *    stride 32 bytes (DenseMatrix<Symbol, U32>, dense.rs:43-55,125-128),
 *    cell (r, c) = symbol at sequence index c * rows + r, default symbol
 *    (N / X) past the end (pli/mod.rs:178-200, seq.rs:369-381);
 *  - score matrices are `[rows][32]` f32 / u8 with a caller-given row stride
 *    in ELEMENTS (DenseMatrix<f32, U32>::stride() == 32);
 *  - a PSSM is `[m][pssm_stride]` with K valid columns per row (K = 5 DNA,
 *    21 protein; DenseMatrix<f32, K> stride is 8 / 24 floats, <u8, K> is 32).
 *    Columns >= K are never read (they are uninitialised padding in Rust);
 *  - an `lm_ctx` is NOT thread-safe: use one per host thread (the reference's
 *    `Pipeline` is stateless and Clone, pli/mod.rs:227-231).  Different
 *    contexts may be used concurrently, also on the same device.
 */
#ifndef LIGHTMOTIF_B200_H
#define LIGHTMOTIF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_ABI_VERSION 2
#define LM_LANES 32 /* <Avx2 as Backend>::Lanes = U32, avx2.rs:33-35 */

/* status codes */
#define LM_OK 0
#define LM_EMPTY 1               /* Option::None (empty scores) */
#define LM_ERR_INVALID_ARG (-1)
#define LM_ERR_INVALID_SYMBOL (-2) /* err.rs:10 InvalidSymbol; *bad_index is set */
#define LM_ERR_WRAP (-3)           /* avx2.rs:832-837 "not enough wrapping rows" */
#define LM_ERR_CUDA (-4)
#define LM_ERR_OOM (-5)
#define LM_ERR_UNSUPPORTED (-6)    /* err.rs:34 UnsupportedBackend / size limits */
#define LM_ERR_OVERFLOW (-7)       /* hit buffer too small; *count holds the total */

/* alphabets (abc.rs:88-110 Dna "ACTGN", abc.rs:190-228 Protein "ACDEFGHIKLMNPQRSTVWYX") */
#define LM_DNA 0
#define LM_PROTEIN 1

/* argmax tie rules: which reference back-end to reproduce */
#define LM_RULE_GENERIC 0 /* pli/mod.rs:135-155: last row, then last column */
#define LM_RULE_AVX2 1    /* avx2.rs:349-426: first column, last row in it; (0,0) quirk */
#define LM_RULE_SSE2 2    /* sse2.rs:175-261: last column, last row in it */

typedef struct lm_ctx lm_ctx;
typedef struct lm_seq lm_seq;

/* MatrixCoordinates (dense.rs:28-39) plus the score of the cell.
 * `position` = col * rows + row = StripedScores::offset (scores.rs:155-157),
 * computed with the number of score rows of the call. */
typedef struct lm_hit {
  uint64_t row;
  uint64_t position;
  uint32_t col;
  float score; /* for u8 calls: the u8 score converted to float */
} lm_hit;

/* per-column winner: the highest score of a column and the LAST row that
 * attains it (what avx2.rs:373-403 keeps in s1..s4 / p1..p4).  `valid` is 0
 * for an empty row range. */
typedef struct lm_colwin {
  uint64_t row;
  float score;
  uint32_t valid;
} lm_colwin;

/* ------------------------------------------------------------- lifecycle */

int lm_abi_version(void);
const char *lm_strerror(int status);
/* message of the last failing CUDA call on this context ("" if none) */
const char *lm_ctx_last_error(const lm_ctx *ctx);

/* Pipeline::<A, Cuda>::cuda() -> Result<Self, UnsupportedBackend>
 * (mirrors pli/mod.rs:401-407 Pipeline::avx2()).  LM_ERR_UNSUPPORTED when no
 * sm_100 device is present. */
int lm_ctx_create(int device, lm_ctx **out);
/* Same, but enqueue all work on an existing CUDA stream (cudaStream_t) so the
 * caller can bracket calls with its own events. */
int lm_ctx_create_on_stream(int device, void *cuda_stream, lm_ctx **out);
void lm_ctx_destroy(lm_ctx *ctx);
void *lm_ctx_stream(lm_ctx *ctx);
int lm_ctx_synchronize(lm_ctx *ctx);
/* number of kernels this library launched through `ctx` so far */
uint64_t lm_ctx_kernel_launches(const lm_ctx *ctx);
/* device time (ms, CUDA events on the context's stream) spent in the kernels
 * of the most recent scoring call */
float lm_ctx_last_kernel_ms(const lm_ctx *ctx);
/* SM clock (MHz) the most recent filtered fused scan actually ran at: SM cycle
 * counter / global nanosecond timer read inside the kernel (0 when unknown).
 * Measured on the device because NVML queries from the host stall concurrent
 * multi-GPU work on some hosts. */
float lm_ctx_last_sm_clock_mhz(const lm_ctx *ctx);
/* Options.  LM_OPT_PRUNE controls how the FUSED calls (score+argmax,
 * score+threshold; DNA) find their result: 0 = score every cell exactly,
 * 1 (default) = exact filter + recheck when the cost model expects it to pay
 * off (an integer upper bound of every cell's score is computed, only cells
 * whose bound reaches the running threshold get the exact sequential f32 sum
 * -- the device analogue of the reference's Scanner, scan.rs:160-250),
 * 2 = filter whenever a plan exists.  Results are bit-identical in all modes. */
#define LM_OPT_PRUNE 1
/* rows scored exactly first to seed the running maximum of a filtered argmax
 * scan (default 262144) */
#define LM_OPT_BOOT_ROWS 2
/* force the number of filter look-ups per position (single-row mode: 4, 6, 8,
 * 10, 12 or 16; paired mode: 3, 4, 5, 6 or 8; 0 = planner decides) */
#define LM_OPT_FILTER_ROWS 3
/* leave this many SMs out of the persistent scoring grids so that a concurrent
 * small kernel (the NCCL all-gather of the previous scan's winners) can run
 * next to a scan (default 0) */
#define LM_OPT_RESERVE_SMS 4
/* filter look-up flavour: 0 = planner decides, 1 = one motif row per look-up,
 * 2 = paired (one row exactly + one class bit of the next row per look-up) */
#define LM_OPT_FILTER_MODE 5
/* protein filter: 1 (default) = keep the class-coded copy of the last scanned
 * sequence and reuse it while sequence and class map are unchanged, 0 = rebuild
 * it on every call */
#define LM_OPT_CLASS_CACHE 6
/* lm_seq_from_striped[_slab] / lm_seq_from_symbols of a large DNA matrix or
 * symbol array (pinned or pageable host memory).  Two-ended upload: host
 * threads pack chunks from the back of the matrix 4 symbols per byte (chunks
 * holding N go as they are), a kernel expands them in place, and the copy
 * engine pulls plain chunks from the front meanwhile -- the device matrix is
 * byte-identical, the call is no longer bound by 1 B/symbol over PCIe.
 * 1 (default) = adaptive: the context times its large uploads both ways (the
 * first two-ended, the second plain) and keeps the faster one, re-trying the
 * other every 32nd upload -- several ranks sharing one host can saturate its
 * memory with plain copies alone; 0 = always the plain copy; 2 = packed chunks
 * only; 3 = two-ended, plain chunks only while no packed copy is pending;
 * 4 = always two-ended. */
#define LM_OPT_UPLOAD 7
/* host threads that pack (0 = default: all hardware threads but one) */
#define LM_OPT_UPLOAD_THREADS 8
/* matrices smaller than this many bytes take the plain copy (default 64 MiB) */
#define LM_OPT_UPLOAD_MIN_BYTES 9
int lm_ctx_set_option(lm_ctx *ctx, int option, int value);
/* the most recent lm_seq_from_striped[_slab] / lm_seq_from_symbols: out[0] = bytes that crossed PCIe,
 * out[1] = chunks sent packed, out[2] = chunks sent plain, out[3] = host
 * threads used (0: plain copy) */
int lm_ctx_last_upload(const lm_ctx *ctx, uint64_t out[4]);
/* statistics of the most recent fused call: out[0] = cells rechecked exactly,
 * out[1] = tiles that fell back to exact scoring, out[2] = 1 if the filter
 * path ran, out[3] = filter look-ups per position | (paired mode << 8) */
int lm_ctx_last_stats(const lm_ctx *ctx, uint64_t out[4]);

/* --------------------------------------------- host-buffer trait mirrors */

/* Encode::encode_into (pli/mod.rs:56-66; avx2.rs:37-99).  On an invalid byte
 * returns LM_ERR_INVALID_SYMBOL with *bad_index = index of the FIRST invalid
 * byte (the Rust shim rebuilds InvalidSymbol(seq[bad_index] as char)). */
int lm_encode(lm_ctx *ctx, int alphabet, const uint8_t *ascii, size_t n,
              uint8_t *dst, size_t *bad_index);
/* EncodedSequence::encode_lossy (seq.rs:122-129): unknown -> default symbol */
int lm_encode_lossy(lm_ctx *ctx, int alphabet, const uint8_t *ascii, size_t n,
                    uint8_t *dst);
/* Stripe::stripe_into (pli/mod.rs:178-200; avx2.rs:561-799).  dst is
 * [ceil(n/32)][dst_stride] bytes; cells past n get the default symbol. */
int lm_stripe(lm_ctx *ctx, int alphabet, const uint8_t *sym, size_t n,
              uint8_t *dst, size_t dst_stride);
/* StripedSequence::configure_wrap (seq.rs:369-381): fills rows
 * [rows, rows + wrap) of `data` ([rows + wrap][stride]). */
int lm_configure_wrap(lm_ctx *ctx, int alphabet, uint8_t *data, size_t rows,
                      size_t stride, size_t wrap);

/* Score<f32, A, U32>::score_rows_into (pli/mod.rs:72-106; avx2.rs:817-904).
 * seq is the host striped matrix [rows + wrap][seq_stride]; `length` is
 * StripedSequence::len().  Writes *out_rows = 0 and returns LM_OK when
 * length < m or the range is empty (the reference's resize(0, 0)); otherwise
 * fills out[0 .. row_end-row_begin][32] and *out_rows = row_end - row_begin.
 * LM_ERR_WRAP if wrap < m - 1. */
int lm_score_f32_rows(lm_ctx *ctx, int alphabet, const float *pssm, size_t m,
                      size_t pssm_stride, const uint8_t *seq, size_t rows,
                      size_t wrap, size_t seq_stride, size_t length,
                      size_t row_begin, size_t row_end, float *out,
                      size_t out_stride, size_t *out_rows);
/* Score<u8, Dna, U32>::score_rows_into (avx2.rs:292-347, 907-941): saturating
 * u8 sums (_mm256_adds_epu8). */
int lm_score_u8_rows(lm_ctx *ctx, int alphabet, const uint8_t *dm, size_t m,
                     size_t dm_stride, const uint8_t *seq, size_t rows,
                     size_t wrap, size_t seq_stride, size_t length,
                     size_t row_begin, size_t row_end, uint8_t *out,
                     size_t out_stride, size_t *out_rows);

/* Maximum::argmax / max (pli/mod.rs:135-160; avx2.rs:349-467) over a host
 * score matrix [rows][stride].  LM_EMPTY when rows == 0.  lm_max_* return the
 * true maximum (the generic definition); the AVX2 `max` quirk (accumulators
 * start at 0.0, avx2.rs:438-441) is NOT reproduced. */
int lm_argmax_f32(lm_ctx *ctx, const float *scores, size_t rows, size_t stride,
                  int rule, lm_hit *best);
int lm_max_f32(lm_ctx *ctx, const float *scores, size_t rows, size_t stride,
               float *best);
/* Maximum::max as a given reference back-end computes it: LM_RULE_GENERIC / LM_RULE_SSE2 (which has
 * no SIMD max and uses the generic one, dispatch.rs:169-177) = the true maximum; LM_RULE_AVX2 =
 * max(0.0, maximum), because the AVX2 accumulators start at 0.0 (avx2.rs:438-441) -- what
 * Pipeline::dispatch() returns on an x86-64 host with AVX2 for an all-negative score matrix. */
int lm_max_f32_rule(lm_ctx *ctx, const float *scores, size_t rows, size_t stride,
                    int rule, float *best);
int lm_argmax_u8(lm_ctx *ctx, const uint8_t *scores, size_t rows, size_t stride,
                 int rule, lm_hit *best);
int lm_max_u8(lm_ctx *ctx, const uint8_t *scores, size_t rows, size_t stride,
              uint8_t *best);
/* Threshold::threshold (pli/mod.rs:210-221): every cell >= t, returned in the
 * reference's row-major order.  *count is always the total number of hits;
 * LM_ERR_OVERFLOW if it exceeds cap (the first `cap` hits are still written). */
int lm_threshold_f32(lm_ctx *ctx, const float *scores, size_t rows,
                     size_t stride, float t, lm_hit *hits, size_t cap,
                     size_t *count);
int lm_threshold_u8(lm_ctx *ctx, const uint8_t *scores, size_t rows,
                    size_t stride, uint8_t t, lm_hit *hits, size_t cap,
                    size_t *count);

/* ------------------------------------------- device-resident sequences */

/* EncodedSequence::encode / encode_lossy -> to_striped -> configure_wrap in
 * ONE device pass from raw ASCII (seq.rs:111-129,168-175,369-381).
 * lossy == 0: LM_ERR_INVALID_SYMBOL + *bad_index as lm_encode. */
int lm_seq_from_ascii(lm_ctx *ctx, int alphabet, const uint8_t *ascii, size_t n,
                      int lossy, size_t wrap, lm_seq **out, size_t *bad_index);
/* Stripe::stripe + configure_wrap from already encoded symbols */
int lm_seq_from_symbols(lm_ctx *ctx, int alphabet, const uint8_t *sym, size_t n,
                        size_t wrap, lm_seq **out);
/* upload an existing host StripedSequence ([rows + wrap][stride]) */
int lm_seq_from_striped(lm_ctx *ctx, int alphabet, const uint8_t *data,
                        size_t rows, size_t wrap, size_t stride, size_t length,
                        lm_seq **out);
/* upload a ROW SLAB of a longer host StripedSequence: `data` points at row
 * `row_offset` of a striped matrix with `rows_global` = ceil(length / 32) rows;
 * the slab holds `rows` rows followed by `halo` rows (the NEXT rows of that
 * matrix, or its wrap rows for the last slab).  Hits / positions of scans over
 * the slab are reported in the coordinates of the whole matrix; the handle can
 * be passed to the lm_comm_* calls of the rank that owns it. */
int lm_seq_from_striped_slab(lm_ctx *ctx, int alphabet, const uint8_t *data,
                             size_t rows, size_t halo, size_t stride, size_t length,
                             uint64_t row_offset, uint64_t rows_global, lm_seq **out);
/* adopt caller-owned DEVICE memory holding [rows + wrap][32] bytes (16-byte
 * aligned, row stride 32); the handle does not free it */
int lm_seq_from_device(lm_ctx *ctx, int alphabet, void *device_ptr, size_t rows,
                       size_t wrap, size_t length, lm_seq **out);
/* benchmark input: counter-based synthetic sequence generated on the device
 * (SURVEY.md 8d).  DNA: w = splitmix64(seed + (i >> 5)),
 * sym(i) = (w >> 2(i & 31)) & 3; protein: (splitmix64(seed + i) >> 11) % 20.
 * The handle holds the row slab [row_begin, row_begin + n_rows) (+ wrap rows)
 * of the striped matrix of the whole `total_length`-symbol sequence, so each
 * GPU of a sharded scan generates its own slab without any transfer; pass
 * row_begin = 0, n_rows = ceil(total_length / 32) for the whole sequence. */
int lm_seq_synthetic(lm_ctx *ctx, int alphabet, uint64_t seed,
                     uint64_t total_length, uint64_t row_begin, uint64_t n_rows,
                     size_t wrap, lm_seq **out);
/* StripedSequence::configure_wrap on the device copy (only grows) */
int lm_seq_configure_wrap(lm_ctx *ctx, lm_seq *seq, size_t wrap);
int lm_seq_info(const lm_seq *seq, int *alphabet, size_t *rows, size_t *wrap,
                size_t *length);
void *lm_seq_device_ptr(const lm_seq *seq);
/* copy [rows + wrap][32] back into dst ([rows + wrap][dst_stride]) */
int lm_seq_download(lm_ctx *ctx, const lm_seq *seq, uint8_t *dst,
                    size_t dst_stride);
void lm_seq_free(lm_ctx *ctx, lm_seq *seq);

/* Score::score_rows_into on a resident sequence.  `out` is a host pointer, or
 * a device pointer when out_is_device != 0. */
int lm_seq_score_f32(lm_ctx *ctx, const lm_seq *seq, const float *pssm,
                     size_t m, size_t pssm_stride, size_t row_begin,
                     size_t row_end, float *out, size_t out_stride,
                     int out_is_device, size_t *out_rows);
int lm_seq_score_u8(lm_ctx *ctx, const lm_seq *seq, const uint8_t *dm, size_t m,
                    size_t dm_stride, size_t row_begin, size_t row_end,
                    uint8_t *out, size_t out_stride, int out_is_device,
                    size_t *out_rows);

/* FUSED  ScoringMatrix::score (pwm/mod.rs:640-648) + StripedScores::argmax
 * (scores.rs:190-192): scores never reach HBM.  Returns the cell the chosen
 * reference back-end would return, bit-identical score.  `per_column`
 * (nullable, [32]) and `first_cell` (nullable; score of cell (row_begin, 0))
 * receive what lm_argmax_merge needs for multi-GPU merging.  With LM_OPT_PRUNE
 * active, per_column[c] is the exact column winner whenever it attains the
 * range maximum and a valid lower bound otherwise -- which is all the tie
 * rules and lm_argmax_merge depend on.
 * LM_EMPTY when length < m or the range is empty. */
int lm_seq_score_argmax_f32(lm_ctx *ctx, const lm_seq *seq, const float *pssm,
                            size_t m, size_t pssm_stride, size_t row_begin,
                            size_t row_end, int rule, lm_hit *best,
                            lm_colwin *per_column, float *first_cell);
/* Split form of lm_seq_score_argmax_f32 for pipelining: `launch` only enqueues
 * the kernels and the copy of the small result block (0 <= slot < LM_SLOTS, so
 * several calls can be in flight: e.g. the scan of motif k+1 runs while the host merges
 * / all-gathers the winners of motif k); `finish` waits for that slot and
 * returns exactly what the one-shot call returns. */
#define LM_SLOTS 4
int lm_seq_score_argmax_f32_launch(lm_ctx *ctx, const lm_seq *seq,
                                   const float *pssm, size_t m,
                                   size_t pssm_stride, size_t row_begin,
                                   size_t row_end, int slot);
int lm_seq_score_argmax_f32_finish(lm_ctx *ctx, int slot, int rule,
                                   lm_hit *best, lm_colwin *per_column,
                                   float *first_cell);
/* Device address of the 272-byte result block a `launch` on `slot` writes
 * (valid once the launch's kernels have run, in stream order): uint64
 * key[32] -- per-column winner, (order-preserving f32 key << 32) | row relative
 * to row_begin, 0 = no valid cell -- then uint32 first_cell (f32 bits of cell
 * (row_begin, 0)) and 12 bytes of padding.  Lets a multi-GPU host hand the
 * winners to NCCL (all-gather) without a round trip through host memory;
 * decode with lm_decode_keys, merge with lm_argmax_merge. */
void *lm_ctx_result_block(lm_ctx *ctx, int slot);
/* keys[32] (as stored in a result block) -> lm_colwin[32] */
void lm_decode_keys(const uint64_t *keys, lm_colwin *out);
/* FUSED score + Threshold::threshold (scores.rs:207-213): all cells >= t of
 * the scored rows, row-major order, exact f32 scores. */
int lm_seq_score_threshold_f32(lm_ctx *ctx, const lm_seq *seq,
                               const float *pssm, size_t m, size_t pssm_stride,
                               size_t row_begin, size_t row_end, float t,
                               lm_hit *hits, size_t cap, size_t *count);
/* Many motifs over one resident sequence (the lightmotif-cli job loop,
 * lightmotif-cli/src/main.rs:540-562): motif i is pssms[offsets[i] ..] with
 * widths[i] rows of pssm_stride floats; best[i] as lm_seq_score_argmax_f32.
 * The sequence needs wrap >= max(widths) - 1. */
int lm_seq_score_argmax_f32_batch(lm_ctx *ctx, const lm_seq *seq,
                                  const float *pssms, const uint32_t *widths,
                                  const uint64_t *offsets, size_t n_motifs,
                                  size_t pssm_stride, int rule, lm_hit *best);
/* The same with strand selection (DNA; pwm/mod.rs:566-577 ScoringMatrix::reverse_complement: rows
 * reversed, A<->T, C<->G): LM_STRAND_DIRECT scans the matrices as given, LM_STRAND_REVERSE their
 * reverse complements, LM_STRAND_BOTH both -- results are then interleaved, best[2i] = direct,
 * best[2i + 1] = reverse complement of motif i.  status (nullable, same length as best) receives
 * LM_OK or LM_EMPTY (motif longer than the sequence: the reference's `None`) per entry. */
#define LM_STRAND_DIRECT 1
#define LM_STRAND_REVERSE 2
#define LM_STRAND_BOTH 3
int lm_seq_score_argmax_f32_batch_rc(lm_ctx *ctx, const lm_seq *seq,
                                     const float *pssms, const uint32_t *widths,
                                     const uint64_t *offsets, size_t n_motifs,
                                     size_t pssm_stride, int strands, int rule,
                                     lm_hit *best, int *status);
/* out = reverse complement of a DNA matrix ([m][pssm_stride], columns A,C,T,G,N); host only */
int lm_pssm_reverse_complement(const float *pssm, size_t m, size_t pssm_stride, float *out);
/* Same job loop with one threshold per motif (what the CLI computes from a
 * p-value, lightmotif-cli/src/main.rs:468-497): hits of motif i go to
 * hits[i * cap ..] in row-major order, counts[i] = total number of hits of
 * motif i.  LM_ERR_OVERFLOW if any count exceeds cap (the first `cap` hits of
 * those motifs are still written).  All scans are enqueued back to back and
 * the host synchronises once. */
int lm_seq_score_threshold_f32_batch(lm_ctx *ctx, const lm_seq *seq,
                                     const float *pssms, const uint32_t *widths,
                                     const uint64_t *offsets, size_t n_motifs,
                                     size_t pssm_stride, const float *thresholds,
                                     lm_hit *hits, size_t cap, size_t *counts);

/* with strand selection: hits / counts of entry e (= motif for one strand, 2 * motif + strand for
 * LM_STRAND_BOTH) at hits[e * cap ..] / counts[e]; thresholds stay one per motif */
int lm_seq_score_threshold_f32_batch_rc(lm_ctx *ctx, const lm_seq *seq,
                                        const float *pssms, const uint32_t *widths,
                                        const uint64_t *offsets, size_t n_motifs,
                                        size_t pssm_stride, int strands,
                                        const float *thresholds, lm_hit *hits,
                                        size_t cap, size_t *counts);

/* The filter the fused DNA calls would use for `pssm` (host only, no GPU
 * needed): introspection for tests and tuning.  The kernel computes, per
 * position p, sum_q(p) = sum over look-ups k < r of table_k[entry_k(p)] with
 *   single mode (pair == 0): entry_k = symbol at motif row j0 + k
 *                            (table = tab_lo bytes A,C,T,G | tab_hi byte 0 = N);
 *   paired mode:             entry_k = s | class(s') << 2, s / s' = symbols at
 *                            rows j0 + 2k / j0 + 2k + 1, class(x) = bit 2 of
 *                            x * pair_mul, N as first symbol = entry 4
 *                            (entries 0..3 = tab_lo bytes, 4..7 = tab_hi bytes);
 * and guarantees  c0 + delta * sum_q(p) >= score(p)  in real arithmetic, with
 * |f32 sequential score - real score| <= eps.  A position is rechecked exactly
 * iff sum_q(p) >= ceil((T - eps - c0) / delta - 1e-6) for the current threshold T.
 * `threshold` is the planning threshold (is_argmax != 0: ignored, the expected
 * maximum over n_positions uniform-random positions is used). */
typedef struct lm_filter_plan {
  int ok;       /* 1 if the cost model prefers the filter over exact scoring */
  int r;        /* look-ups per position */
  int pair;     /* 0 single rows, 1 paired rows */
  uint32_t pair_mul;
  int j0;
  int smax;     /* largest attainable sum_q (<= 127) */
  uint32_t tab_lo[16], tab_hi[16];
  double c0, delta, eps;
  double est_pass_rate; /* under a uniform ACGT background at the planning threshold */
  double est_cost;      /* modelled issue slots per 128 positions */
} lm_filter_plan;
int lm_plan_filter(const float *pssm, size_t m, size_t pssm_stride,
                   double threshold, int is_argmax, double n_positions,
                   int force_lookups, int force_mode, lm_filter_plan *out);
/* Protein flavour: the 21 residues are partitioned into 8 classes chosen for
 * this matrix (class_of[a], a = residue code 0..20, X = 20); the filter runs on
 * a class-coded copy of the sequence with single-row look-ups,
 * entry_k = class of the residue at motif row j0 + k (entries 0..3 = tab_lo
 * bytes, 4..7 = tab_hi bytes), table entry = best weight within the class. */
int lm_plan_filter_protein(const float *pssm, size_t m, size_t pssm_stride,
                           double threshold, int is_argmax, double n_positions,
                           int force_lookups, lm_filter_plan *out,
                           uint8_t class_of[21]);

/* Merge per-column winners of `n_parts` consecutive row slabs (slab p scored
 * rows [row_offsets[p], row_offsets[p+1])) into the cell the single-call
 * argmax would return (multi-GPU reduce after an all-gather; host only).
 * first_cell = score of cell (row_offsets[0], 0), needed by LM_RULE_AVX2.
 * total_rows is the number of score rows used for `position`. */
int lm_argmax_merge(const lm_colwin *parts, size_t n_parts,
                    const uint64_t *row_offsets, float first_cell,
                    uint64_t total_rows, int rule, lm_hit *best);

/* ------------------------------------------------ multi-GPU (row-sharded) scans
 *
 * The reference has no multi-device path; its sharding axis is the row range of
 * Score::score_rows_into (pli/mod.rs:72-77).  A communicator couples `world`
 * ranks (one GPU each): rank g scans rows [rows * g / world, rows * (g+1) / world)
 * of the global striped matrix (+ halo rows).  During a fused argmax scan the GPUs
 * share their running maximum through peer-mapped HBM (system-scope atomicMax over
 * NVLink), so every GPU filters at the global maximum; at the end of a scan the
 * scan kernel itself writes the rank's result block into every rank's memory and
 * raises an arrival flag.  No collective library is involved.  Results are the
 * cell / the hit list ONE call over the whole matrix would return (same tie rules,
 * positions = col * rows_global + row).
 *
 * lm_comm_init:        ONE process drives `n_dev` GPUs (devices[i] or 0..n_dev-1;
 *                      the same device may be listed twice, e.g. for testing).
 *                      LM_ERR_UNSUPPORTED when two devices cannot address each
 *                      other's memory.
 * lm_comm_init_rank:   one process per GPU: creates this rank's exchange buffer on
 *                      ctx's device and returns its cudaIpc handle; exchange the
 *                      handles by any means (MPI, torch.distributed, a file) and
 *                      pass all `world` of them, in rank order, to lm_comm_connect.
 * All ranks must issue the same sequence of sharded calls (like any collective).
 * `hit_capacity` = hits per rank a sharded threshold scan can hold (0: 1 << 20). */
#define LM_MAX_RANKS 16
#define LM_COMM_HANDLE_BYTES 64
typedef struct lm_comm lm_comm;
int lm_comm_init(int n_dev, const int *devices, size_t hit_capacity, lm_comm **out);
int lm_comm_init_rank(lm_ctx *ctx, int rank, int world, size_t hit_capacity,
                      uint8_t handle_out[LM_COMM_HANDLE_BYTES], lm_comm **out);
int lm_comm_connect(lm_comm *comm, const uint8_t *handles /* [world][LM_COMM_HANDLE_BYTES] */);
void lm_comm_destroy(lm_comm *comm);
const char *lm_comm_last_error(const lm_comm *comm);
int lm_comm_world(const lm_comm *comm);
/* ranks driven by this process (n_dev for lm_comm_init, 1 for lm_comm_init_rank) */
int lm_comm_n_local(const lm_comm *comm);
lm_ctx *lm_comm_ctx(lm_comm *comm, int local_index);
int lm_comm_rank(const lm_comm *comm, int local_index);
/* Row slabs of the local ranks: slabs[i] (i < lm_comm_n_local) is created on
 * lm_comm_ctx(comm, i).  From a host StripedSequence ([rows + wrap][stride]; the
 * halo of a slab is the next `wrap` rows of that matrix) or from the synthetic
 * generator (lm_seq_synthetic).  Free each with lm_seq_free(lm_comm_ctx(comm, i), .). */
int lm_comm_shard_from_striped(lm_comm *comm, int alphabet, const uint8_t *data,
                               size_t rows, size_t wrap, size_t stride,
                               size_t length, lm_seq **slabs);
int lm_comm_seq_synthetic(lm_comm *comm, int alphabet, uint64_t seed,
                          uint64_t total_length, size_t wrap, lm_seq **slabs);
/* Row-sharded lm_seq_score_argmax_f32 (ScoringMatrix::score + StripedScores::argmax
 * over the whole matrix).  launch enqueues the scans of the local ranks on their
 * contexts' streams (0 <= slot < LM_SLOTS, several may be in flight); finish waits
 * for the blocks of ALL ranks and returns the global cell on every rank. */
int lm_comm_score_argmax_f32_launch(lm_comm *comm, const lm_seq *const *slabs,
                                    const float *pssm, size_t m, size_t pssm_stride,
                                    int slot);
int lm_comm_score_argmax_f32_finish(lm_comm *comm, int slot, int rule, lm_hit *best);
/* launch + finish on slot 0 */
int lm_seq_score_argmax_f32_sharded(lm_comm *comm, const lm_seq *const *slabs,
                                    const float *pssm, size_t m, size_t pssm_stride,
                                    int rule, lm_hit *best);
/* Row-sharded lm_seq_score_threshold_f32: every rank returns the COMPLETE hit list
 * (rows / positions of the global matrix, row-major order, pli/mod.rs:210-221):
 * the per-rank lists are gathered over NVLink straight out of the ranks' exchange
 * buffers.  *count = total number of hits, local_counts (nullable, [world]) = hits
 * found by each rank.  LM_ERR_OVERFLOW if *count > cap or a rank found more than
 * the communicator's hit_capacity. */
int lm_seq_score_threshold_f32_sharded(lm_comm *comm, const lm_seq *const *slabs,
                                       const float *pssm, size_t m, size_t pssm_stride,
                                       float t, lm_hit *hits, size_t cap, size_t *count,
                                       size_t *local_counts);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTMOTIF_B200_H */
