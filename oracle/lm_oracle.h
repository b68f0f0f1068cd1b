/*
 * lm_oracle.h -- CPU restatement of lightmotif's *generic* scoring pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (lightmotif_b200/, the
 * C-ABI library) may include, link or call this.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it,
 * as the checker.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the lightmotif repository root, reference @ 9df9a159 / v0.10.1).
 *
 * Parity status: PINNED for DNA f32 score / argmax / threshold, encode,
 * stripe, configure_wrap and the u8 upper-bound property, against the
 * reference's own golden vectors (tests/golden/, see tests/test_oracle_kat.py).
 * Protein scoring is parity-UNPINNED (the reference holds no value test).
 */
#ifndef LM_ORACLE_H
#define LM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LMO_DNA = 0, LMO_PROTEIN = 1 };
enum { LMO_RULE_GENERIC = 0, LMO_RULE_AVX2 = 1, LMO_RULE_SSE2 = 2 };

/* alphabet constants: lightmotif/src/abc.rs:88-110 (Dna, "ACTGN"),
 * abc.rs:190-228 (Protein, "ACDEFGHIKLMNPQRSTVWYX") */
int lmo_alphabet_k(int alphabet);
const char *lmo_alphabet_str(int alphabet);
uint8_t lmo_default_symbol(int alphabet);

/* pli/mod.rs:56-66 + abc.rs:161-171,298-324.  Returns 0, or -1 with
 * *bad_index = index of the first byte that is not in the alphabet. */
int lmo_encode(int alphabet, const uint8_t *ascii, size_t n, uint8_t *dst,
               size_t *bad_index);
/* seq.rs:122-129 */
void lmo_encode_lossy(int alphabet, const uint8_t *ascii, size_t n, uint8_t *dst);

/* pli/mod.rs:178-200.  dst is [rows][stride] with rows = ceil(n / c). */
size_t lmo_stripe_rows(size_t n, size_t c);
void lmo_stripe(const uint8_t *sym, size_t n, size_t c, uint8_t pad,
                uint8_t *dst, size_t stride);
/* seq.rs:369-381.  data has room for rows + new_wrap rows. */
void lmo_configure_wrap(uint8_t *data, size_t rows, size_t c, size_t stride,
                        size_t new_wrap, uint8_t pad);

/* pli/mod.rs:72-106 with T = f32.  Returns the number of score rows written
 * (0 when len < m or the range is empty). out is [row_end-row_begin][out_stride]. */
size_t lmo_score_f32_rows(const float *pssm, size_t m, size_t pssm_stride,
                          const uint8_t *seq, size_t seq_stride, size_t c,
                          size_t len, size_t row_begin, size_t row_end,
                          float *out, size_t out_stride);
/* same with T = u8; `saturate` != 0 is the SIMD contract
 * (avx2.rs:331 _mm256_adds_epu8), 0 is the wrapping release-mode generic. */
size_t lmo_score_u8_rows(const uint8_t *dm, size_t m, size_t dm_stride,
                         const uint8_t *seq, size_t seq_stride, size_t c,
                         size_t len, size_t row_begin, size_t row_end,
                         uint8_t *out, size_t out_stride, int saturate);

/* pli/mod.rs:135-155 (generic), avx2.rs:349-426, sse2.rs:175-261.
 * Return 0 and (row, col), or -1 when rows == 0. */
int lmo_argmax_f32(const float *s, size_t rows, size_t c, size_t stride,
                   int rule, size_t *row, size_t *col);
int lmo_argmax_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
                  int rule, size_t *row, size_t *col);
/* pli/mod.rs:158-160 (generic) / avx2.rs:428-467 (accumulators start at 0) */
int lmo_max_f32(const float *s, size_t rows, size_t c, size_t stride, int rule,
                float *out);
int lmo_max_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
               uint8_t *out);
/* pli/mod.rs:210-221: row-major, x >= t.  Writes up to cap (row, col) pairs,
 * returns the total count. */
size_t lmo_threshold_f32(const float *s, size_t rows, size_t c, size_t stride,
                         float t, uint64_t *rows_out, uint32_t *cols_out,
                         size_t cap);
size_t lmo_threshold_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
                        uint8_t t, uint64_t *rows_out, uint32_t *cols_out,
                        size_t cap);

/* --- PSSM construction (host side, restated to build identical inputs) --- */
/* pwm/mod.rs:209-236 CountMatrix::from_sequences; counts is [m][k] u32 */
void lmo_counts_from_sequences(int alphabet, const uint8_t *const *seqs,
                               size_t nseq, size_t m, uint32_t *counts);
/* pwm/mod.rs:240-258 to_freq with scalar pseudocount (abc.rs:558-574) */
void lmo_to_freq(int alphabet, const uint32_t *counts, size_t m, float pseudo,
                 float *freq);
/* pwm/mod.rs:415-430 into_scoring; bg == NULL -> uniform (abc.rs:473-487) */
void lmo_to_scoring(int alphabet, const float *freq, size_t m, const float *bg,
                    float *pssm);
/* pwm/mod.rs:597-620 */
float lmo_min_score(int alphabet, const float *pssm, size_t m);
float lmo_max_score(int alphabet, const float *pssm, size_t m);
/* pwm/mod.rs:665-696 to_discrete; offsets is [m] */
void lmo_to_discrete(int alphabet, const float *pssm, size_t m, uint8_t *dm,
                     float *factor, float *offsets, float *offset);
/* pwm/mod.rs:782-790 */
uint8_t lmo_scale(float score, float factor, float offset);
float lmo_unscale(uint8_t score, float factor, float offset);
/* pwm/mod.rs:651-662 on the striped (configured) sequence, seq.rs:433-442 */
float lmo_score_position(const float *pssm, size_t m, size_t k,
                         const uint8_t *seq, size_t seq_stride, size_t rows,
                         size_t pos);

/* --- Scanner (scan.rs:160-250) ------------------------------------------- */
/* Collects every hit (position, score) with score >= threshold exactly as
 * Scanner::next does (u8 prefilter per block, f32 recheck). Order = the
 * reference's pop order is not reproduced; hits are returned sorted by
 * position (the reference's tests sort too, tests/scan.rs:36). */
size_t lmo_scanner_collect(int alphabet, const float *pssm, size_t m,
                           const uint8_t *seq, size_t seq_stride, size_t c,
                           size_t rows, size_t wrap, size_t len,
                           float threshold, size_t block_size,
                           uint64_t *pos_out, float *score_out, size_t cap);
/* Scanner::max (scan.rs:200-249). Returns 0 and the hit, or -1 for none. */
int lmo_scanner_max(int alphabet, const float *pssm, size_t m,
                    const uint8_t *seq, size_t seq_stride, size_t c,
                    size_t rows, size_t wrap, size_t len, float threshold,
                    size_t block_size, uint64_t *pos, float *score);

/* --- synthetic inputs (no reference analogue) -------------- */
void lmo_synth_dna(uint64_t seed, uint64_t offset, size_t n, uint8_t *dst);
void lmo_synth_protein(uint64_t seed, uint64_t offset, size_t n, uint8_t *dst);

#ifdef __cplusplus
}
#endif
#endif
