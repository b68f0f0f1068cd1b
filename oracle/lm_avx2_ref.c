/*
 * lm_avx2_ref.c -- x86-64 AVX2 restatement of the reference's fast CPU path,
 * used ONLY as the timed CPU baseline (bench.py cpu_baseline / --impl
 * reference) and cross-checked bit-for-bit against lm_oracle.c in tests/.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * The reference itself (Rust) cannot be built in this environment (no rustc /
 * cargo), so this file restates the instruction recipe of
 *   lightmotif/src/pli/platform/avx2.rs:101-199  score_f32_avx2_permute
 *     (pshufb byte->dword widening, vpermps 8-entry LUT, vaddps, NT stores)
 *   lightmotif/src/pli/platform/avx2.rs:349-426  argmax_f32_avx2
 *     (per-column running best with `<=`, strict `>` across columns)
 * for the DNA alphabet (K = 5 <= 8, PSSM rows padded to 8 floats = 32 bytes,
 * dense.rs:125-128).  The reference is single-threaded per call; the
 * `threads` argument below splits the row range the way a caller of
 * Score::score_rows_into (pli/mod.rs:72-77) would.
 */
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 32

/* Score rows [r0, r1) of a configured striped DNA sequence.
 * pssm: [m][8] f32, 32-byte aligned; seq: [>= r1 + m - 1][32] u8, 32-byte
 * aligned; out: [r1 - r0][32] f32, 32-byte aligned. */
__attribute__((target("avx2"))) void lmo_avx2_score_f32_dna(
    const float *pssm, size_t m, const uint8_t *seq, size_t r0, size_t r1,
    float *out) {
  /* byte -> dword widening masks: sub-vector q takes bytes 4q..4q+3 of each
   * 128-bit half (avx2.rs:126-144) */
  __m256i widen[4];
  for (int q = 0; q < 4; q++) {
    int b = 4 * q;
    widen[q] = _mm256_set_epi32(
        (int)0xFFFFFF03 + b, (int)0xFFFFFF02 + b, (int)0xFFFFFF01 + b, (int)0xFFFFFF00 + b,
        (int)0xFFFFFF03 + b, (int)0xFFFFFF02 + b, (int)0xFFFFFF01 + b, (int)0xFFFFFF00 + b);
  }
  for (size_t r = r0; r < r1; r++) {
    __m256 acc[4];
    for (int q = 0; q < 4; q++) acc[q] = _mm256_setzero_ps();
    const uint8_t *sp = seq + r * LANES;
    const float *pp = pssm;
    for (size_t j = 0; j < m; j++) {
      __m256i x = _mm256_load_si256((const __m256i *)sp);
      __m256 t = _mm256_load_ps(pp);
      for (int q = 0; q < 4; q++) {
        __m256i idx = _mm256_shuffle_epi8(x, widen[q]);
        acc[q] = _mm256_add_ps(acc[q], _mm256_permutevar8x32_ps(t, idx));
      }
      sp += LANES;
      pp += 8;
    }
    /* acc[q] = columns {4q..4q+3, 16+4q..16+4q+3}: put them back in order */
    float *o = out + (r - r0) * LANES;
    _mm256_stream_ps(o + 0, _mm256_permute2f128_ps(acc[0], acc[1], 0x20));
    _mm256_stream_ps(o + 8, _mm256_permute2f128_ps(acc[2], acc[3], 0x20));
    _mm256_stream_ps(o + 16, _mm256_permute2f128_ps(acc[0], acc[1], 0x31));
    _mm256_stream_ps(o + 24, _mm256_permute2f128_ps(acc[2], acc[3], 0x31));
  }
  _mm_sfence();
}

/* argmax over [rows][32] f32 with the AVX2 back-end's tie rule. */
__attribute__((target("avx2"))) int lmo_avx2_argmax_f32(const float *s,
                                                        size_t rows,
                                                        size_t *row,
                                                        size_t *col) {
  if (rows == 0) return -1;
  __m256 best[4];
  __m256i where[4];
  for (int q = 0; q < 4; q++) {
    best[q] = _mm256_load_ps(s + 8 * q);
    where[q] = _mm256_setzero_si256();
  }
  for (size_t i = 0; i < rows; i++) {
    __m256i idx = _mm256_set1_epi32((int)i);
    for (int q = 0; q < 4; q++) {
      __m256 x = _mm256_load_ps(s + i * LANES + 8 * q);
      __m256 le = _mm256_cmp_ps(best[q], x, _CMP_LE_OS);
      where[q] = _mm256_blendv_epi8(where[q], idx, _mm256_castps_si256(le));
      best[q] = _mm256_blendv_ps(best[q], x, le);
    }
  }
  uint32_t w[LANES];
  for (int q = 0; q < 4; q++) _mm256_storeu_si256((__m256i *)(w + 8 * q), where[q]);
  size_t br = 0, bc = 0;
  float bs = s[0];
  for (size_t c = 0; c < LANES; c++) {
    float x = s[(size_t)w[c] * LANES + c];
    if (x > bs) {
      bs = x;
      br = w[c];
      bc = c;
    }
  }
  *row = br;
  *col = bc;
  return 0;
}

/* ---- threaded "score then argmax" driver (the reference's usage pattern:
 * pssm.score(&striped) followed by scores.argmax(), lightmotif-bench/dna.rs) */

typedef struct {
  const float *pssm;
  size_t m;
  const uint8_t *seq;
  size_t r0, r1;
  float *scratch; /* [r1 - r0][32] */
  float best;
  size_t row, col;
  int ok;
} slab_t;

static void *slab_main(void *arg) {
  slab_t *s = (slab_t *)arg;
  s->ok = 0;
  if (s->r1 <= s->r0) return NULL;
  lmo_avx2_score_f32_dna(s->pssm, s->m, s->seq, s->r0, s->r1, s->scratch);
  size_t r, c;
  if (lmo_avx2_argmax_f32(s->scratch, s->r1 - s->r0, &r, &c) == 0) {
    s->best = s->scratch[r * LANES + c];
    s->row = s->r0 + r;
    s->col = c;
    s->ok = 1;
  }
  return NULL;
}

/* Scores rows [0, rows) with `threads` workers, each materialising its slab of
 * StripedScores in `scratch` ([rows][32] f32, caller-allocated, 32-byte
 * aligned) and reducing it.  Slab winners are merged "first column, then last
 * row" (the single-call AVX2 rule applied to the concatenation). */
int lmo_avx2_score_argmax_mt(const float *pssm, size_t m, const uint8_t *seq,
                             size_t rows, int threads, float *scratch,
                             float *best, size_t *row, size_t *col) {
  if (threads < 1) threads = 1;
  slab_t *sl = (slab_t *)calloc((size_t)threads, sizeof(slab_t));
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  size_t per = (rows + (size_t)threads - 1) / (size_t)threads;
  for (int t = 0; t < threads; t++) {
    size_t r0 = per * (size_t)t, r1 = r0 + per;
    if (r0 > rows) r0 = rows;
    if (r1 > rows) r1 = rows;
    sl[t] = (slab_t){pssm, m, seq, r0, r1, scratch + r0 * LANES, 0.f, 0, 0, 0};
    if (t + 1 < threads) pthread_create(&th[t], NULL, slab_main, &sl[t]);
  }
  slab_main(&sl[threads - 1]);
  for (int t = 0; t + 1 < threads; t++) pthread_join(th[t], NULL);
  int have = 0;
  for (int t = 0; t < threads; t++) {
    if (!sl[t].ok) continue;
    int take = !have || sl[t].best > *best ||
               (sl[t].best == *best &&
                (sl[t].col < *col || (sl[t].col == *col && sl[t].row > *row)));
    if (take) {
      have = 1;
      *best = sl[t].best;
      *row = sl[t].row;
      *col = sl[t].col;
    }
  }
  free(sl);
  free(th);
  return have ? 0 : -1;
}
