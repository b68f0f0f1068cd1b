/*
 * lm_oracle.c -- CPU restatement of lightmotif's generic pipeline (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see lm_oracle.h).  Build with
 *   gcc -O2 -fno-fast-math -ffp-contract=off
 * so that every f32 sum keeps the reference's strictly sequential order.
 *
 * The loops below follow the *generic* (scalar) trait defaults of
 * lightmotif/src/pli/mod.rs; the SIMD back-ends of the reference are
 * bit-identical to them on every f32 score (sequential adds from +0.0) and
 * differ only in the argmax tie rule and the u8 saturation, both modelled
 * here behind a `rule` / `saturate` argument.
 */
#include "lm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- alphabets */

static const char DNA_STR[] = "ACTGN";                      /* abc.rs:107-109 */
static const char PROTEIN_STR[] = "ACDEFGHIKLMNPQRSTVWYX";  /* abc.rs:225-227 */

int lmo_alphabet_k(int alphabet) { return alphabet == LMO_DNA ? 5 : 21; }

const char *lmo_alphabet_str(int alphabet) {
  return alphabet == LMO_DNA ? DNA_STR : PROTEIN_STR;
}

/* abc.rs:128-135 (#[default] N = 4), abc.rs:254-256 (#[default] X = 20) */
uint8_t lmo_default_symbol(int alphabet) {
  return (uint8_t)(lmo_alphabet_k(alphabet) - 1);
}

/* Symbol::from_ascii, abc.rs:161-171 and abc.rs:298-324: exact upper-case
 * match only; returns K on failure. */
static int from_ascii(int alphabet, uint8_t ch) {
  const char *s = lmo_alphabet_str(alphabet);
  int k = lmo_alphabet_k(alphabet);
  for (int a = 0; a < k; a++)
    if ((uint8_t)s[a] == ch) return a;
  return k;
}

int lmo_encode(int alphabet, const uint8_t *ascii, size_t n, uint8_t *dst,
               size_t *bad_index) {
  int k = lmo_alphabet_k(alphabet);
  for (size_t i = 0; i < n; i++) { /* pli/mod.rs:62-64 */
    int a = from_ascii(alphabet, ascii[i]);
    if (a == k) {
      if (bad_index) *bad_index = i;
      return -1;
    }
    dst[i] = (uint8_t)a;
  }
  return 0;
}

void lmo_encode_lossy(int alphabet, const uint8_t *ascii, size_t n, uint8_t *dst) {
  int k = lmo_alphabet_k(alphabet);
  for (size_t i = 0; i < n; i++) { /* seq.rs:123-128: unwrap_or_default */
    int a = from_ascii(alphabet, ascii[i]);
    dst[i] = (uint8_t)(a == k ? k - 1 : a);
  }
}

/* ------------------------------------------------------------------- stripe */

size_t lmo_stripe_rows(size_t n, size_t c) { return n / c + (n % c > 0); }

void lmo_stripe(const uint8_t *sym, size_t n, size_t c, uint8_t pad,
                uint8_t *dst, size_t stride) {
  size_t rows = lmo_stripe_rows(n, c); /* pli/mod.rs:182 */
  if (rows == 0) return;
  for (size_t i = 0; i < n; i++) /* pli/mod.rs:191-193 */
    dst[(i % rows) * stride + i / rows] = sym[i];
  for (size_t i = n; i < rows * c; i++) /* pli/mod.rs:194-196 */
    dst[(i % rows) * stride + i / rows] = pad;
}

void lmo_configure_wrap(uint8_t *data, size_t rows, size_t c, size_t stride,
                        size_t new_wrap, uint8_t pad) {
  /* seq.rs:373-378; the cascade (row i may itself be a wrap row when
   * new_wrap > rows) is kept on purpose. */
  for (size_t i = 0; i < new_wrap; i++) {
    for (size_t j = 0; j + 1 < c; j++)
      data[(rows + i) * stride + j] = data[i * stride + j + 1];
    data[(rows + i) * stride + c - 1] = pad;
  }
}

/* -------------------------------------------------------------------- score */

size_t lmo_score_f32_rows(const float *pssm, size_t m, size_t pssm_stride,
                          const uint8_t *seq, size_t seq_stride, size_t c,
                          size_t len, size_t row_begin, size_t row_end,
                          float *out, size_t out_stride) {
  if (len < m || row_end <= row_begin) return 0; /* pli/mod.rs:85-88 */
  for (size_t r = row_begin; r < row_end; r++) { /* pli/mod.rs:96-105 */
    for (size_t col = 0; col < c; col++) {
      volatile float score = 0.0f; /* T::default() */
      for (size_t j = 0; j < m; j++) {
        uint8_t s = seq[(r + j) * seq_stride + col];
        score = score + pssm[j * pssm_stride + s];
      }
      out[(r - row_begin) * out_stride + col] = score;
    }
  }
  return row_end - row_begin;
}

size_t lmo_score_u8_rows(const uint8_t *dm, size_t m, size_t dm_stride,
                         const uint8_t *seq, size_t seq_stride, size_t c,
                         size_t len, size_t row_begin, size_t row_end,
                         uint8_t *out, size_t out_stride, int saturate) {
  if (len < m || row_end <= row_begin) return 0;
  for (size_t r = row_begin; r < row_end; r++) {
    for (size_t col = 0; col < c; col++) {
      unsigned score = 0;
      for (size_t j = 0; j < m; j++) {
        uint8_t s = seq[(r + j) * seq_stride + col];
        score += dm[j * dm_stride + s];
        if (saturate) { /* avx2.rs:331, neon.rs vqaddq_u8 */
          if (score > 255u) score = 255u;
        } else { /* pli/mod.rs:101 `+=` on u8, release build wraps */
          score &= 255u;
        }
      }
      out[(r - row_begin) * out_stride + col] = (uint8_t)score;
    }
  }
  return row_end - row_begin;
}

/* ------------------------------------------------------------ argmax / max */

int lmo_argmax_f32(const float *s, size_t rows, size_t c, size_t stride,
                   int rule, size_t *row, size_t *col) {
  if (rows == 0) return -1;
  if (rule == LMO_RULE_GENERIC) {
    /* pli/mod.rs:140-152: row-major scan, `>=` => last (row, col) wins */
    size_t br = 0, bc = 0;
    float best = s[0]; /* scores[0] == data[0][0] (scores.rs:246-254) */
    for (size_t i = 0; i < rows; i++)
      for (size_t j = 0; j < c; j++) {
        float x = s[i * stride + j];
        if (x >= best) {
          br = i;
          bc = j;
          best = x;
        }
      }
    *row = br;
    *col = bc;
    return 0;
  }
  /* per-column pass shared by AVX2 (avx2.rs:373-403) and SSE2: running best
   * starts at row 0, replaced when best <= x  => last row wins ties */
  size_t *p = (size_t *)calloc(c, sizeof(size_t));
  float *b = (float *)malloc(c * sizeof(float));
  for (size_t j = 0; j < c; j++) b[j] = s[j];
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < c; j++) {
      float x = s[i * stride + j];
      if (b[j] <= x) {
        p[j] = i;
        b[j] = x;
      }
    }
  size_t br = 0, bc = 0;
  float best = s[0]; /* avx2.rs:411-412: best_pos = (0, 0) */
  for (size_t j = 0; j < c; j++) {
    float x = s[p[j] * stride + j];
    int take = (rule == LMO_RULE_AVX2) ? (x > best)   /* avx2.rs:417 */
                                       : (x >= best); /* sse2.rs:252 */
    if (take) {
      best = x;
      br = p[j];
      bc = j;
    }
  }
  free(p);
  free(b);
  *row = br;
  *col = bc;
  return 0;
}

int lmo_argmax_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
                  int rule, size_t *row, size_t *col) {
  if (rows == 0) return -1;
  if (rule == LMO_RULE_GENERIC || rule == LMO_RULE_SSE2) {
    /* generic default (SSE2 has no u8 argmax, pli/mod.rs:376-381) */
    size_t br = 0, bc = 0;
    uint8_t best = s[0];
    for (size_t i = 0; i < rows; i++)
      for (size_t j = 0; j < c; j++) {
        uint8_t x = s[i * stride + j];
        if (x >= best) {
          br = i;
          bc = j;
          best = x;
        }
      }
    *row = br;
    *col = bc;
    return 0;
  }
  /* avx2.rs:484-529: per column `r > s - 1` i.e. r >= s => last row wins;
   * across columns Iterator::max_by_key => last column among equal maxima */
  size_t br = 0, bc = 0;
  int have = 0;
  uint8_t best = 0;
  for (size_t j = 0; j < c; j++) {
    size_t pj = 0;
    int sj = -1; /* avx2.rs:492: s = -1; after a hit s = r - 1 (avx2.rs:515) */
    for (size_t i = 0; i < rows; i++) {
      int x = s[i * stride + j];
      if (x > sj) { /* avx2.rs:507 cmpgt_epi16(r, s) */
        pj = i;
        sj = x - 1;
      }
    }
    uint8_t x = s[pj * stride + j];
    if (!have || x >= best) {
      have = 1;
      best = x;
      br = pj;
      bc = j;
    }
  }
  *row = br;
  *col = bc;
  return 0;
}

int lmo_max_f32(const float *s, size_t rows, size_t c, size_t stride, int rule,
                float *out) {
  if (rows == 0) return -1;
  if (rule == LMO_RULE_AVX2) {
    /* avx2.rs:438-464: accumulators start at 0.0 => max(0, true max) */
    float m = 0.0f;
    for (size_t i = 0; i < rows; i++)
      for (size_t j = 0; j < c; j++) {
        float x = s[i * stride + j];
        if (x > m) m = x;
      }
    *out = m;
    return 0;
  }
  size_t r, col;
  lmo_argmax_f32(s, rows, c, stride, rule, &r, &col); /* pli/mod.rs:158-160 */
  *out = s[r * stride + col];
  return 0;
}

int lmo_max_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
               uint8_t *out) {
  if (rows == 0) return -1;
  uint8_t m = 0;
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < c; j++)
      if (s[i * stride + j] > m) m = s[i * stride + j];
  *out = m;
  return 0;
}

/* ---------------------------------------------------------------- threshold */

size_t lmo_threshold_f32(const float *s, size_t rows, size_t c, size_t stride,
                         float t, uint64_t *rows_out, uint32_t *cols_out,
                         size_t cap) {
  size_t n = 0;
  for (size_t i = 0; i < rows; i++) /* pli/mod.rs:212-219 */
    for (size_t j = 0; j < c; j++)
      if (s[i * stride + j] >= t) {
        if (n < cap) {
          rows_out[n] = i;
          cols_out[n] = (uint32_t)j;
        }
        n++;
      }
  return n;
}

size_t lmo_threshold_u8(const uint8_t *s, size_t rows, size_t c, size_t stride,
                        uint8_t t, uint64_t *rows_out, uint32_t *cols_out,
                        size_t cap) {
  size_t n = 0;
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < c; j++)
      if (s[i * stride + j] >= t) {
        if (n < cap) {
          rows_out[n] = i;
          cols_out[n] = (uint32_t)j;
        }
        n++;
      }
  return n;
}

/* ---------------------------------------------------------- PSSM building */

void lmo_counts_from_sequences(int alphabet, const uint8_t *const *seqs,
                               size_t nseq, size_t m, uint32_t *counts) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  memset(counts, 0, m * k * sizeof(uint32_t));
  for (size_t q = 0; q < nseq; q++) /* pwm/mod.rs:228-230 */
    for (size_t i = 0; i < m; i++) counts[i * k + seqs[q][i]] += 1;
}

void lmo_to_freq(int alphabet, const uint32_t *counts, size_t m, float pseudo,
                 float *freq) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  for (size_t i = 0; i < m; i++) { /* pwm/mod.rs:246-256 */
    for (size_t j = 0; j < k; j++) {
      /* Pseudocounts::from(f32): default symbol gets 0 (abc.rs:558-574) */
      float p = (j == k - 1) ? 0.0f : pseudo;
      freq[i * k + j] = (float)counts[i * k + j] + p;
    }
    volatile float s = 0.0f; /* dst.iter().sum() */
    for (size_t j = 0; j < k; j++) s = s + freq[i * k + j];
    for (size_t j = 0; j < k; j++) freq[i * k + j] = freq[i * k + j] / s;
  }
}

void lmo_to_scoring(int alphabet, const float *freq, size_t m, const float *bg,
                    float *pssm) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  float uni[32];
  if (!bg) { /* Background::uniform, abc.rs:473-487 */
    for (size_t j = 0; j < k; j++)
      uni[j] = (j == k - 1) ? 0.0f : 1.0f / (float)(k - 1);
    bg = uni;
  }
  for (size_t i = 0; i < m; i++) /* pwm/mod.rs:420-428 */
    for (size_t j = 0; j < k; j++) {
      float f = bg[j];
      if (f == 0.0f)
        pssm[i * k + j] = -INFINITY;
      else
        pssm[i * k + j] = log2f(freq[i * k + j] / f);
    }
}

float lmo_min_score(int alphabet, const float *pssm, size_t m) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  volatile float sum = 0.0f;
  for (size_t i = 0; i < m; i++) { /* pwm/mod.rs:592-602 */
    float mn = pssm[i * k];
    for (size_t j = 1; j + 1 < k; j++)
      if (pssm[i * k + j] < mn) mn = pssm[i * k + j];
    sum = sum + mn;
  }
  return sum;
}

float lmo_max_score(int alphabet, const float *pssm, size_t m) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  volatile float sum = 0.0f;
  for (size_t i = 0; i < m; i++) { /* pwm/mod.rs:605-615 */
    float mx = pssm[i * k];
    for (size_t j = 1; j + 1 < k; j++)
      if (pssm[i * k + j] > mx) mx = pssm[i * k + j];
    sum = sum + mx;
  }
  return sum;
}

/* Rust `as u8` on f32: saturating, NaN -> 0 */
static uint8_t f32_as_u8(float x) {
  if (!(x == x)) return 0;
  if (x <= 0.0f) return 0;
  if (x >= 255.0f) return 255;
  return (uint8_t)x;
}

void lmo_to_discrete(int alphabet, const float *pssm, size_t m, uint8_t *dm,
                     float *factor, float *offsets, float *offset) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  float max_score = lmo_max_score(alphabet, pssm, m); /* pwm/mod.rs:666 */
  volatile float off = 0.0f;
  for (size_t i = 0; i < m; i++) { /* pwm/mod.rs:667-681 */
    float mn = 0.0f;
    for (size_t j = 0; j + 1 < k; j++) {
      float x = pssm[i * k + j];
      if (isinf(x)) x = -max_score;
      if (j == 0 || x < mn) mn = x;
    }
    offsets[i] = mn;
    off = off + mn;
  }
  float fac = (max_score - off) / 255.0f; /* pwm/mod.rs:683 */
  for (size_t i = 0; i < m; i++) /* pwm/mod.rs:686-690 */
    for (size_t j = 0; j < k; j++)
      dm[i * k + j] = f32_as_u8(ceilf((pssm[i * k + j] - offsets[i]) / fac));
  *factor = fac;
  *offset = off;
}

uint8_t lmo_scale(float score, float factor, float offset) {
  return f32_as_u8(floorf((score - offset) / factor)); /* pwm/mod.rs:782-784 */
}

float lmo_unscale(uint8_t score, float factor, float offset) {
  return (float)score * factor + offset; /* pwm/mod.rs:788-790 */
}

float lmo_score_position(const float *pssm, size_t m, size_t k,
                         const uint8_t *seq, size_t seq_stride, size_t rows,
                         size_t pos) {
  volatile float score = 0.0f; /* pwm/mod.rs:656-661 */
  for (size_t j = 0; j < m; j++) {
    size_t idx = pos + j; /* seq.rs:436-440: col = idx / rows, row = idx % rows */
    uint8_t s = seq[(idx % rows) * seq_stride + idx / rows];
    score = score + pssm[j * k + s];
  }
  return score;
}

/* ------------------------------------------------------------------ Scanner */

typedef struct {
  uint64_t pos;
  float score;
} hit_t;

static int hit_cmp(const void *a, const void *b) {
  const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
  return (x->pos > y->pos) - (x->pos < y->pos);
}

size_t lmo_scanner_collect(int alphabet, const float *pssm, size_t m,
                           const uint8_t *seq, size_t seq_stride, size_t c,
                           size_t rows, size_t wrap, size_t len,
                           float threshold, size_t block_size,
                           uint64_t *pos_out, float *score_out, size_t cap) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  uint8_t *dm = (uint8_t *)malloc(m * k);
  float *offsets = (float *)malloc(m * sizeof(float));
  float factor, offset;
  lmo_to_discrete(alphabet, pssm, m, dm, &factor, offsets, &offset); /* scan.rs:129 */
  uint8_t t = lmo_scale(threshold, factor, offset);                  /* scan.rs:171 */
  uint8_t *ds = (uint8_t *)malloc(block_size * c);
  size_t nh = 0, hcap = 1024;
  hit_t *hits = (hit_t *)malloc(hcap * sizeof(hit_t));
  size_t total_rows = rows + wrap; /* seq.matrix().rows() */
  for (size_t row = 0; row < total_rows; row += block_size) { /* scan.rs:172 */
    size_t end = row + block_size < rows ? row + block_size : rows; /* :174-175 */
    size_t got = lmo_score_u8_rows(dm, m, k, seq, seq_stride, c, len, row,
                                   end > row ? end : row, ds, c, 1);
    if (got == 0) continue; /* empty scores: max() is None -> 0; threshold() is empty */
    uint8_t mx;
    lmo_max_u8(ds, got, c, c, &mx);
    if (mx >= t) { /* scan.rs:181 */
      for (size_t i = 0; i < got; i++)
        for (size_t j = 0; j < c; j++)
          if (ds[i * c + j] >= t) { /* pli/mod.rs:215 */
            size_t index = j * rows + row + i; /* scan.rs:185 */
            if (index + m <= len) {
              float sc = lmo_score_position(pssm, m, k, seq, seq_stride, rows, index);
              if (sc >= threshold) {
                if (nh == hcap) {
                  hcap *= 2;
                  hits = (hit_t *)realloc(hits, hcap * sizeof(hit_t));
                }
                hits[nh].pos = index;
                hits[nh].score = sc;
                nh++;
              }
            }
          }
    }
  }
  qsort(hits, nh, sizeof(hit_t), hit_cmp);
  for (size_t i = 0; i < nh && i < cap; i++) {
    pos_out[i] = hits[i].pos;
    score_out[i] = hits[i].score;
  }
  free(hits);
  free(ds);
  free(offsets);
  free(dm);
  return nh;
}

int lmo_scanner_max(int alphabet, const float *pssm, size_t m,
                    const uint8_t *seq, size_t seq_stride, size_t c,
                    size_t rows, size_t wrap, size_t len, float threshold,
                    size_t block_size, uint64_t *pos, float *score) {
  size_t k = (size_t)lmo_alphabet_k(alphabet);
  uint8_t *dm = (uint8_t *)malloc(m * k);
  float *offsets = (float *)malloc(m * sizeof(float));
  float factor, offset;
  lmo_to_discrete(alphabet, pssm, m, dm, &factor, offsets, &offset);
  uint8_t best_discrete = lmo_scale(threshold, factor, offset); /* scan.rs:211-214 */
  int have = 0;
  uint64_t bpos = 0;
  float bscore = 0.0f;
  uint8_t *ds = (uint8_t *)malloc(block_size * c);
  size_t total_rows = rows + wrap;
  for (size_t row = 0; row < total_rows; row += block_size) { /* scan.rs:220 */
    size_t end = row + block_size < rows ? row + block_size : rows;
    size_t got = lmo_score_u8_rows(dm, m, k, seq, seq_stride, c, len, row,
                                   end > row ? end : row, ds, c, 1);
    if (got == 0) continue;
    uint8_t mx;
    lmo_max_u8(ds, got, c, c, &mx);
    if (mx >= best_discrete) { /* scan.rs:227 */
      for (size_t i = 0; i < got; i++)
        for (size_t j = 0; j < c; j++) {
          uint8_t d = ds[i * c + j];
          if (d >= best_discrete) { /* scan.rs:230-232 (re-checked: it moves) */
            size_t index = j * rows + row + i;
            /* NB scan.rs:234 does not bound-check index + m <= len: cells
             * past the end read N padding / wrap rows. score_position indexes
             * the striped matrix with idx % rows, idx / rows which is only in
             * range while idx < rows * c; guard like the Vec bounds check
             * would (a panic in the reference) by skipping. */
            if (index + m > rows * c) continue;
            float sc = lmo_score_position(pssm, m, k, seq, seq_stride, rows, index);
            if (have) {
              if ((sc > bscore) | (sc == bscore && index > bpos)) { /* scan.rs:235 */
                bpos = index;
                bscore = sc;
                best_discrete = d;
              }
            } else {
              have = 1;
              bpos = index;
              bscore = sc;
            }
          }
        }
    }
  }
  free(ds);
  free(offsets);
  free(dm);
  if (!have) return -1;
  *pos = bpos;
  *score = bscore;
  return 0;
}

/* ---------------------------------------------------------------- synthetic */

static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void lmo_synth_dna(uint64_t seed, uint64_t offset, size_t n, uint8_t *dst) {
  for (size_t q = 0; q < n; q++) {
    uint64_t i = offset + q;
    uint64_t w = splitmix64(seed + (i >> 5));
    dst[q] = (uint8_t)((w >> (2 * (i & 31))) & 3);
  }
}

void lmo_synth_protein(uint64_t seed, uint64_t offset, size_t n, uint8_t *dst) {
  for (size_t q = 0; q < n; q++) {
    uint64_t i = offset + q;
    dst[q] = (uint8_t)((splitmix64(seed + i) >> 11) % 20);
  }
}
