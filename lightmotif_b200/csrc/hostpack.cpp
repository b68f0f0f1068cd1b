// hostpack.cpp -- host side of the packed upload (upload.cu): DNA symbols 0..3 (abc.rs:231-256:
// A, C, T, G; N = 4 does not fit) packed 4 per byte, little end first:
//   out[k] = s[4k] | s[4k+1] << 2 | s[4k+2] << 4 | s[4k+3] << 6
// Plain host code (no CUDA): this is transfer compression, not scoring -- the bytes are expanded
// again on the device by unpack2_kernel before anything looks at them.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace lmi {

namespace {

// 8 symbols -> 2 bytes (SWAR); `bad` collects every bit that no symbol 0..3 has
inline uint16_t pack8(uint64_t x, uint64_t& bad) {
  bad |= x;
  uint64_t y = (x & 0x0003000300030003ull) | ((x >> 6) & 0x000C000C000C000Cull);  // 16-bit units: 2 symbols
  uint64_t z = (y | (y >> 12)) & 0x000000FF000000FFull;                            // 32-bit units: 4 symbols
  return (uint16_t)(z | (z >> 24));
}

bool pack_scalar(const uint8_t* src, size_t n, uint8_t* dst) {
  uint64_t bad = 0;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t x;
    std::memcpy(&x, src + i, 8);
    const uint16_t p = pack8(x, bad);
    std::memcpy(dst + i / 4, &p, 2);
  }
  if (i < n) {  // (n is a multiple of 4 for every caller; a tail of 4 symbols)
    uint64_t x = 0;
    std::memcpy(&x, src + i, n - i);
    const uint16_t p = pack8(x, bad);
    std::memcpy(dst + i / 4, &p, (n - i + 3) / 4);
  }
  return (bad & 0xFCFCFCFCFCFCFCFCull) == 0;
}

#if defined(__x86_64__)
// 128 symbols per iteration: (s0 + 4 s1) by maddubs, (.. + 16 (s2 + 4 s3)) by madd, then 32 -> 8 bit packs
__attribute__((target("avx2"))) bool pack_avx2(const uint8_t* src, size_t n, uint8_t* dst) {
  const __m256i c1 = _mm256_set1_epi16(0x0401);
  const __m256i c2 = _mm256_set1_epi32(0x00100001);
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  __m256i bad = _mm256_setzero_si256();
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    // one thread streaming from DRAM is bound by its outstanding misses: asking for the lines 4 KB ahead
    // raised the rate by a third (a prefetch never faults, so running past the end is harmless)
    _mm_prefetch((const char*)(src + i + 4096), _MM_HINT_T0);
    _mm_prefetch((const char*)(src + i + 4096 + 64), _MM_HINT_T0);
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64));
    const __m256i d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
    const __m256i pa = _mm256_madd_epi16(_mm256_maddubs_epi16(a, c1), c2);
    const __m256i pb = _mm256_madd_epi16(_mm256_maddubs_epi16(b, c1), c2);
    const __m256i pc = _mm256_madd_epi16(_mm256_maddubs_epi16(c, c1), c2);
    const __m256i pd = _mm256_madd_epi16(_mm256_maddubs_epi16(d, c1), c2);
    const __m256i w = _mm256_packus_epi16(_mm256_packus_epi32(pa, pb), _mm256_packus_epi32(pc, pd));
    _mm256_storeu_si256((__m256i*)(dst + i / 4), _mm256_permutevar8x32_epi32(w, order));
  }
  bool ok = _mm256_testz_si256(bad, _mm256_set1_epi8((char)0xFC)) != 0;
  if (i < n) ok = pack_scalar(src + i, n - i, dst + i / 4) && ok;
  return ok;
}
#endif

}  // namespace

// Pack `n` symbols (n % 4 == 0) into n / 4 bytes.  Returns false when a symbol > 3 was met (the
// output is then meaningless: the caller sends that chunk as it is).
bool pack2(const uint8_t* src, size_t n, uint8_t* dst) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return pack_avx2(src, n, dst);
#endif
  return pack_scalar(src, n, dst);
}

bool pack2_scalar(const uint8_t* src, size_t n, uint8_t* dst) { return pack_scalar(src, n, dst); }

}  // namespace lmi
