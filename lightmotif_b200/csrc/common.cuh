// common.cuh -- shared device/host helpers for the B200 (sm_100a) kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lightmotif_b200.h"

namespace lm {

constexpr int LANES = LM_LANES;  // striped columns == warp width

// alphabet constants (lightmotif/src/abc.rs:88-110,190-228)
__host__ __device__ inline int alphabet_k(int alphabet) { return alphabet == LM_DNA ? 5 : 21; }
inline const char* alphabet_str(int alphabet) {
  return alphabet == LM_DNA ? "ACTGN" : "ACDEFGHIKLMNPQRSTVWYX";
}

// ---- order-preserving float <-> u32 key (for atomicMax on packed keys) ----
__host__ __device__ inline uint32_t f32_to_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } v; v.f = f; uint32_t b = v.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float key_to_f32(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } v; v.u = b; return v.f;
#endif
}

// ---- coupling of one rank's fused scan to the other ranks of a row-sharded scan (comm.cu) -------
// Every rank owns an exchange buffer in its HBM that all ranks map (peer access within a process,
// cudaIpc between processes).  During a scan the ranks share their running maximum through a
// tagged word per launch slot ("floor": epoch << 32 | score key, raised with system-scope atomicMax
// over NVLink), so every GPU filters at the GLOBAL maximum found so far; when a rank's scan ends,
// the last warp of its scan kernel writes the rank's result block into every rank's buffer and
// raises an arrival flag there (release at system scope).  No collective kernel is involved.
constexpr int XBLOCK_WORDS = 40;  // u64 words of a published result block (320 bytes)
// words: [0,32) per-column winner keys, [32] first cell (f32 bits), [33] hit count (threshold scans),
//        [34] first global row of the rank's slab, [35] rows of the slab, [36] epoch, [37] candidates,
//        [38] overflow tiles
struct PeerLink {
  int world = 0, rank = 0;
  uint32_t epoch = 0;                                  // same on all ranks for one sharded scan
  unsigned long long* floor_local = nullptr;           // this rank's floor word of the slot
  unsigned long long* const* floor_peers = nullptr;    // device array [world]: that word on every rank
  unsigned long long* const* gather_peers = nullptr;   // device array [world]: where THIS rank's block goes on every rank
  unsigned long long* const* flag_peers = nullptr;     // device array [world]: the arrival flag to raise there
  uint64_t row_offset = 0, n_rows = 0;
};

#ifdef __CUDACC__
// score key of the shared floor if it belongs to this scan, else 0
__device__ __forceinline__ uint32_t peer_floor_key(const PeerLink& l) {
  const unsigned long long v = __ldcg(l.floor_local);
  return (uint32_t)(v >> 32) == l.epoch ? (uint32_t)v : 0u;
}
// raise the floor on every rank (called by a whole warp; lane p serves rank p, p + 32, ...)
__device__ __forceinline__ void peer_push_floor(const PeerLink& l, uint32_t key, int lane) {
  const unsigned long long v = ((unsigned long long)l.epoch << 32) | key;
  for (int p = lane; p < l.world; p += 32) atomicMax_system(l.floor_peers[p], v);
}
// publish this rank's result block (`small` = SmallBlock in local HBM) to every rank; one warp.
// All earlier writes to `small` by this grid must be visible to the calling warp.
__device__ __forceinline__ void peer_publish(const PeerLink& l, const unsigned long long* small, int lane) {
  unsigned long long w0 = __ldcg(&small[lane]);   // keys
  unsigned long long w1 = 0;                      // words 32..39
  if (lane == 0) w1 = __ldcg(&small[32]) & 0xffffffffull;  // first cell
  else if (lane == 1) w1 = __ldcg(&small[33]);             // hit count
  else if (lane == 2) w1 = l.row_offset;
  else if (lane == 3) w1 = l.n_rows;
  else if (lane == 4) w1 = l.epoch;
  else if (lane == 5) w1 = __ldcg(&small[34]);             // candidates
  else if (lane == 6) w1 = __ldcg(&small[35]);             // overflow tiles
  for (int p = 0; p < l.world; ++p) {
    volatile unsigned long long* dst = l.gather_peers[p];
    dst[lane] = w0;
    if (lane < XBLOCK_WORDS - 32) dst[32 + lane] = w1;
  }
  __threadfence_system();
  __syncwarp();
  for (int p = lane; p < l.world; p += 32) {
    unsigned long long* f = l.flag_peers[p];
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)l.epoch) : "memory");
  }
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// wait with back-off: between polls the warp really sleeps (nanosleep), so a waiting warp costs a
// handful of issue slots per microsecond instead of spinning through the retry loop (the hinted form
// of try_wait compiles to a TRYWAIT + NANOSLEEP.SYNCS loop that wakes on every barrier event of the
// CTA: profiles/README.md, r01 v5)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// wait with a suspend-time hint: the warp is parked by the hardware for up to `hint_ns` per attempt
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITH_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONEH_%=;\n"
      "bra WAITH_%=;\n"
      "DONEH_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(hint_ns)
      : "memory");
}
// L2 prefetch of a contiguous global range (bytes % 16 == 0, 16-byte aligned): no shared-memory
// destination, so the bytes in flight are not limited by the tile ring
__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// make generic-proxy smem accesses visible/ordered w.r.t. the async proxy
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace lm
