// comm.cu -- row-sharded scans across the GPUs of one box (include/lightmotif_b200.h, lm_comm_*).
//
// The reference has no multi-device path; its natural sharding axis is the row range of
// Score::score_rows_into (lightmotif/src/pli/mod.rs:72-77): rank g scans rows [r_g, r_{g+1}) of the
// global striped matrix (+ M - 1 halo rows), a contiguous byte range.  The scan itself needs no
// data exchange.  What the GPUs do exchange, entirely through peer-mapped HBM over NVLink / NVSwitch
// and without any collective kernel:
//
//   * DURING the scan, the running maximum: every rank owns a tagged "floor" word per launch slot;
//     a rank whose own maximum rises pushes it to all ranks with a system-scope atomicMax, and every
//     producer warp of the filter kernel reads its local floor with each tile it claims.  All GPUs
//     therefore filter at the GLOBAL maximum found so far (fewer exact rechecks, and the slowest
//     rank is no slower than the luckiest one) -- prune.cu.
//   * AT THE END of the scan, the result block: the last warp of the scan kernel to finish writes
//     the rank's 320-byte block (32 per-column winners, first cell, hit count, slab geometry) into
//     every rank's gather area and raises an arrival flag there (st.release.sys) -- common.cuh
//     peer_publish.  A one-warp kernel on a side stream waits for the `world` flags and the blocks
//     are copied to pinned host memory, where the tie rule is applied (lm_argmax_merge semantics:
//     later slabs win row ties, then the chosen back-end's rule).
//   * threshold scans additionally gather the hit lists: every rank reads the other ranks' lists
//     straight out of their exchange buffers (peer copies), acknowledges, and sorts.
//
// Two ways to form a communicator: lm_comm_init (ONE process drives N GPUs: what a Rust caller of
// the C ABI uses) and lm_comm_init_rank + lm_comm_connect (one process per GPU, e.g. under torchrun;
// the buffers are shared with cudaIpc handles that the caller exchanges by any means).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include <mutex>

#include "internal.cuh"

using namespace lm;
using namespace lmi;

namespace {

constexpr int MAXR = LM_MAX_RANKS;
constexpr int XSLOTS = LM_SLOTS + 1;                 // launch slots + one for the synchronous threshold scans
constexpr int THR_SLOT = LM_SLOTS;
constexpr size_t XBLOCK_BYTES = XBLOCK_WORDS * 8;   // 320
// exchange buffer layout (bytes)
constexpr size_t XOFF_FLOOR = 0;                                        // [XSLOTS] x 128 B
constexpr size_t XOFF_FLAGS = XOFF_FLOOR + XSLOTS * 128;                // [XSLOTS][2][MAXR] u64 arrival flags
constexpr size_t XOFF_ACKS = XOFF_FLAGS + XSLOTS * 2 * MAXR * 8;        // [MAXR] u64 "hits consumed" flags
constexpr size_t XOFF_GATHER = (XOFF_ACKS + MAXR * 8 + 127) & ~(size_t)127;  // [XSLOTS][2][MAXR] blocks
constexpr size_t XOFF_STATUS = XOFF_GATHER + XSLOTS * 2 * MAXR * XBLOCK_BYTES;  // [XSLOTS] u64 wait status
// local-only pointer tables (what the scan kernel dereferences)
constexpr size_t XOFF_PTR_FLOOR = (XOFF_STATUS + XSLOTS * 8 + 127) & ~(size_t)127;  // [XSLOTS][MAXR]
constexpr size_t XOFF_PTR_GATHER = XOFF_PTR_FLOOR + XSLOTS * MAXR * 8;              // [XSLOTS][2][MAXR]
constexpr size_t XOFF_PTR_FLAG = XOFF_PTR_GATHER + XSLOTS * 2 * MAXR * 8;           // [XSLOTS][2][MAXR]
constexpr size_t XOFF_PTR_ACK = XOFF_PTR_FLAG + XSLOTS * 2 * MAXR * 8;              // [MAXR]
constexpr size_t XOFF_HITS = (XOFF_PTR_ACK + MAXR * 8 + 4095) & ~(size_t)4095;      // lm_hit[hcap]

constexpr unsigned long long WAIT_TIMEOUT_NS = 30ull * 1000 * 1000 * 1000;

__global__ void publish_kernel(const PeerLink link, const unsigned long long* small) {
  peer_publish(link, small, threadIdx.x);
}

// one warp: lane r waits until rank r's flag of this (slot, parity) carries `epoch`
__global__ void gather_wait_kernel(const unsigned long long* flags, int world, unsigned long long epoch,
                                   unsigned long long* status) {
  const int lane = threadIdx.x;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  bool ok = true;
  if (lane < world) {
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + lane) : "memory");
      if (v == epoch) break;
      __nanosleep(200);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > WAIT_TIMEOUT_NS) { ok = false; break; }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) *status = ok ? epoch : ~0ull;
}

// one warp: tell every rank that this rank has copied their hit lists of `epoch`
__global__ void ack_kernel(unsigned long long* const* ack_peers, int world, unsigned long long epoch) {
  __threadfence_system();
  for (int p = threadIdx.x; p < world; p += 32)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ack_peers[p]), "l"(epoch) : "memory");
}

// one warp: wait until every rank has acknowledged `epoch` (0 = nothing to wait for)
__global__ void ack_wait_kernel(const unsigned long long* acks, int world, unsigned long long epoch,
                                unsigned long long* status) {
  const int lane = threadIdx.x;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  bool ok = true;
  if (lane < world) {
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(acks + lane) : "memory");
      if (v >= epoch) break;
      __nanosleep(200);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > WAIT_TIMEOUT_NS) { ok = false; break; }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0 && !ok) *status = ~0ull;
}

}  // namespace

namespace lm {
cudaError_t launch_publish(const PeerLink& link, const unsigned long long* small, cudaStream_t stream) {
  publish_kernel<<<1, 32, 0, stream>>>(link, small);
  return cudaGetLastError();
}
}  // namespace lm

struct lm_comm {
  int world = 0, n_local = 0;
  bool multiproc = false, connected = false;
  size_t hcap = 0;
  struct Local {
    int rank = 0;
    lm_ctx* ctx = nullptr;
    bool own_ctx = false;
    uint8_t* xbuf = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t gathered[XSLOTS] = {};
    cudaEvent_t scan_done = nullptr;
    uint64_t* h_gather[XSLOTS] = {};  // pinned: [world][XBLOCK_WORDS] + status word
    DevBuf d_stage;                   // gathered hit lists
  } local[MAXR];
  uint8_t* peer_xbuf[MAXR] = {};      // every rank's exchange buffer as mapped in this process
  bool ipc_opened[MAXR] = {};
  uint32_t epoch = 0;
  uint32_t slot_uses[XSLOTS] = {};
  uint32_t slot_epoch[XSLOTS] = {};
  int slot_parity[XSLOTS] = {};
  bool slot_busy[XSLOTS] = {};
  uint32_t last_thr_epoch = 0;
  char err[256] = {0};
};

namespace {

size_t xbuf_bytes(size_t hcap) { return XOFF_HITS + hcap * sizeof(lm_hit); }

int comm_fail(lm_comm* c, cudaError_t e, const char* what) {
  snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? LM_ERR_OOM : LM_ERR_CUDA;
}
#define CM_TRY(expr)                                        \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return comm_fail(comm, _e, #expr); \
  } while (0)

// CUDA loads kernels lazily, at their first launch, and a load may synchronise with whatever runs on the
// device.  A sharded scan keeps gather_wait_kernel spinning until every rank has published; if a rank that
// shares the device (a communicator may list a device twice) then launches an instance nobody has launched
// before -- e.g. the second scan of a matrix, re-planned for the maximum the first one found --, its load
// waits for the kernel that waits for it (observed as the 30 s time-out).  Communicators therefore load every
// scan kernel up front, once per device and process.
int preload_kernels(lm_comm* comm, int device) {
  static std::mutex mu;
  static std::vector<int> done;
  std::lock_guard<std::mutex> lock(mu);
  if (std::find(done.begin(), done.end(), device) != done.end()) return LM_OK;
  CM_TRY(cudaSetDevice(device));
  CM_TRY(lm::preload_prune_kernels());
  CM_TRY(lm::preload_score_kernels());
  cudaFuncAttributes at;
  CM_TRY(cudaFuncGetAttributes(&at, publish_kernel));
  CM_TRY(cudaFuncGetAttributes(&at, gather_wait_kernel));
  CM_TRY(cudaFuncGetAttributes(&at, ack_kernel));
  CM_TRY(cudaFuncGetAttributes(&at, ack_wait_kernel));
  done.push_back(device);
  return LM_OK;
}

int alloc_local(lm_comm* comm, int li) {
  lm_comm::Local& L = comm->local[li];
  CM_TRY(cudaSetDevice(L.ctx->device));
  LM_CHECK(preload_kernels(comm, L.ctx->device));
  const size_t bytes = xbuf_bytes(comm->hcap);
  CM_TRY(cudaMalloc((void**)&L.xbuf, bytes));  // plain cudaMalloc: exportable with cudaIpcGetMemHandle
  CM_TRY(cudaMemset(L.xbuf, 0, XOFF_HITS));
  CM_TRY(cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking));
  CM_TRY(cudaEventCreateWithFlags(&L.scan_done, cudaEventDisableTiming));
  for (int s = 0; s < XSLOTS; ++s) {
    CM_TRY(cudaEventCreateWithFlags(&L.gathered[s], cudaEventDisableTiming));
    CM_TRY(cudaHostAlloc((void**)&L.h_gather[s], (size_t)MAXR * XBLOCK_BYTES + 64, cudaHostAllocDefault));
  }
  comm->peer_xbuf[L.rank] = L.xbuf;
  return LM_OK;
}

// write the pointer tables the scan kernels of local rank `li` dereference
int wire_local(lm_comm* comm, int li) {
  lm_comm::Local& L = comm->local[li];
  std::vector<uint64_t> pf(XSLOTS * MAXR, 0), pg(XSLOTS * 2 * MAXR, 0), pl(XSLOTS * 2 * MAXR, 0), pa(MAXR, 0);
  for (int r = 0; r < comm->world; ++r) {
    const uint8_t* xb = comm->peer_xbuf[r];
    pa[r] = (uint64_t)(xb + XOFF_ACKS + (size_t)L.rank * 8);
    for (int s = 0; s < XSLOTS; ++s) {
      pf[s * MAXR + r] = (uint64_t)(xb + XOFF_FLOOR + (size_t)s * 128);
      for (int p = 0; p < 2; ++p) {
        const size_t idx = ((size_t)(s * 2 + p) * MAXR + L.rank);  // where THIS rank's block / flag lives on rank r
        pg[(s * 2 + p) * MAXR + r] = (uint64_t)(xb + XOFF_GATHER + idx * XBLOCK_BYTES);
        pl[(s * 2 + p) * MAXR + r] = (uint64_t)(xb + XOFF_FLAGS + idx * 8);
      }
    }
  }
  CM_TRY(cudaSetDevice(L.ctx->device));
  CM_TRY(cudaMemcpy(L.xbuf + XOFF_PTR_FLOOR, pf.data(), pf.size() * 8, cudaMemcpyHostToDevice));
  CM_TRY(cudaMemcpy(L.xbuf + XOFF_PTR_GATHER, pg.data(), pg.size() * 8, cudaMemcpyHostToDevice));
  CM_TRY(cudaMemcpy(L.xbuf + XOFF_PTR_FLAG, pl.data(), pl.size() * 8, cudaMemcpyHostToDevice));
  CM_TRY(cudaMemcpy(L.xbuf + XOFF_PTR_ACK, pa.data(), pa.size() * 8, cudaMemcpyHostToDevice));
  CM_TRY(cudaDeviceSynchronize());
  return LM_OK;
}

PeerLink make_link(const lm_comm* comm, int li, int slot, int parity, uint32_t epoch, const lm_seq* slab) {
  const lm_comm::Local& L = comm->local[li];
  PeerLink k;
  k.world = comm->world;
  k.rank = L.rank;
  k.epoch = epoch;
  k.floor_local = (unsigned long long*)(L.xbuf + XOFF_FLOOR + (size_t)slot * 128);
  k.floor_peers = (unsigned long long* const*)(L.xbuf + XOFF_PTR_FLOOR + (size_t)slot * MAXR * 8);
  k.gather_peers = (unsigned long long* const*)(L.xbuf + XOFF_PTR_GATHER + (size_t)(slot * 2 + parity) * MAXR * 8);
  k.flag_peers = (unsigned long long* const*)(L.xbuf + XOFF_PTR_FLAG + (size_t)(slot * 2 + parity) * MAXR * 8);
  k.row_offset = slab ? slab->row_offset : 0;
  k.n_rows = slab ? slab->rows : 0;
  return k;
}

// side stream of local rank li: wait for this rank's scan, then for every rank's block; copy out
int enqueue_gather(lm_comm* comm, int li, int slot, int parity, uint32_t epoch) {
  lm_comm::Local& L = comm->local[li];
  CM_TRY(cudaSetDevice(L.ctx->device));
  CM_TRY(cudaEventRecord(L.scan_done, L.ctx->stream));
  CM_TRY(cudaStreamWaitEvent(L.side, L.scan_done, 0));
  const size_t base = (size_t)(slot * 2 + parity) * MAXR;
  unsigned long long* status = (unsigned long long*)(L.xbuf + XOFF_STATUS + (size_t)slot * 8);
  gather_wait_kernel<<<1, 32, 0, L.side>>>((const unsigned long long*)(L.xbuf + XOFF_FLAGS + base * 8), comm->world,
                                            epoch, status);
  CM_TRY(cudaGetLastError());
  L.ctx->launches += 1;
  CM_TRY(cudaMemcpyAsync(L.h_gather[slot], L.xbuf + XOFF_GATHER + base * XBLOCK_BYTES,
                         (size_t)comm->world * XBLOCK_BYTES, cudaMemcpyDeviceToHost, L.side));
  CM_TRY(cudaMemcpyAsync((uint8_t*)L.h_gather[slot] + (size_t)MAXR * XBLOCK_BYTES, status, 8, cudaMemcpyDeviceToHost,
                         L.side));
  CM_TRY(cudaEventRecord(L.gathered[slot], L.side));
  return LM_OK;
}

int begin_collective(lm_comm* comm, int slot, uint32_t* epoch, int* parity) {
  if (!comm || !comm->connected || slot < 0 || slot >= XSLOTS) return LM_ERR_INVALID_ARG;
  if (comm->slot_busy[slot]) return LM_ERR_INVALID_ARG;
  *epoch = ++comm->epoch;
  *parity = (int)(comm->slot_uses[slot]++ & 1u);
  comm->slot_epoch[slot] = *epoch;
  comm->slot_parity[slot] = *parity;
  comm->slot_busy[slot] = true;
  return LM_OK;
}

// wait for the gathered blocks of `slot` on local rank li; returns the pinned copy
int wait_gather(lm_comm* comm, int li, int slot, const uint64_t** blocks) {
  lm_comm::Local& L = comm->local[li];
  CM_TRY(cudaSetDevice(L.ctx->device));
  CM_TRY(cudaEventSynchronize(L.gathered[slot]));
  const uint64_t status = *(const uint64_t*)((const uint8_t*)L.h_gather[slot] + (size_t)MAXR * XBLOCK_BYTES);
  if (status != comm->slot_epoch[slot]) {
    snprintf(comm->err, sizeof(comm->err), "rank %d: timed out waiting for the other ranks' result blocks", L.rank);
    return LM_ERR_CUDA;
  }
  *blocks = L.h_gather[slot];
  return LM_OK;
}

}  // namespace

extern "C" {

const char* lm_comm_last_error(const lm_comm* comm) { return comm ? comm->err : ""; }
int lm_comm_world(const lm_comm* comm) { return comm ? comm->world : 0; }
int lm_comm_n_local(const lm_comm* comm) { return comm ? comm->n_local : 0; }
lm_ctx* lm_comm_ctx(lm_comm* comm, int local_index) {
  return comm && local_index >= 0 && local_index < comm->n_local ? comm->local[local_index].ctx : nullptr;
}
int lm_comm_rank(const lm_comm* comm, int local_index) {
  return comm && local_index >= 0 && local_index < comm->n_local ? comm->local[local_index].rank : -1;
}

void lm_comm_destroy(lm_comm* comm) {
  if (!comm) return;
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    if (!L.ctx) continue;
    cudaSetDevice(L.ctx->device);
    if (L.ctx->stream) cudaStreamSynchronize(L.ctx->stream);
    if (L.side) { cudaStreamSynchronize(L.side); cudaStreamDestroy(L.side); }
    if (L.scan_done) cudaEventDestroy(L.scan_done);
    for (int s = 0; s < XSLOTS; ++s) {
      if (L.gathered[s]) cudaEventDestroy(L.gathered[s]);
      if (L.h_gather[s]) cudaFreeHost(L.h_gather[s]);
    }
    if (L.d_stage.p) cudaFree(L.d_stage.p);
  }
  if (comm->multiproc)
    for (int r = 0; r < comm->world; ++r)
      if (comm->ipc_opened[r]) cudaIpcCloseMemHandle(comm->peer_xbuf[r]);
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    if (!L.ctx) continue;
    cudaSetDevice(L.ctx->device);
    if (L.xbuf) cudaFree(L.xbuf);
    if (L.own_ctx) lm_ctx_destroy(L.ctx);
  }
  cudaGetLastError();
  delete comm;
}

int lm_comm_init(int n_dev, const int* devices, size_t hit_capacity, lm_comm** out) {
  if (!out || n_dev < 1 || n_dev > MAXR) return LM_ERR_INVALID_ARG;
  *out = nullptr;
  lm_comm* comm = new (std::nothrow) lm_comm();
  if (!comm) return LM_ERR_OOM;
  comm->world = comm->n_local = n_dev;
  comm->hcap = hit_capacity ? hit_capacity : (size_t)1 << 20;
  int st = LM_OK;
  for (int r = 0; r < n_dev && st == LM_OK; ++r) {
    lm_comm::Local& L = comm->local[r];
    L.rank = r;
    st = lm_ctx_create(devices ? devices[r] : r, &L.ctx);
    L.own_ctx = st == LM_OK;
  }
  // every device must be able to address every other device's HBM (NVLink / NVSwitch peer access)
  for (int a = 0; a < n_dev && st == LM_OK; ++a) {
    for (int b = 0; b < n_dev && st == LM_OK; ++b) {
      const int da = comm->local[a].ctx->device, db = comm->local[b].ctx->device;
      if (da == db) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, da, db) != cudaSuccess || !can) { st = LM_ERR_UNSUPPORTED; break; }
      cudaSetDevice(da);
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) st = comm_fail(comm, e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
  }
  for (int r = 0; r < n_dev && st == LM_OK; ++r) st = alloc_local(comm, r);
  for (int r = 0; r < n_dev && st == LM_OK; ++r) st = wire_local(comm, r);
  if (st != LM_OK) { lm_comm_destroy(comm); return st; }
  comm->connected = true;
  *out = comm;
  return LM_OK;
}

int lm_comm_init_rank(lm_ctx* ctx, int rank, int world, size_t hit_capacity, uint8_t handle_out[LM_COMM_HANDLE_BYTES],
                      lm_comm** out) {
  if (!ctx || !out || !handle_out || world < 1 || world > MAXR || rank < 0 || rank >= world) return LM_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) <= LM_COMM_HANDLE_BYTES, "handle size");
  *out = nullptr;
  lm_comm* comm = new (std::nothrow) lm_comm();
  if (!comm) return LM_ERR_OOM;
  comm->world = world;
  comm->n_local = 1;
  comm->multiproc = true;
  comm->hcap = hit_capacity ? hit_capacity : (size_t)1 << 20;
  comm->local[0].rank = rank;
  comm->local[0].ctx = ctx;
  int st = alloc_local(comm, 0);
  if (st == LM_OK) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, comm->local[0].xbuf);
    if (e != cudaSuccess) st = comm_fail(comm, e, "cudaIpcGetMemHandle");
    else {
      std::memset(handle_out, 0, LM_COMM_HANDLE_BYTES);
      std::memcpy(handle_out, &h, sizeof(h));
    }
  }
  if (st != LM_OK) {
    snprintf(ctx->err, sizeof(ctx->err), "%s", comm->err);
    lm_comm_destroy(comm);
    return st;
  }
  *out = comm;
  return LM_OK;
}

int lm_comm_connect(lm_comm* comm, const uint8_t* handles) {
  if (!comm || !handles || !comm->multiproc || comm->connected) return LM_ERR_INVALID_ARG;
  lm_comm::Local& L = comm->local[0];
  CM_TRY(cudaSetDevice(L.ctx->device));
  for (int r = 0; r < comm->world; ++r) {
    if (r == L.rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * LM_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    CM_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    comm->peer_xbuf[r] = (uint8_t*)p;
    comm->ipc_opened[r] = true;
  }
  LM_CHECK(wire_local(comm, 0));
  comm->connected = true;
  return LM_OK;
}

int lm_comm_score_argmax_f32_launch(lm_comm* comm, const lm_seq* const* slabs, const float* pssm, size_t m,
                                    size_t pssm_stride, int slot) {
  if (!comm || !slabs || slot >= LM_SLOTS) return LM_ERR_INVALID_ARG;
  for (int li = 0; li < comm->n_local; ++li)
    if (!slabs[li] || comm->local[li].ctx->slots[slot].state != 0) return LM_ERR_INVALID_ARG;
  uint32_t epoch;
  int parity;
  LM_CHECK(begin_collective(comm, slot, &epoch, &parity));
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    const PeerLink link = make_link(comm, li, slot, parity, epoch, slabs[li]);
    int st = argmax_launch(L.ctx, slabs[li], pssm, m, pssm_stride, 0, slabs[li]->rows, slot, &link);
    if (st < 0) {
      snprintf(comm->err, sizeof(comm->err), "rank %d: %s", L.rank, L.ctx->err);
      return st;  // the communicator is unusable after a failed collective launch
    }
    LM_CHECK(enqueue_gather(comm, li, slot, parity, epoch));
  }
  return LM_OK;
}

int lm_comm_score_argmax_f32_finish(lm_comm* comm, int slot, int rule, lm_hit* best) {
  if (!comm || !best || slot < 0 || slot >= LM_SLOTS || !comm->slot_busy[slot]) return LM_ERR_INVALID_ARG;
  comm->slot_busy[slot] = false;
  // local bookkeeping: statistics, kernel time, slot state
  for (int li = 0; li < comm->n_local; ++li) {
    lm_hit tmp;
    int st = lm_seq_score_argmax_f32_finish(comm->local[li].ctx, slot, rule, &tmp, nullptr, nullptr);
    if (st < 0) return st;
  }
  const uint64_t* blocks = nullptr;
  LM_CHECK(wait_gather(comm, 0, slot, &blocks));
  for (int li = 1; li < comm->n_local; ++li) {
    const uint64_t* other = nullptr;
    LM_CHECK(wait_gather(comm, li, slot, &other));
  }
  // merge (lm_argmax_merge): slabs in row order; later slabs hold larger rows
  std::vector<int> order(comm->world);
  for (int r = 0; r < comm->world; ++r) order[r] = r;
  std::sort(order.begin(), order.end(), [&](int a, int b) {  // by first row; ranks without rows first
    const uint64_t *x = blocks + (size_t)a * XBLOCK_WORDS, *y = blocks + (size_t)b * XBLOCK_WORDS;
    return x[34] != y[34] ? x[34] < y[34] : x[35] < y[35];
  });
  std::vector<lm_colwin> parts((size_t)comm->world * 32);
  std::vector<uint64_t> offs(comm->world + 1, 0);
  uint64_t total_rows = 0;
  for (int i = 0; i < comm->world; ++i) {
    const uint64_t* b = blocks + (size_t)order[i] * XBLOCK_WORDS;
    decode_colwins((const unsigned long long*)b, false, &parts[(size_t)i * 32]);
    offs[i] = b[34];
    total_rows = std::max<uint64_t>(total_rows, b[34] + b[35]);
  }
  offs[comm->world] = total_rows;
  float first_cell = 0.f;  // cell (0, 0) of the global matrix: first cell of the first rank that owns rows
  for (int i = 0; i < comm->world; ++i) {
    const uint64_t* b = blocks + (size_t)order[i] * XBLOCK_WORDS;
    if (b[35]) { first_cell = bits_to_score((uint32_t)b[32], false); break; }
  }
  return lm_argmax_merge(parts.data(), (size_t)comm->world, offs.data(), first_cell, total_rows, rule, best);
}

int lm_seq_score_argmax_f32_sharded(lm_comm* comm, const lm_seq* const* slabs, const float* pssm, size_t m,
                                    size_t pssm_stride, int rule, lm_hit* best) {
  int st = lm_comm_score_argmax_f32_launch(comm, slabs, pssm, m, pssm_stride, 0);
  if (st < 0) return st;
  return lm_comm_score_argmax_f32_finish(comm, 0, rule, best);
}

int lm_seq_score_threshold_f32_sharded(lm_comm* comm, const lm_seq* const* slabs, const float* pssm, size_t m,
                                       size_t pssm_stride, float t, lm_hit* hits, size_t cap, size_t* count,
                                       size_t* local_counts) {
  if (!comm || !slabs || !pssm || (cap && !hits)) return LM_ERR_INVALID_ARG;
  if (count) *count = 0;
  for (int li = 0; li < comm->n_local; ++li) {
    if (!slabs[li] || pssm_stride < (size_t)alphabet_k(slabs[li]->alphabet)) return LM_ERR_INVALID_ARG;
    if (slabs[li]->device != comm->local[li].ctx->device) return LM_ERR_INVALID_ARG;
  }
  uint32_t epoch;
  int parity;
  LM_CHECK(begin_collective(comm, THR_SLOT, &epoch, &parity));
  comm->slot_busy[THR_SLOT] = false;  // synchronous call
  const size_t small_off = (size_t)(LM_SLOTS + 1) * 1024;
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    lm_ctx* ctx = L.ctx;
    LM_CHECK(bind(ctx));
    const lm_seq* slab = slabs[li];
    const PeerLink link = make_link(comm, li, THR_SLOT, parity, epoch, slab);
    // the hit region is about to be overwritten: every rank must have copied the previous lists
    if (comm->last_thr_epoch) {
      ack_wait_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long*)(L.xbuf + XOFF_ACKS), comm->world,
                                                 comm->last_thr_epoch,
                                                 (unsigned long long*)(L.xbuf + XOFF_STATUS + (size_t)THR_SLOT * 8));
      CM_TRY(cudaGetLastError());
      ctx->launches += 1;
    }
    int st = check_score_args(m, slab->rows, slab->wrap, slab->length, 0, slab->rows);
    if (st < 0) return st;
    uint8_t* small = (uint8_t*)ctx->d_small.p + small_off;
    if (st == LM_EMPTY) {
      CM_TRY(cudaMemsetAsync(small, 0, 512, ctx->stream));
      CM_TRY(launch_publish(link, (const unsigned long long*)small, ctx->stream));
      ctx->launches += 1;
    } else {
      st = threshold_launch(ctx, slab, pssm, m, pssm_stride, 0, slab->rows, t, (lm_hit*)(L.xbuf + XOFF_HITS),
                            comm->hcap, small_off, &link, nullptr);
      if (st < 0) { snprintf(comm->err, sizeof(comm->err), "rank %d: %s", L.rank, ctx->err); return st; }
    }
    LM_CHECK(enqueue_gather(comm, li, THR_SLOT, parity, epoch));
  }
  // counts of every rank, then the lists themselves: peer copies straight out of the other ranks'
  // exchange buffers (NVLink), concatenated in rank order on local rank 0's device
  const uint64_t* blocks = nullptr;
  for (int li = comm->n_local - 1; li >= 0; --li) LM_CHECK(wait_gather(comm, li, THR_SLOT, &blocks));
  size_t total = 0;
  bool overflow = false;
  std::vector<size_t> cnt(comm->world), off(comm->world);
  for (int r = 0; r < comm->world; ++r) {
    const uint64_t c = blocks[(size_t)r * XBLOCK_WORDS + 33];
    if (local_counts) local_counts[r] = (size_t)c;
    overflow = overflow || c > comm->hcap;
    cnt[r] = (size_t)std::min<uint64_t>(c, comm->hcap);
    off[r] = total;
    total += cnt[r];
  }
  size_t grand = 0;
  for (int r = 0; r < comm->world; ++r) grand += (size_t)blocks[(size_t)r * XBLOCK_WORDS + 33];
  if (count) *count = grand;
  {
    lm_comm::Local& L = comm->local[0];
    lm_ctx* ctx = L.ctx;
    LM_CHECK(bind(ctx));
    LM_CHECK(ensure(ctx, L.d_stage, std::max<size_t>(total, 1) * sizeof(lm_hit)));
    for (int r = 0; r < comm->world; ++r)
      if (cnt[r])
        CM_TRY(cudaMemcpyAsync((lm_hit*)L.d_stage.p + off[r], comm->peer_xbuf[r] + XOFF_HITS, cnt[r] * sizeof(lm_hit),
                               cudaMemcpyDefault, ctx->stream));
    const size_t n = std::min(total, cap);
    std::vector<lm_hit> all(total);
    if (total) CM_TRY(cudaMemcpyAsync(all.data(), L.d_stage.p, total * sizeof(lm_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CM_TRY(cudaStreamSynchronize(ctx->stream));
    sort_hits(all.data(), total);  // the reference's row-major order (pli/mod.rs:212-219), global rows
    if (n) std::memcpy(hits, all.data(), n * sizeof(lm_hit));
  }
  // acknowledge: the other ranks may reuse their hit regions
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    LM_CHECK(bind(L.ctx));
    ack_kernel<<<1, 32, 0, L.ctx->stream>>>((unsigned long long* const*)(L.xbuf + XOFF_PTR_ACK), comm->world, epoch);
    CM_TRY(cudaGetLastError());
    L.ctx->launches += 1;
    CM_TRY(cudaStreamSynchronize(L.ctx->stream));
  }
  comm->last_thr_epoch = epoch;
  if (overflow || grand > cap) return LM_ERR_OVERFLOW;
  return LM_OK;
}

int lm_comm_shard_from_striped(lm_comm* comm, int alphabet, const uint8_t* data, size_t rows, size_t wrap,
                               size_t stride, size_t length, lm_seq** slabs) {
  if (!comm || !slabs || (rows + wrap && !data) || stride < 32) return LM_ERR_INVALID_ARG;
  for (int li = 0; li < comm->n_local; ++li) slabs[li] = nullptr;
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    const size_t r0 = rows * (size_t)L.rank / comm->world, r1 = rows * (size_t)(L.rank + 1) / comm->world;
    // the slab's halo = the next `wrap` rows of the host matrix (base rows of the next slab, or the
    // sequence's own wrap rows for the last slab)
    lm_seq* s = nullptr;
    int st = rows == (length + 31) / 32
                 ? lm_seq_from_striped_slab(L.ctx, alphabet, data + r0 * stride, r1 - r0, wrap, stride, length, r0, rows, &s)
                 : LM_ERR_INVALID_ARG;
    if (st < 0) {
      for (int k = 0; k < li; ++k) { lm_seq_free(comm->local[k].ctx, slabs[k]); slabs[k] = nullptr; }
      return st;
    }
    slabs[li] = s;
  }
  return LM_OK;
}

int lm_comm_seq_synthetic(lm_comm* comm, int alphabet, uint64_t seed, uint64_t total_length, size_t wrap,
                          lm_seq** slabs) {
  if (!comm || !slabs) return LM_ERR_INVALID_ARG;
  const uint64_t rows = (total_length + 31) / 32;
  for (int li = 0; li < comm->n_local; ++li) slabs[li] = nullptr;
  for (int li = 0; li < comm->n_local; ++li) {
    lm_comm::Local& L = comm->local[li];
    const uint64_t r0 = rows * (uint64_t)L.rank / comm->world, r1 = rows * (uint64_t)(L.rank + 1) / comm->world;
    int st = lm_seq_synthetic(L.ctx, alphabet, seed, total_length, r0, r1 - r0, wrap, &slabs[li]);
    if (st < 0) {
      for (int k = 0; k < li; ++k) { lm_seq_free(comm->local[k].ctx, slabs[k]); slabs[k] = nullptr; }
      return st;
    }
  }
  return LM_OK;
}

}  // extern "C"
