// upload.cu -- host -> device transfer of a large striped DNA matrix (lm_seq_from_striped).
//
// A StripedSequence in host memory is 1 byte per symbol (seq.rs:262-271) and a fused scan reads it at
// several TB/s, so a call that starts from a HOST matrix is bound by PCIe (~55 GB/s here), not by the
// scan.  DNA symbols other than N need 2 bits, so the upload is done from BOTH ends of the matrix at once:
//
//   front   whenever the link would be idle, the copy engine pulls a chunk of the caller's
//           (pinned) matrix as it is                                                      1 B / symbol
//   back    host threads pack chunks 4 symbols per byte (hostpack.cpp) into a ring of
//           pinned slots; each goes over PCIe packed and unpack2_kernel expands it into
//           its place in the device matrix                                                0.25 B / symbol
//
// until the two ends meet: the link is never idle, packed chunks never wait behind plain ones, and the
// split adapts to the machine (host cores, memory bandwidth, other ranks on the same host).  A chunk
// that holds a symbol > 3 (N) is sent as it is.  The device matrix is byte-identical to a plain copy
// (tests/test_gpu_parity.py::test_packed_upload_*).  Pageable source buffers go through the packed path
// only (the ring is the pinned staging a plain copy would need anyway).  LM_UPLOAD_DEBUG=1 in the environment
// prints one timing line per upload to stderr.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif

#include "internal.cuh"

namespace lmi {

bool pack2(const uint8_t* src, size_t n, uint8_t* dst);  // hostpack.cpp

namespace {

constexpr size_t CHUNK = 4u << 20;        // symbols (= bytes of the matrix) per chunk
constexpr size_t SLOT = CHUNK / 4;        // packed bytes per chunk
constexpr int RAW_DEPTH = 2;              // plain copies kept in flight
constexpr int MAX_THREADS = 64;

// 4 symbols of a packed byte -> 4 bytes
__device__ __forceinline__ uint32_t spread4(uint32_t b) {
  uint32_t t = (b | (b << 12)) & 0x000F000Fu;
  return (t | (t << 6)) & 0x03030303u;
}

__global__ void __launch_bounds__(256) unpack2_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, size_t n_words) {
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += step) {
    const uint32_t w = __ldg(&in[i]);
    uint4 o;
    o.x = spread4(w & 0xffu);
    o.y = spread4((w >> 8) & 0xffu);
    o.z = spread4((w >> 16) & 0xffu);
    o.w = spread4(w >> 24);
    out[i] = o;
  }
}

}  // namespace
}  // namespace lmi

struct lm_uploader {
  int nslots = 0;
  uint8_t* h_ring = nullptr;  // pinned, nslots x SLOT
  uint8_t* d_ring = nullptr;  // device, nslots x SLOT
  // ALL host -> device copies go through s_raw in issue order (two streams that both copy were observed to
  // starve one another: the packed chunks did not move until the plain stream ran dry); the expansion
  // kernels run on s_pack behind per-slot events, so the copy stream never waits for a kernel
  cudaStream_t s_raw = nullptr, s_pack = nullptr;
  std::vector<cudaEvent_t> slot_ev, copy_ev;  // per slot: expanded on the device / packed bytes landed
  cudaEvent_t raw_ev[lmi::RAW_DEPTH] = {};
  cudaEvent_t ev_begin = nullptr, ev_raw = nullptr, ev_pack = nullptr;
};

void free_uploader(lm_uploader* u) {
  if (!u) return;
  if (u->h_ring) cudaFreeHost(u->h_ring);
  if (u->d_ring) cudaFree(u->d_ring);
  for (auto e : u->slot_ev)
    if (e) cudaEventDestroy(e);
  for (auto e : u->copy_ev)
    if (e) cudaEventDestroy(e);
  for (auto e : u->raw_ev)
    if (e) cudaEventDestroy(e);
  for (auto e : {u->ev_begin, u->ev_raw, u->ev_pack})
    if (e) cudaEventDestroy(e);
  if (u->s_raw) cudaStreamDestroy(u->s_raw);
  if (u->s_pack) cudaStreamDestroy(u->s_pack);
  delete u;
}

namespace lmi {

namespace {

int upload_threads(const lm_ctx* ctx) {
  int t = ctx->opt_upload_threads;
  if (t <= 0) {
    unsigned hw = std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;  // (a process confined to fewer CPUs -- taskset, cpusets, one rank of many -- must not oversubscribe
    CPU_ZERO(&set);  //  them: twice as many packing threads as cores measured slower than the plain copy)
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
      const unsigned allowed = (unsigned)CPU_COUNT(&set);
      if (allowed && (hw == 0 || allowed < hw)) hw = allowed;
    }
#endif
    t = hw > 2 ? (int)hw - 1 : 1;  // one core is left to the thread that feeds the copy engine
  }
  return t > MAX_THREADS ? MAX_THREADS : t;
}

int get_uploader(lm_ctx* ctx, int threads, lm_uploader** out) {
  const int want = 2 * threads + 8;  // (a packed chunk may wait behind plain ones: spare slots keep the packers busy)
  ctx->upload_ring_created = false;
  if (ctx->uploader && ctx->uploader->nslots >= want) { *out = ctx->uploader; return LM_OK; }
  ctx->upload_ring_created = true;
  free_uploader(ctx->uploader);
  ctx->uploader = nullptr;
  lm_uploader* u = new (std::nothrow) lm_uploader();
  if (!u) return LM_ERR_OOM;
  u->nslots = want;
  u->slot_ev.assign(want, nullptr);
  u->copy_ev.assign(want, nullptr);
  cudaError_t e = cudaHostAlloc((void**)&u->h_ring, (size_t)want * SLOT, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc((void**)&u->d_ring, (size_t)want * SLOT);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&u->s_raw, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&u->s_pack, cudaStreamNonBlocking);
  for (int i = 0; i < want && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&u->slot_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < want && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&u->copy_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < RAW_DEPTH && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&u->raw_ev[i], cudaEventDisableTiming);
  for (cudaEvent_t* p : {&u->ev_begin, &u->ev_raw, &u->ev_pack})
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(p, cudaEventDisableTiming);
  if (e != cudaSuccess) { free_uploader(u); return fail(ctx, e, "packed upload: staging"); }
  ctx->uploader = u;
  *out = u;
  return LM_OK;
}

}  // namespace

// Does this upload take the two-ended path?  LM_OPT_UPLOAD = 1 (default) decides from what this context has
// measured: on a host whose memory system the copy engines of several ranks already saturate (4 ranks x 53 GB/s
// on a 32-thread host), packing threads only take bandwidth away from them (measured: 175 vs 212 Gbp/s), while
// one or two ranks per host gain 1.3-2.4x.  The first large upload goes two-ended, the second plain, later
// ones take the faster of the two; the other is tried again every 32nd upload.  Pageable sources always pack.
bool upload_wants_packed(lm_ctx* ctx, int alphabet, size_t stride, size_t bytes, const void* src) {
  if (ctx->opt_upload == 0 || alphabet != LM_DNA || stride != 32 || bytes < (size_t)ctx->opt_upload_min_bytes) return false;
  if (ctx->opt_upload != 1) return true;
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) return true;
  const uint32_t call = ctx->upload_calls++;
  if (ctx->upload_rate_packed == 0.0) return true;
  if (ctx->upload_rate_plain == 0.0) return false;
  // (a tie goes to the plain copy: it costs no host threads)
  const bool packed_is_faster = ctx->upload_rate_packed >= 1.05 * ctx->upload_rate_plain;
  return (call % 32 == 31) ? !packed_is_faster : packed_is_faster;
}
void upload_note_rate(lm_ctx* ctx, bool packed, size_t bytes, double seconds) {
  if (seconds <= 0.0) return;
  if (packed && ctx->upload_ring_created) return;  // (pinned allocations took part of that time: measure the next one)
  (packed ? ctx->upload_rate_packed : ctx->upload_rate_plain) = (double)bytes / seconds;
}

// Copy `bytes` DNA symbols (a striped matrix, or the flat symbols of lm_seq_from_symbols) from host memory to
// `dst` (device, 16-byte aligned, allocated before or in order on ctx->stream).  A tail of fewer than 64 bytes is
// copied as it is.  On return the copies are ordered before later work on ctx->stream.
int upload_striped_dna(lm_ctx* ctx, uint8_t* dst, const uint8_t* src, size_t bytes_all) {
  const size_t bytes = bytes_all & ~(size_t)63;  // whole 16-symbol words of the expansion kernel, 4 per packed word
  if (bytes_all > bytes) {
    cudaError_t e = cudaMemcpyAsync(dst + bytes, src + bytes, bytes_all - bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, e, "packed upload: tail");
  }
  if (bytes == 0) return LM_OK;
  const int threads = upload_threads(ctx);
  lm_uploader* U = nullptr;
  LM_CHECK(get_uploader(ctx, threads, &U));
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                      attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const bool raw_front = pinned && ctx->opt_upload != 2;  // 2 = packed chunks only
  const bool raw_gaps_only = ctx->opt_upload == 3;       // 3 = plain chunks only while no packed copy is pending
  const int64_t nchunks = (int64_t)((bytes + CHUNK - 1) / CHUNK);
  auto chunk_len = [&](int64_t c) { return (size_t)c * CHUNK + CHUNK <= bytes ? CHUNK : bytes - (size_t)c * CHUNK; };

  // the device matrix was allocated in stream order on ctx->stream
  LM_TRY(cudaEventRecord(U->ev_begin, ctx->stream));
  LM_TRY(cudaStreamWaitEvent(U->s_raw, U->ev_begin, 0));
  LM_TRY(cudaStreamWaitEvent(U->s_pack, U->ev_begin, 0));

  const bool dbg = getenv("LM_UPLOAD_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  std::atomic<long long> wait_us{0};
  struct Ready { int64_t chunk; int slot; bool ok; };
  std::mutex mu;
  std::condition_variable cv_ready, cv_free;
  std::deque<int> free_slots;
  std::deque<Ready> ready;
  int64_t front = 0, back = nchunks;  // chunks [front, back) are unclaimed
  bool stop = false;
  for (int i = 0; i < U->nslots; ++i) free_slots.push_back(i);

  auto worker = [&]() {
    for (;;) {
      int64_t c;
      int slot;
      {
        std::unique_lock<std::mutex> l(mu);
        const double w0 = dbg ? now() : 0.0;
        cv_free.wait(l, [&] { return stop || back <= front || !free_slots.empty(); });
        if (dbg) wait_us += (long long)((now() - w0) * 1000.0);
        if (stop || back <= front) return;
        c = --back;
        slot = free_slots.front();
        free_slots.pop_front();
      }
      const bool ok = pack2(src + (size_t)c * CHUNK, chunk_len(c), U->h_ring + (size_t)slot * SLOT);
      {
        std::lock_guard<std::mutex> l(mu);
        ready.push_back({c, slot, ok});
      }
      cv_ready.notify_one();
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(threads);
  try {
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
  } catch (...) {  // (thread limit of the process: carry on with the threads that did start -- the copy engine takes
  }                //  whatever they do not; nothing may unwind across the C ABI)
  const bool no_packers = pool.empty();

  cudaError_t err = cudaSuccess;
  uint64_t pcie = 0, n_packed = 0, n_raw = 0;
  std::deque<int> packed_inflight;  // slots, in issue order
  int raw_inflight = 0, raw_head = 0;  // raw_ev ring: [raw_head, raw_head + raw_inflight)
  auto issue_raw = [&](int64_t c) {
    const size_t len = chunk_len(c);
    while (raw_inflight >= RAW_DEPTH && err == cudaSuccess) {  // (chunks with N only: the front claims when there is room)
      err = cudaEventSynchronize(U->raw_ev[raw_head]);
      raw_head = (raw_head + 1) % RAW_DEPTH;
      --raw_inflight;
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(dst + (size_t)c * CHUNK, src + (size_t)c * CHUNK, len, cudaMemcpyHostToDevice, U->s_raw);
    if (err == cudaSuccess) err = cudaEventRecord(U->raw_ev[(raw_head + raw_inflight) % RAW_DEPTH], U->s_raw);
    ++raw_inflight;
    pcie += len;
    ++n_raw;
  };
  int64_t issued = 0;
  int last_packed = -1;  // slot of the packed chunk issued last
  while (issued < nchunks && err == cudaSuccess) {
    // recycle what the device is done with
    while (!packed_inflight.empty() && cudaEventQuery(U->slot_ev[packed_inflight.front()]) == cudaSuccess) {
      {
        std::lock_guard<std::mutex> l(mu);
        free_slots.push_back(packed_inflight.front());
      }
      cv_free.notify_one();
      packed_inflight.pop_front();
    }
    while (raw_inflight && cudaEventQuery(U->raw_ev[raw_head]) == cudaSuccess) {
      raw_head = (raw_head + 1) % RAW_DEPTH;
      --raw_inflight;
    }
    cudaGetLastError();  // (cudaErrorNotReady of the queries)
    // (policy 3 only: plain chunks are issued only while no packed copy is pending, so that a packed chunk --
    // four times the symbols per byte -- never queues behind plain ones.  Measured equal to policies 1 and 2 on
    // a 16-thread host, where the packing threads, not the link, are the limit.)
    const bool packed_copy_pending = last_packed >= 0 && cudaEventQuery(U->copy_ev[last_packed]) != cudaSuccess;
    cudaGetLastError();
    Ready r{-1, 0, false};
    int64_t raw_chunk = -1;
    bool exhausted = false;
    {
      std::lock_guard<std::mutex> l(mu);
      if (!ready.empty()) {
        r = ready.front();
        ready.pop_front();
      } else if ((raw_front || no_packers) && raw_inflight < RAW_DEPTH && front < back &&
                 !(raw_gaps_only && packed_copy_pending && !no_packers)) {
        raw_chunk = front++;
        exhausted = front >= back;
      }
    }
    if (r.chunk >= 0) {
      if (r.ok) {
        const size_t len = chunk_len(r.chunk);
        uint8_t* d_slot = U->d_ring + (size_t)r.slot * SLOT;
        err = cudaMemcpyAsync(d_slot, U->h_ring + (size_t)r.slot * SLOT, len / 4, cudaMemcpyHostToDevice, U->s_raw);
        if (err == cudaSuccess) err = cudaEventRecord(U->copy_ev[r.slot], U->s_raw);
        if (err == cudaSuccess) err = cudaStreamWaitEvent(U->s_pack, U->copy_ev[r.slot], 0);
        if (err == cudaSuccess) {
          const size_t n_words = len / 16;
          const unsigned grid = (unsigned)std::min<size_t>((n_words + 255) / 256, (size_t)ctx->sm_count * 8);
          unpack2_kernel<<<grid, 256, 0, U->s_pack>>>((const uint32_t*)d_slot, (uint4*)(dst + (size_t)r.chunk * CHUNK), n_words);
          err = cudaGetLastError();
          ctx->launches += 1;
        }
        if (err == cudaSuccess) err = cudaEventRecord(U->slot_ev[r.slot], U->s_pack);
        packed_inflight.push_back(r.slot);
        last_packed = r.slot;
        pcie += len / 4;
        ++n_packed;
      } else {  // holds N (or anything else that is not 0..3): sent as it is
        {
          std::lock_guard<std::mutex> l(mu);
          free_slots.push_back(r.slot);
        }
        cv_free.notify_one();
        issue_raw(r.chunk);
      }
      ++issued;
      continue;
    }
    if (raw_chunk >= 0) {
      issue_raw(raw_chunk);
      ++issued;
      if (exhausted) cv_free.notify_all();  // idle workers may leave
      continue;
    }
    std::unique_lock<std::mutex> l(mu);
    if (back <= front) cv_free.notify_all();  // nothing left to claim: idle workers may leave
    cv_ready.wait_for(l, std::chrono::microseconds(25), [&] { return !ready.empty(); });
  }
  {
    std::lock_guard<std::mutex> l(mu);
    stop = true;
  }
  cv_free.notify_all();
  const double t_issued = now();
  for (auto& t : pool) t.join();
  if (dbg) {
    cudaStreamSynchronize(U->s_raw);
    cudaStreamSynchronize(U->s_pack);
    fprintf(stderr, "[upload] %.1f MB: spawn+issue %.2f ms, drain %.2f ms, workers waited for slots %.2f ms in total (%d threads), packed %llu plain %llu\n",
            bytes / 1e6, t_issued - t_begin, now() - t_issued, wait_us.load() / 1000.0, threads, (unsigned long long)n_packed,
            (unsigned long long)n_raw);
  }
  if (err == cudaSuccess) err = cudaEventRecord(U->ev_raw, U->s_raw);
  if (err == cudaSuccess) err = cudaEventRecord(U->ev_pack, U->s_pack);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(ctx->stream, U->ev_raw, 0);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(ctx->stream, U->ev_pack, 0);
  if (err != cudaSuccess) {
    cudaStreamSynchronize(U->s_raw);
    cudaStreamSynchronize(U->s_pack);
    return fail(ctx, err, "packed upload");
  }
  ctx->upload_stats[0] = pcie + (bytes_all - bytes);
  ctx->upload_stats[1] = n_packed;
  ctx->upload_stats[2] = n_raw;
  ctx->upload_stats[3] = (uint64_t)threads;
  return LM_OK;
}

}  // namespace lmi
