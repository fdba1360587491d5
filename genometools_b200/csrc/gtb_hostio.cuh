// gtb_hostio.cuh -- moving results from HBM into the caller's host buffers.
//
// The suffix table lives in HBM as uint32 (n + 1 < 2^32) while the reference's .suf holds
// uint64 entries (gt_suffixsortspace_to_file, /root/reference/src/match/sfx-suffixgetset.c:462-477).
// PCIe is the bottleneck of every end-to-end run (c4: 24.8 GB of uint64 against 187 ms of
// kernels), so only the 4 significant bytes per entry cross the bus: chunks are copied into
// pinned staging buffers owned by the handle and widened into the caller's buffer by a few
// host threads with streaming stores while the next chunks are in flight.  The same staging
// serves byte tables (.lcp, .bwt) whose destination is pageable memory -- a cudaMemcpy straight
// into pageable memory is bounced through one driver thread at a fraction of the bus rate.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "gtb_common.cuh"

namespace gtb {

struct HostStage {
  static constexpr int NB = 4;                     // chunks in flight
  // bytes of device data per chunk (GTB200_STAGE_KB).  Small on purpose: 4 x 8 MiB stay in the
  // last-level cache, so the widening threads read what the DMA engine just wrote without a trip
  // to DRAM (measured on the pool's hosts for c4: 8 MiB 334 ms, 64 MiB 372 ms, 4 MiB 419 ms)
  size_t chunk = size_t(8) << 20;
  void *buf[NB] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[NB] = {nullptr, nullptr, nullptr, nullptr};
  bool ready = false;
  int nthreads = 0;
  int thread_div = 1;          // handles of one job that copy at the same time share the host's cores

  int ensure(ErrBuf &err)
  {
    if (ready) return 0;
    if (const char *e = getenv("GTB200_STAGE_KB")) {
      const long kb = atol(e);
      if (kb >= 64 && kb <= (1l << 20)) chunk = (size_t) kb << 10;
    }
    for (int i = 0; i < NB; i++) {
      GTB_CUDA(cudaHostAlloc(&buf[i], chunk, cudaHostAllocDefault));
      GTB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    int t = 0;
    if (const char *e = getenv("GTB200_HOST_THREADS")) t = atoi(e);
    if (t <= 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      t = hw > 3 ? (int) hw - 2 : 1;
      if (t > 32) t = 32;
      if (thread_div > 1) t = t / thread_div > 0 ? t / thread_div : 1;
    }
    nthreads = t > 128 ? 128 : t;
    ready = true;
    return 0;
  }
  void release()
  {
    for (int i = 0; i < NB; i++) {
      if (buf[i]) cudaFreeHost(buf[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      buf[i] = nullptr; ev[i] = nullptr;
    }
    ready = false;
  }
};

// chunks of one transfer dealt from both ends: the staged (narrow) path takes them from the
// front, the direct (wide) path from the back, until they meet -- whichever resource (host
// memory system or PCIe) is faster on this machine ends up with the larger share
struct ChunkDealer {
  std::mutex mu;
  u64 lo = 0, hi = 0;
  explicit ChunkDealer(u64 nchunks) : hi(nchunks) {}
  bool take_front(u64 *k) { std::lock_guard<std::mutex> g(mu); if (lo >= hi) return false; *k = lo++; return true; }
  bool take_back(u64 *k) { std::lock_guard<std::mutex> g(mu); if (lo >= hi) return false; *k = --hi; return true; }
  u64 front_taken() { std::lock_guard<std::mutex> g(mu); return lo; }
};

// dst[i] = src[i] for i < n with non-temporal stores, widest vector unit of the host (gtb_widen.cpp)
extern "C" void gtb_widen_u32_u64(const uint32_t *src, uint64_t *dst, uint64_t n, int which);
static inline void widen_u32_u64_host(const u32 *src, u64 *dst, u64 n) { gtb_widen_u32_u64(src, dst, n, 0); }

// Copy `count` elements of `elem` bytes from device memory to host memory through the pinned
// staging buffers; widen = true turns uint32 elements into uint64 on the way.  All device work
// is queued on `st` (the handle's stream), so the copy sees the finished results.  With a
// dealer only the chunks (of hs.chunk / elem elements) it hands out are moved.
static int staged_d2h(HostStage &hs, cudaStream_t st, const void *dsrc, void *hdst, u64 count,
                      unsigned elem, bool widen, ErrBuf &err, ChunkDealer *dealer = nullptr)
{
  if (count == 0) return 0;
  GTB_TRY(hs.ensure(err));
  const u64 per = hs.chunk / elem;                         // elements per chunk
  const u64 nchunks = div_up(count, per);
  const u64 outelem = widen ? 8 : elem;
  ChunkDealer own(nchunks);
  if (!dealer) dealer = &own;
  // small transfers: this thread alone
  int T = hs.nthreads;
  if (count * (u64) elem < (u64(4) << 20)) T = 1;

  std::mutex mu;
  std::condition_variable cv;
  std::vector<u64> chunk_of(nchunks);                      // chunk_of[s] = chunk moved as number s
  u64 arrived = 0;                                         // numbers whose bytes are in their staging buffer
  bool failed = false, finished = false;
  std::atomic<u64> done[HostStage::NB];
  for (auto &d : done) d.store(0);

  auto slice = [&](u64 s, int t) {
    const u64 k = chunk_of[s];
    const u64 c = count - k * per < per ? count - k * per : per;
    const u64 lo = c * (u64) t / (u64) T, hi = c * (u64) (t + 1) / (u64) T;
    const u8 *src = static_cast<const u8 *>(hs.buf[s % HostStage::NB]) + lo * elem;
    u8 *d = static_cast<u8 *>(hdst) + (k * per + lo) * outelem;
    if (widen) widen_u32_u64_host(reinterpret_cast<const u32 *>(src), reinterpret_cast<u64 *>(d), hi - lo);
    else memcpy(d, src, (hi - lo) * elem);
    done[s % HostStage::NB].fetch_add(1, std::memory_order_release);
  };
  auto work = [&](int t) {
    for (u64 s = 0;; s++) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return arrived > s || failed || finished; });
        if (failed || arrived <= s) return;                // (finished: nothing more will arrive)
      }
      slice(s, t);
    }
  };
  std::vector<std::thread> pool;
  int rc = 0;
  try {
    pool.reserve((size_t) T);
    for (int t = 1; t < T; t++) pool.emplace_back(work, t);
  } catch (const std::exception &e) {          // no thread to be had: nothing was copied yet
    { std::lock_guard<std::mutex> lk(mu); failed = true; }
    cv.notify_all();
    for (auto &th : pool) th.join();
    err.set("could not start the host threads of the result copy: %s", e.what());
    return -1;
  }

  auto fail = [&](cudaError_t e, const char *what) {
    err.set("%s failed: %s", what, cudaGetErrorString(e));
    { std::lock_guard<std::mutex> lk(mu); failed = true; }
    cv.notify_all();
    rc = -1;
  };
  u64 issued = 0;                                          // numbers handed to the copy engine
  bool dry = false;                                        // the dealer has nothing left
  for (u64 s = 0; rc == 0; s++) {
    while (!dry && issued < s + HostStage::NB && rc == 0) {
      u64 k;
      if (!dealer->take_front(&k)) { dry = true; break; }
      const int b = (int) (issued % HostStage::NB);
      // the buffer's previous chunk must have left it (all T slices taken)
      const u64 uses = issued / HostStage::NB;
      while (done[b].load(std::memory_order_acquire) < uses * (u64) T) std::this_thread::yield();
      const u64 c = count - k * per < per ? count - k * per : per;
      chunk_of[issued] = k;
      cudaError_t e = cudaMemcpyAsync(hs.buf[b], static_cast<const u8 *>(dsrc) + k * per * elem, c * elem,
                                      cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(hs.ev[b], st);
      if (e != cudaSuccess) { fail(e, "staged device-to-host copy"); break; }
      issued++;
    }
    if (rc != 0 || s >= issued) break;
    cudaError_t e = cudaEventSynchronize(hs.ev[s % HostStage::NB]);
    if (e != cudaSuccess) { fail(e, "cudaEventSynchronize"); break; }
    { std::lock_guard<std::mutex> lk(mu); arrived = s + 1; }
    cv.notify_all();
    slice(s, 0);                                           // this thread takes slice 0 of the chunk
  }
  { std::lock_guard<std::mutex> lk(mu); finished = true; }
  cv.notify_all();
  for (auto &th : pool) th.join();
  return rc;
}

// true when `p` is page-locked host memory the copy engine can write directly
static inline bool host_pointer_is_pinned(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

} // namespace gtb
