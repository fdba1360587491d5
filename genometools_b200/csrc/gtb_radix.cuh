// gtb_radix.cuh -- hand-written onesweep-style LSD radix sort for (u64 key, u32 value)
// pairs on sm_100a.  No Thrust/CUB.
//
// One "pass" = one launch of rs_onesweep_kernel: every CTA takes a tile of
// RS_TILE elements (ticket order), ranks them by one digit with warp-level
// match/ballot histograms, obtains the global offset of each of its digit bins
// with a decoupled look-back over the per-tile status words, stages the tile in
// shared memory in digit order and writes it out so that runs of equal digits
// go to consecutive global addresses.  The digit histograms of ALL passes are
// computed up front by one rs_hist_kernel launch (one read of the keys).
//
// The first pass can read its keys from a "source" object instead of memory, so
// the 2-bit text -> 64-bit key generation is fused into the first pass (the keys
// are never written unsorted).
//
// Replaces (as the engine of every sorting step) the per-bucket CPU sorters of
// /root/reference/src/match/sfx-bentsedg.c:797-1341, sfx-shortreadsort.c and
// sfx-bltrie.c; as a stand-alone pair sort it is the counterpart of
// gt_radixsort_inplace_GtUwordPair (/root/reference/src/core/radix_sort.h:107).
#pragma once
#include "gtb_common.cuh"

namespace gtb {

constexpr int RS_NT    = 256;               // threads per CTA ( == number of bins )
constexpr int RS_IPT   = 16;                // items per thread
constexpr int RS_TILE  = RS_NT * RS_IPT;    // 4096 pairs per tile
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_BINS  = 256;
constexpr int RS_MAXPASS = 8;

// status word: [63:48] epoch  [47:46] flag  [45:0] value
constexpr u64 RS_FLAG_AGG  = 1ull << 46;
constexpr u64 RS_FLAG_INCL = 2ull << 46;
constexpr u64 RS_VALUE_MASK = (1ull << 46) - 1;

struct PassPlan {
  int npass;
  int shift[RS_MAXPASS];
  int bits[RS_MAXPASS];
};

// plan 8-bit digits covering key bits [begin_bit, end_bit)
static inline void plan_add_bits(PassPlan &p, int begin_bit, int end_bit)
{
  for (int b = begin_bit; b < end_bit; b += 8) {
    p.shift[p.npass] = b;
    p.bits[p.npass] = (end_bit - b) < 8 ? (end_bit - b) : 8;
    p.npass++;
  }
}

// ---- key sources -------------------------------------------------------------
struct PairSrc {                 // pairs already in memory
  const u64 *keys;
  const u32 *vals;
  __device__ __forceinline__ bool load(u64 idx, u64 &k, u32 &v) const
  { k = keys[idx]; v = vals[idx]; return true; }
  __device__ __forceinline__ bool load_key(u64 idx, u64 &k) const
  { k = keys[idx]; return true; }
};

constexpr size_t RS_SMEM_BYTES =
    sizeof(u64) * RS_TILE + sizeof(u32) * RS_TILE + sizeof(u32) * RS_WARPS * RS_BINS +
    sizeof(u32) * RS_BINS + sizeof(u64) * RS_BINS + sizeof(u32) * (RS_WARPS + 2);

// ---- histogram of all digits in one read ---------------------------------------
template <class Src>
__global__ void __launch_bounds__(RS_NT)
rs_hist_kernel(Src src, u64 N, PassPlan plan, unsigned long long *__restrict__ ghist)
{
  __shared__ u32 s_h[RS_MAXPASS * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAXPASS * RS_BINS; i += RS_NT) s_h[i] = 0;
  __syncthreads();
  const u64 ntiles = (N + RS_TILE - 1) / RS_TILE;
  const unsigned lane = lane_id();
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const u64 base = tile * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_IPT; k++) {
      const u64 idx = base + (u64) k * RS_NT + threadIdx.x;
      u64 key = 0;
      bool ok = idx < N;
      if (ok) ok = src.load_key(idx, key);
      for (int p = 0; p < plan.npass; p++) {
        const unsigned d = (unsigned) (key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
        const unsigned dd = ok ? d : 0x1ffu;
        int pred;
        __match_all_sync(FULL_MASK, dd, &pred);
        if (pred) {                       // whole warp in one bin: one atomic
          if (lane == 0 && ok) atomicAdd(&s_h[p * RS_BINS + d], 32u);
        } else if (ok) {
          atomicAdd(&s_h[p * RS_BINS + d], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RS_NT)
    if (s_h[i]) atomicAdd(&ghist[i], (unsigned long long) s_h[i]);
}

// exclusive scan of each pass's 256 counts -> global digit starts
__global__ void __launch_bounds__(RS_BINS)
rs_scan_kernel(const unsigned long long *__restrict__ ghist, u64 *__restrict__ gbase)
{
  __shared__ u64 scratch[RS_BINS / 32 + 1];
  const int p = blockIdx.x;
  u64 v = ghist[p * RS_BINS + threadIdx.x], total;
  u64 ex = block_exclusive_sum<RS_BINS, u64>(v, scratch, &total);
  gbase[p * RS_BINS + threadIdx.x] = ex;
}

// ---- one onesweep pass ---------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(RS_NT)
rs_onesweep_kernel(Src src, u64 *__restrict__ okeys, u32 *__restrict__ ovals, u64 N,
                   int shift, unsigned dmask, const u64 *__restrict__ gbase,
                   u64 *status, u32 epoch, u32 *ticket, u32 ticket_base)
{
  extern __shared__ __align__(16) unsigned char rs_smem[];
  u64 *s_keys     = reinterpret_cast<u64 *>(rs_smem);
  u64 *s_adj      = s_keys + RS_TILE;
  u32 *s_vals     = reinterpret_cast<u32 *>(s_adj + RS_BINS);
  u32 *s_whist    = s_vals + RS_TILE;                 // [RS_WARPS][RS_BINS]
  u32 *s_binstart = s_whist + RS_WARPS * RS_BINS;
  u32 *s_scan     = s_binstart + RS_BINS;             // RS_WARPS + 1
  u32 *s_tile     = s_scan + RS_WARPS + 1;

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) *s_tile = atomicAdd(ticket, 1u) - ticket_base;
  for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_NT) s_whist[i] = 0;
  __syncthreads();
  const u64 tile = *s_tile;
  const u64 base = tile * (u64) RS_TILE;
  const u32 count = (N - base) < (u64) RS_TILE ? (u32) (N - base) : (u32) RS_TILE;

  u64 key[RS_IPT];
  u32 val[RS_IPT];
  u32 rnk[RS_IPT];
  unsigned okmask = 0;

  // warp-striped load: warp w owns [w*32*IPT, (w+1)*32*IPT), lane-contiguous rows
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    const u32 idx = warp * (32u * RS_IPT) + (u32) k * 32u + lane;
    bool ok = idx < count;
    key[k] = 0; val[k] = 0;
    if (ok) ok = src.load(base + idx, key[k], val[k]);
    okmask |= (ok ? 1u : 0u) << k;
  }

  // stable ranking inside the warp: rows in order, lanes in order
  u32 *wh = s_whist + warp * RS_BINS;
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    const bool ok = (okmask >> k) & 1u;
    const unsigned d = (unsigned) (key[k] >> shift) & dmask;
    const unsigned dd = ok ? d : 0x1ffu;
    const unsigned peers = __match_any_sync(FULL_MASK, dd);
    const int leader = __ffs(peers) - 1;
    u32 old = 0;
    if (ok && (int) lane == leader) {
      old = wh[d];
      wh[d] = old + __popc(peers);
    }
    old = __shfl_sync(FULL_MASK, old, leader);
    rnk[k] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();

  // one thread per bin: warp offsets, tile-local bin starts, decoupled look-back
  {
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      const u32 c = s_whist[w * RS_BINS + tid];
      s_whist[w * RS_BINS + tid] = run;
      run += c;
    }
    const u32 cnt = run;
    u32 total;
    const u32 excl = block_exclusive_sum<RS_NT, u32>(cnt, s_scan, &total);
    s_binstart[tid] = excl;
    if (tid == 0) s_scan[RS_WARPS + 0] = total;   // keep the tile total (slot reused)
    const u64 ep = (u64) epoch << 48;
    u64 *mine = status + tile * RS_BINS + tid;
    u64 prefix = 0;
    if (tile == 0) {
      st_relaxed_u64(mine, ep | RS_FLAG_INCL | (u64) cnt);
    } else {
      st_relaxed_u64(mine, ep | RS_FLAG_AGG | (u64) cnt);
      for (u64 t = tile; t-- > 0; ) {
        const u64 *pp = status + t * RS_BINS + tid;
        u64 s;
        do {
          s = ld_relaxed_u64(pp);
        } while ((s >> 48) != (u64) epoch || (s & (3ull << 46)) == 0);
        prefix += s & RS_VALUE_MASK;
        if (s & RS_FLAG_INCL) break;
      }
      st_relaxed_u64(mine, ep | RS_FLAG_INCL | ((prefix + cnt) & RS_VALUE_MASK));
    }
    s_adj[tid] = gbase[tid] + prefix - (u64) excl;
  }
  __syncthreads();
  const u32 total = s_scan[RS_WARPS + 0];

  // stage the tile in digit order
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    if ((okmask >> k) & 1u) {
      const unsigned d = (unsigned) (key[k] >> shift) & dmask;
      const u32 slot = s_binstart[d] + wh[d] + rnk[k];
      s_keys[slot] = key[k];
      s_vals[slot] = val[k];
    }
  }
  __syncthreads();
  // coalesced scatter: consecutive threads -> consecutive slots -> runs per bin
  for (u32 i = tid; i < total; i += RS_NT) {
    const u64 kk = s_keys[i];
    const unsigned d = (unsigned) (kk >> shift) & dmask;
    const u64 dst = s_adj[d] + i;
    okeys[dst] = kk;
    ovals[dst] = s_vals[i];
  }
}

// ---- host-side driver ------------------------------------------------------------
struct RadixWork {
  u64 *status = nullptr;            // [status_tiles][256]
  u64  status_tiles = 0;
  u32 *ticket = nullptr;
  u32  ticket_base = 0;
  u32  epoch = 0;
  unsigned long long *ghist = nullptr;   // [8][256] device
  u64 *gbase = nullptr;                  // [8][256] device
  unsigned long long *h_hist = nullptr;  // pinned host copy
  // statistics
  u32 passes = 0;
  u64 pairs_moved = 0;
  u32 launches = 0;
  float ms_hist = 0, ms_radix = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static inline int radix_work_init(RadixWork &w, ErrBuf &err)
{
  GTB_CUDA(cudaMalloc(&w.ticket, sizeof(u32)));
  GTB_CUDA(cudaMemset(w.ticket, 0, sizeof(u32)));
  GTB_CUDA(cudaMalloc(&w.ghist, sizeof(unsigned long long) * RS_MAXPASS * RS_BINS));
  GTB_CUDA(cudaMalloc(&w.gbase, sizeof(u64) * RS_MAXPASS * RS_BINS));
  GTB_CUDA(cudaMallocHost(&w.h_hist, sizeof(unsigned long long) * RS_MAXPASS * RS_BINS));
  for (int i = 0; i < 4; i++) GTB_CUDA(cudaEventCreate(&w.ev[i]));
  return 0;
}

static inline void radix_work_free(RadixWork &w)
{
  cudaFree(w.status); cudaFree(w.ticket); cudaFree(w.ghist); cudaFree(w.gbase);
  if (w.h_hist) cudaFreeHost(w.h_hist);
  for (int i = 0; i < 4; i++) if (w.ev[i]) cudaEventDestroy(w.ev[i]);
  w = RadixWork();
}

static inline int radix_work_reserve(RadixWork &w, u64 nitems, ErrBuf &err)
{
  const u64 tiles = div_up(nitems, RS_TILE) + 1;
  if (tiles > w.status_tiles) {
    if (w.status) GTB_CUDA(cudaFree(w.status));
    w.status = nullptr; w.status_tiles = 0;
    GTB_CUDA(cudaMalloc(&w.status, sizeof(u64) * RS_BINS * tiles));
    GTB_CUDA(cudaMemset(w.status, 0, sizeof(u64) * RS_BINS * tiles));
    w.status_tiles = tiles;
    w.epoch = 0;
  }
  return 0;
}

template <class Src>
static int rs_launch_pass(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc,
                          u64 *okeys, u32 *ovals, int shift, int bits, int passidx,
                          ErrBuf &err)
{
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    GTB_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<Src>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int) RS_SMEM_BYTES));
    attr_set = true;
  }
  const u64 tiles = div_up(nsrc, RS_TILE);
  if (tiles == 0) return 0;
  if (tiles > w.status_tiles) { err.set("radix: status array too small"); return -1; }
  if (++w.epoch >= 0xffffu) {           // epoch space exhausted: start over
    GTB_CUDA(cudaMemsetAsync(w.status, 0, sizeof(u64) * RS_BINS * w.status_tiles, st));
    w.epoch = 1;
  }
  rs_onesweep_kernel<Src><<<(unsigned) tiles, RS_NT, RS_SMEM_BYTES, st>>>(
      src, okeys, ovals, nsrc, shift, (1u << bits) - 1u, w.gbase + passidx * RS_BINS,
      w.status, w.epoch, w.ticket, w.ticket_base);
  GTB_LAUNCH_CHECK();
  w.ticket_base += (u32) tiles;
  w.passes++; w.launches++;
  return 0;
}

// Sort pairs by the digits of `plan` (LSD, stable).  The first executed pass reads
// from `src` (nsrc items, some possibly invalid); the sorted pairs end up in
// kbuf[*res]/vbuf[*res], *nout = number of valid pairs.  Passes whose digit is the
// same for every key are skipped.  Synchronises the stream once (histogram
// read-back).
template <class Src>
static int radix_sort(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc,
                      u64 *kbuf[2], u32 *vbuf[2], const PassPlan &plan,
                      int *res, u64 *nout, ErrBuf &err)
{
  *res = 0; *nout = 0;
  if (nsrc == 0) return 0;
  if (plan.npass < 1 || plan.npass > RS_MAXPASS) { err.set("radix: bad pass plan"); return -1; }
  GTB_TRY(radix_work_reserve(w, nsrc, err));
  GTB_CUDA(cudaEventRecord(w.ev[0], st));
  GTB_CUDA(cudaMemsetAsync(w.ghist, 0, sizeof(unsigned long long) * RS_MAXPASS * RS_BINS, st));
  {
    u64 tiles = div_up(nsrc, RS_TILE);
    unsigned grid = (unsigned) (tiles < 148ull * 8 ? tiles : 148ull * 8);
    rs_hist_kernel<Src><<<grid, RS_NT, 0, st>>>(src, nsrc, plan, w.ghist);
    GTB_LAUNCH_CHECK();
    rs_scan_kernel<<<plan.npass, RS_BINS, 0, st>>>(w.ghist, w.gbase);
    GTB_LAUNCH_CHECK();
    w.launches += 2;
  }
  GTB_CUDA(cudaMemcpyAsync(w.h_hist, w.ghist, sizeof(unsigned long long) * plan.npass * RS_BINS,
                           cudaMemcpyDeviceToHost, st));
  GTB_CUDA(cudaEventRecord(w.ev[1], st));
  GTB_CUDA(cudaStreamSynchronize(st));
  u64 total = 0;
  for (int d = 0; d < RS_BINS; d++) total += w.h_hist[d];
  *nout = total;
  bool skip[RS_MAXPASS];
  int nexec = 0;
  for (int p = 0; p < plan.npass; p++) {
    skip[p] = false;
    for (int d = 0; d < RS_BINS; d++)
      if (w.h_hist[p * RS_BINS + d] == total) { skip[p] = true; break; }
    if (!skip[p]) nexec++;
  }
  if (nexec == 0) skip[plan.npass - 1] = false;   // materialise at least once
  if (total == 0) {
    float ms = 0; cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]); w.ms_hist += ms;
    return 0;
  }
  int cur = -1;                                    // -1: data still in `src`
  GTB_CUDA(cudaEventRecord(w.ev[2], st));
  for (int p = 0; p < plan.npass; p++) {
    if (skip[p]) continue;
    if (cur < 0) {
      GTB_TRY(rs_launch_pass(w, st, src, nsrc, kbuf[0], vbuf[0], plan.shift[p], plan.bits[p], p, err));
      cur = 0;
    } else {
      PairSrc ps{kbuf[cur], vbuf[cur]};
      GTB_TRY(rs_launch_pass(w, st, ps, total, kbuf[cur ^ 1], vbuf[cur ^ 1], plan.shift[p],
                             plan.bits[p], p, err));
      cur ^= 1;
    }
    w.pairs_moved += total;
  }
  GTB_CUDA(cudaEventRecord(w.ev[3], st));
  GTB_CUDA(cudaEventSynchronize(w.ev[3]));
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]); w.ms_hist += ms;
    cudaEventElapsedTime(&ms, w.ev[2], w.ev[3]); w.ms_radix += ms;
  }
  *res = cur;
  return 0;
}

} // namespace gtb
