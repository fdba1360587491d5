// gtb_radix.cuh -- hand-written onesweep-style LSD radix sort for (u64 key, u32 value)
// pairs on sm_100a.  No Thrust/CUB.
//
// One "pass" = one launch of rs_onesweep_kernel: every CTA takes a tile of
// Cfg::TILE elements (ticket order), ranks them by one 8-bit digit with warp-level
// match/ballot histograms, obtains the global offset of each of its digit bins
// with a decoupled look-back over the per-tile status words, stages the tile in
// shared memory in digit order and writes it out so that runs of equal digits
// go to consecutive global addresses.  The digit histograms of ALL passes are
// computed up front by one rs_hist_kernel launch (one read of the keys).
//
// Look-back: with ~450 tiles in flight on 148 SMs a tile usually has to walk over several
// predecessors that have published their aggregate but not yet their inclusive prefix,
// and every hop is an L2 round trip.  Each bin thread therefore fetches Cfg::LB
// predecessor words at once with weak L1-bypassing loads (LDG.NA -- the strong
// ld.relaxed.gpu form completes one at a time per thread on sm_100a) and consumes them in
// order; only a word that is not ready yet is re-read with a strong load.
// (Also measured with tools/rs_bench.cu and not kept: dedicated scanner CTAs that turn
// aggregates into prefixes -- correct, but latency-bound at ~80 cycles per tile.)
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
#include <atomic>
#include <type_traits>
#include "gtb_common.cuh"

namespace gtb {

constexpr int RS_BINS  = 256;
constexpr int RS_MAXPASS = 8;

// how many times the context of a device was destroyed by this library (gtb_release_devices): what is cached
// per context -- the shared-memory attribute of a kernel -- is set again when the count has moved on
inline std::atomic<unsigned> g_context_generation[64];

// Kernel shape.  NT threads x IPT pairs per tile; MINB = CTAs per SM the register
// allocation is bounded for; VAL_EARLY: values are loaded together with the keys
// (IPT more live registers) instead of just before they are staged; LB: predecessor
// status words fetched at once by the look-back (0 = timing experiment without it).
// (Measured on B200 with tools/rs_bench.cu: match.any.sync is slower than the ballot
// loop for 8-bit digits, shared-memory atomics cost 2 cycles per active lane on the
// LSU and lose against load + leader store -- neither is kept.)
template <int NT_, int IPT_, int MINB_, bool VAL_EARLY_, int LB_ = 4, int PF_ = 0>
struct RsCfg {
  static constexpr int NT = NT_, IPT = IPT_, MINB = MINB_, TILE = NT_ * IPT_, WARPS = NT_ / 32, LB = LB_;
  // PF > 0: a CTA asks L2 for the pairs of the tile PF tickets ahead (one 128-byte line per thread) --
  // the tile some CTA will load about one tile time later
  static constexpr int PF = PF_;
  static constexpr bool VAL_EARLY = VAL_EARLY_;
  static_assert(NT_ >= RS_BINS && NT_ % 32 == 0, "one thread per bin needed");
  static_assert(IPT_ * 32 < 4096, "warp-local ranks are kept in 12 bits");
  // staged keys + values, warp histograms (16-bit: a warp ranks at most 32*IPT <= 65535 pairs;
  // at least 4 KB because the per-bin output pointer tables reuse the space), scan scratch
  static constexpr size_t HIST = sizeof(unsigned short) * WARPS * RS_BINS < 4096 ? 4096 : sizeof(unsigned short) * WARPS * RS_BINS;
  static constexpr size_t SMEM = sizeof(u64) * TILE + sizeof(u32) * TILE + HIST + sizeof(u32) * (WARPS + 4);
};

// the shape the library uses (chosen with tools/rs_bench.cu on B200, see profiles/)
#ifndef GTB_RS_DEFAULT_CFG
#define GTB_RS_DEFAULT_CFG RsCfg<384, 16, 2, false>
#endif
typedef GTB_RS_DEFAULT_CFG RsDefault;
constexpr int RS_TILE = RsDefault::TILE;      // granularity the status array is sized for

// status word: [63:48] epoch  [47:46] flag  [45:0] value
constexpr u64 RS_FLAG_AGG  = 1ull << 46;
constexpr u64 RS_FLAG_INCL = 2ull << 46;
constexpr u64 RS_VALUE_MASK = (1ull << 46) - 1;

struct PassPlan {
  int npass;
  int shift[RS_MAXPASS];
  int bits[RS_MAXPASS];
  bool padded = false;    // keys are zero in the unused upper bits of every partial digit's byte
};

// plan digits of at most 8 bits covering key bits [begin_bit, end_bit)
static inline void plan_add_bits(PassPlan &p, int begin_bit, int end_bit)
{
  for (int b = begin_bit; b < end_bit; b += 8) {
    p.shift[p.npass] = b;
    p.bits[p.npass] = (end_bit - b) < 8 ? (end_bit - b) : 8;
    p.npass++;
  }
}

// ---- key sources -------------------------------------------------------------
// A source may generate the keys of consecutive items much cheaper than one by one (rolling
// keys over the packed text): BLOCKED_GEN sources fill a thread's IPT consecutive items with
// gen_block(first item, items, out) -- RS_INVALID_KEY marks items that do not take part --
// and the tile is transposed through shared memory into the warp-striped order of the ranking.
constexpr u64 RS_INVALID_KEY = ~0ull;

struct PairSrc {                 // pairs already in memory
  static constexpr bool ALWAYS_VALID = true;
  static constexpr bool BLOCKED_GEN = false;
  static constexpr bool IN_MEMORY = true;
  const u64 *keys;
  const u32 *vals;
  __device__ __forceinline__ bool load_key(u64 idx, u64 &k) const
  { k = keys[idx]; return true; }
  __device__ __forceinline__ u32 load_val(u64 idx) const { return vals[idx]; }
};

template <class S, class = void> struct rs_src_in_memory { static constexpr bool value = false; };
template <class S> struct rs_src_in_memory<S, typename std::enable_if<S::IN_MEMORY>::type> { static constexpr bool value = true; };

// ---- histogram of all digits in one read ---------------------------------------
constexpr int RH_NT = 256, RH_IPT = 16, RH_TILE = RH_NT * RH_IPT;

// SKEWED: whole warps tend to fall into one bin (keys that arrive sorted by their high
// digits, as in the refinement rounds): one atomic per warp then instead of 32 on one
// shared-memory address
template <class Src, bool SKEWED>
__global__ void __launch_bounds__(RH_NT)
rs_hist_kernel(Src src, u64 N, PassPlan plan, unsigned long long *__restrict__ ghist)
{
  __shared__ u32 s_h[RS_MAXPASS * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAXPASS * RS_BINS; i += RH_NT) s_h[i] = 0;
  __syncthreads();
  const u64 ntiles = (N + RH_TILE - 1) / RH_TILE;
  const unsigned lane = lane_id();
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const u64 base = tile * RH_TILE;
#pragma unroll 4
    for (int k = 0; k < RH_IPT; k++) {
      const u64 idx = base + (u64) k * RH_NT + threadIdx.x;
      u64 key = 0;
      bool ok = idx < N;
      if (ok) ok = src.load_key(idx, key);
      if (SKEWED) {
#pragma unroll
        for (int p = 0; p < RS_MAXPASS; p++) {    // (fully unrolled: the plan stays in registers)
          if (p < plan.npass) {
            const unsigned d = (unsigned) (key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
            int pred;
            __match_all_sync(FULL_MASK, ok ? d : 0x1ffu, &pred);
            if (pred) {                       // whole warp in one bin: one atomic
              if (lane == 0 && ok) atomicAdd(&s_h[p * RS_BINS + d], 32u);
            } else if (ok) {
              atomicAdd(&s_h[p * RS_BINS + d], 1u);
            }
          }
        }
      } else if (ok) {
#pragma unroll
        for (int p = 0; p < RS_MAXPASS; p++)
          if (p < plan.npass)
            atomicAdd(&s_h[p * RS_BINS + ((unsigned) (key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RH_NT)
    if (s_h[i]) atomicAdd(&ghist[i], (unsigned long long) s_h[i]);
}

// launch of the histogram kernel: a source may bring a cheaper way to walk its keys
template <class Src>
struct RsHistLauncher {
  static void launch(const Src &src, u64 nsrc, const PassPlan &plan, unsigned long long *ghist, cudaStream_t st)
  {
    u64 tiles = div_up(nsrc, RH_TILE);
    unsigned grid = (unsigned) (tiles < 148ull * 8 ? tiles : 148ull * 8);
    rs_hist_kernel<Src, Src::ALWAYS_VALID><<<grid, RH_NT, 0, st>>>(src, nsrc, plan, ghist);   // pairs in memory may be skewed
  }
};

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

// lanes of the warp whose 8-bit digit equals mine: one ballot per bit; lanes with
// valid == false match nobody
template <bool CHECK_VALID>
__device__ __forceinline__ unsigned warp_peers(unsigned d, bool valid)
{
  unsigned peers;
  if (CHECK_VALID) peers = __ballot_sync(FULL_MASK, valid);
  else peers = FULL_MASK;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    unsigned m;
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "@!p not.b32 %0, %0;\n\t}"
        : "=r"(m) : "r"(d), "r"(1u << b));
    peers &= m;
  }
  return CHECK_VALID ? (valid ? peers : 0u) : peers;
}

// first keys of up to 64 key ranges (MODE 3: the "digit" of a key is the range that owns it).
// The bounds live in a device buffer of the RadixWork (per handle: two handles partitioning
// at the same time on one GPU do not share them) and are staged in shared memory by the CTA.
constexpr int RS_MAX_OWNERS = 64;

__device__ __forceinline__ unsigned rs_owner(u64 key, const u64 *bnd, int nb)
{
  unsigned d = 0;
  for (int i = 1; i < nb; i++) d += key >= bnd[i] ? 1u : 0u;
  return d;
}

// digit extraction.  MODE 0 / 1: the digit is byte `bsel` of the low / high half of the
// key (one PRMT); MODE 2: any (shift, mask); MODE 3: the owning key range
struct RsOwners {               // MODE 3 only
  const u64 *bnd;               // first keys of the ranges (shared memory inside the pass)
  int nb;
  const u64 *binbase;           // null, or per range the address its values are stored to (peer memory)
};

template <int MODE>
__device__ __forceinline__ unsigned rs_digit(u64 key, unsigned bsel, unsigned dmask, const RsOwners &ow)
{
  if (MODE == 0) return __byte_perm((u32) key, 0u, bsel);
  if (MODE == 1) return __byte_perm((u32) (key >> 32), 0u, bsel);
  if (MODE == 3) return rs_owner(key, ow.bnd, ow.nb);
  return (unsigned) (key >> bsel) & dmask;
}

// optional phase timing (tools/rs_bench.cu): cycles per phase summed over all tiles
#ifdef GTB_RS_PROFILE
__device__ unsigned long long g_rs_phase[8];
#define RS_PHASE(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); \
    atomicAdd(&g_rs_phase[i], (unsigned long long) (now_ - t_phase_)); t_phase_ = now_; } } while (0)
#else
#define RS_PHASE(i) do { } while (0)
#endif

__device__ __forceinline__ bool rs_status_ready(u64 s, u32 epoch)
{
  return (s >> 48) == (u64) epoch && (s & (3ull << 46)) != 0;
}

// status words are self-validating (tag + payload in one 64-bit store).  Prefetches and
// first polls use weak L1-bypassing loads, which pipeline; a poll that keeps failing
// falls back to the strong form so that progress never depends on L1 behaviour.
#ifdef GTB_RS_STRONG
#define RS_LD(p) ld_relaxed_u64(p)
#define RS_ST(p, v) st_relaxed_u64(p, v)
#else
#define RS_LD(p) ld_na_u64(p)
#define RS_ST(p, v) st_weak_u64(p, v)
#endif

// ---- one onesweep pass ---------------------------------------------------------
// FULL: the tile is complete and every item takes part: no guards at all.
// rs_tile = the load of a tile (warp-striped, or generated and transposed) + rs_tile_sort.
template <class Src, class Cfg, int MODE, bool FULL>
__device__ __forceinline__ void
rs_tile_sort(const Src &src, u64 (&key)[Cfg::IPT], u32 (&val)[Cfg::VAL_EARLY ? Cfg::IPT : 1], unsigned okmask,
             u64 *__restrict__ okeys, u32 *__restrict__ ovals, u64 base, u64 tile, unsigned bsel, unsigned dmask,
             const u64 *__restrict__ gbase, u64 *status, u32 epoch, unsigned char *rs_smem, const RsOwners &ow);

template <class Src, class Cfg, int MODE, bool FULL>
__device__ __forceinline__ void
rs_tile(const Src &src, u64 *__restrict__ okeys, u32 *__restrict__ ovals, u64 base, u32 count,
        u64 tile, unsigned bsel, unsigned dmask, const u64 *__restrict__ gbase,
        u64 *status, u32 epoch, unsigned char *rs_smem, const RsOwners &ow)
{
  constexpr int IPT = Cfg::IPT, TILE = Cfg::TILE;
  u64 *s_keys = reinterpret_cast<u64 *>(rs_smem);
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 wbase = warp * (32u * IPT) + lane;        // index of this thread's row-0 item

  u64 key[IPT];
  u32 val[Cfg::VAL_EARLY ? IPT : 1];
  unsigned okmask = FULL ? ~0u : 0u;

  // warp-striped load: warp w owns [w*32*IPT, (w+1)*32*IPT), lane-contiguous rows
  if constexpr (Src::BLOCKED_GEN) {
    // thread t generates items [t*IPT, (t+1)*IPT) into a padded row (IPT + 1 words: the
    // blocked writes and the striped reads are both free of bank conflicts); the staging
    // area is still unused at this point
    static_assert(!FULL, "generated tiles carry invalid items");
    static_assert(IPT == 16, "gen_block fills rows of 16 items");
    static_assert(!Cfg::VAL_EARLY, "values of a generated tile are computed when they are staged");
    static_assert(sizeof(u64) * (TILE + TILE / IPT) <= (sizeof(u64) + sizeof(u32)) * TILE, "row padding fits the staging area");
    const u32 first = tid * (u32) IPT;
    src.gen_block(base + first, first < count ? (count - first < (u32) IPT ? count - first : (u32) IPT) : 0u,
                  s_keys + first + tid);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const u32 idx = wbase + (u32) k * 32u;
      key[k] = s_keys[idx + idx / (u32) IPT];
      if (key[k] != RS_INVALID_KEY) okmask |= 1u << k;
    }
    __syncthreads();              // separates the reads above from the staging writes
    // (a guard-free variant for tiles without invalid items does not pay: two inlined copies of
    // rs_tile_sort in one kernel also make ptxas 12.9 fail with C7600)
    rs_tile_sort<Src, Cfg, MODE, false>(src, key, val, okmask, okeys, ovals, base, tile, bsel, dmask, gbase, status, epoch, rs_smem, ow);
    return;
  } else {
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const u32 idx = wbase + (u32) k * 32u;
      if (FULL) {
        src.load_key(base + idx, key[k]);
        if (Cfg::VAL_EARLY) val[k] = src.load_val(base + idx);
      } else {
        bool ok = idx < count;
        key[k] = 0;
        if (ok) ok = src.load_key(base + idx, key[k]);
        if (Cfg::VAL_EARLY) val[k] = idx < count ? src.load_val(base + idx) : 0u;
        okmask |= (ok ? 1u : 0u) << k;
      }
    }
    rs_tile_sort<Src, Cfg, MODE, FULL>(src, key, val, okmask, okeys, ovals, base, tile, bsel, dmask, gbase, status, epoch, rs_smem, ow);
  }
}

template <class Src, class Cfg, int MODE, bool FULL>
__device__ __forceinline__ void
rs_tile_sort(const Src &src, u64 (&key)[Cfg::IPT], u32 (&val)[Cfg::VAL_EARLY ? Cfg::IPT : 1], unsigned okmask,
             u64 *__restrict__ okeys, u32 *__restrict__ ovals, u64 base, u64 tile, unsigned bsel, unsigned dmask,
             const u64 *__restrict__ gbase, u64 *status, u32 epoch, unsigned char *rs_smem, const RsOwners &ow)
{
  constexpr int NT = Cfg::NT, IPT = Cfg::IPT, TILE = Cfg::TILE, WARPS = Cfg::WARPS;
  u64 *s_keys = reinterpret_cast<u64 *>(rs_smem);
  u32 *s_vals = reinterpret_cast<u32 *>(s_keys + TILE);
  typedef unsigned short hist_t;
  hist_t *s_wh = reinterpret_cast<hist_t *>(s_vals + TILE);   // [WARPS][RS_BINS] warp histograms -> slots
  u64 **s_pk  = reinterpret_cast<u64 **>(s_wh);       // [RS_BINS] per-bin output pointers, reuse
  u32 **s_pv  = reinterpret_cast<u32 **>(s_pk + RS_BINS);   // s_wh after staging
  u32 *s_scan = reinterpret_cast<u32 *>(reinterpret_cast<unsigned char *>(s_wh) + Cfg::HIST);   // WARPS + 2
  static_assert(TILE <= 65535, "16-bit slots");
#ifdef GTB_RS_PROFILE
  long long t_phase_ = clock64();
#endif

  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u64 ep = (u64) epoch << 48;
  const u32 wbase = warp * (32u * IPT) + lane;        // index of this thread's row-0 item
  u32 rk[IPT];                  // warp-local rank, later the staging slot

  // stable ranking inside the warp (rows in order, lanes in order): every lane reads the
  // warp's bin counter, the first lane of each set of equal digits writes it back
  // bumped by the size of the set
  hist_t *wh = s_wh + warp * RS_BINS;
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    const bool ok = FULL || ((okmask >> k) & 1u);
    const unsigned d = rs_digit<MODE>(key[k], bsel, dmask, ow);
    const unsigned peers = warp_peers<!FULL>(d, ok);
    const unsigned lower = peers & lt;
    const u32 old = wh[d];
    __syncwarp();
    if (ok && lower == 0u) wh[d] = (hist_t) (old + (u32) __popc(peers));
    __syncwarp();
    rk[k] = old + (u32) __popc(lower);
  }
  __syncthreads();
  RS_PHASE(0);                  // load + ranking

  // one thread per bin: warp offsets, tile-local bin starts, publish the tile aggregate
  u32 cnt = 0, excl = 0;
  {
    if (tid < RS_BINS) {
#pragma unroll
      for (int w = 0; w < WARPS; w++) cnt += s_wh[w * RS_BINS + tid];
    }
    u32 total;
    excl = block_exclusive_sum<NT, u32>(cnt, s_scan, &total);
    if (!FULL && tid == 0) s_scan[WARPS + 1] = total;  // slot not used by the scan any more
    if (tid < RS_BINS) {
      u32 run = excl;
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const u32 c = s_wh[w * RS_BINS + tid];
        s_wh[w * RS_BINS + tid] = (hist_t) run;
        run += c;
      }
      RS_ST(status + tile * RS_BINS + tid, ep | (tile == 0 ? RS_FLAG_INCL : RS_FLAG_AGG) | (u64) cnt);
    }
  }
  __syncthreads();
  RS_PHASE(1);                  // bin scan + publish
  const u32 total = FULL ? (u32) TILE : s_scan[WARPS + 1];   // staged (= valid) items

  // stage the tile in digit order
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    if (FULL || ((okmask >> k) & 1u)) {
      const unsigned d = rs_digit<MODE>(key[k], bsel, dmask, ow);
      rk[k] += wh[d];
      s_keys[rk[k]] = key[k];
    }
  }
  if (Cfg::VAL_EARLY) {
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (FULL || ((okmask >> k) & 1u)) s_vals[rk[k]] = val[k];
  } else {
    u32 v[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const u32 idx = wbase + (u32) k * 32u;
      v[k] = (FULL || ((okmask >> k) & 1u)) ? src.load_val(base + idx) : 0u;
    }
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (FULL || ((okmask >> k) & 1u)) s_vals[rk[k]] = v[k];
  }
  __syncthreads();
  RS_PHASE(2);                  // staging

  // decoupled look-back, LB predecessors per round trip (the warp histograms are dead:
  // the pointer tables reuse their memory)
  if (tid < RS_BINS) {
    u64 prefix = gbase[tid];
    if (Cfg::LB > 0 && tile != 0) {              // (LB = 0: timing experiment, wrong output)
      constexpr int LB = Cfg::LB > 0 ? Cfg::LB : 1;
      u64 sum = 0;
      u64 t = tile;                              // predecessors t-1, t-2, ... are still to visit
      bool done = false;
      while (!done) {
        u64 s[LB];
#pragma unroll
        for (int i = 0; i < LB; i++)
          s[i] = t > (u64) i ? RS_LD(status + (t - 1 - i) * RS_BINS + tid) : 0ull;
#pragma unroll
        for (int i = 0; i < LB; i++) {
          if (!done) {
            u64 v = s[i];
#ifdef GTB_RS_PROFILE
            if (tid == 0) atomicAdd(&g_rs_phase[5], 1ull);
#endif
            while (!rs_status_ready(v, epoch)) {
              v = ld_relaxed_u64(status + (t - 1 - i) * RS_BINS + tid);
#ifdef GTB_RS_PROFILE
              if (tid == 0) atomicAdd(&g_rs_phase[6], 1ull);
#endif
            }
            sum += v & RS_VALUE_MASK;
            if (v & RS_FLAG_INCL) done = true;   // tile 0 is always inclusive: never walks past it
          }
        }
        t -= LB;
      }
      RS_ST(status + tile * RS_BINS + tid, ep | RS_FLAG_INCL | ((sum + cnt) & RS_VALUE_MASK));
      prefix += sum;
    }
    const u64 adj = prefix - (u64) excl;   // output index of slot 0 of this bin
    s_pk[tid] = okeys + adj;
    if (MODE == 3 && ow.binbase != nullptr)    // partition straight into the owners' buffers (peer memory): gbase is
      s_pv[tid] = reinterpret_cast<u32 *>(ow.binbase[tid]) + adj;   // zero, the prefix counts inside the bin
    else
      s_pv[tid] = ovals + adj;
  }
  __syncthreads();
  RS_PHASE(3);                  // look-back

  // coalesced scatter: consecutive threads -> consecutive slots -> runs per bin
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    const u32 i = tid + (u32) k * NT;
    if (FULL || i < total) {
      const u64 kk = s_keys[i];
      const unsigned d = rs_digit<MODE>(kk, bsel, dmask, ow);
      if (MODE != 3 || okeys != nullptr) s_pk[d][i] = kk;     // (a partition pass may keep the values only)
      s_pv[d][i] = s_vals[i];
    }
  }
  RS_PHASE(4);                  // scatter (issue only: the stores drain asynchronously)
}

template <class Src, class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB)
rs_onesweep_kernel(Src src, u64 *__restrict__ okeys, u32 *__restrict__ ovals, u64 N,
                   unsigned bsel, unsigned dmask, const u64 *__restrict__ gbase,
                   u64 *status, u32 epoch, u32 *ticket, u32 ticket_base,
                   const u64 *__restrict__ owner_bounds, int nowners, const u64 *__restrict__ binbase,
                   u32 tile_offset)
{
  constexpr int TILE = Cfg::TILE;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  __shared__ u32 s_ticket;
  __shared__ u64 s_bnd[MODE == 3 ? RS_MAX_OWNERS : 1];
  if (MODE == 3 && threadIdx.x < (unsigned) nowners) s_bnd[threadIdx.x] = owner_bounds[threadIdx.x];
  const RsOwners ow{s_bnd, nowners, binbase};
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u) - ticket_base;
  {
    u32 *s_wh = reinterpret_cast<u32 *>(rs_smem + (sizeof(u64) + sizeof(u32)) * TILE);
    for (int i = threadIdx.x; i < (int) (Cfg::HIST / 4); i += Cfg::NT) s_wh[i] = 0;
  }
  __syncthreads();
  // a pass may be fed by two launches (a second source appended to the first, radix_sort): the tile
  // number continues (look-back across both), the items are counted from the launch's own source
  const u64 tile = s_ticket;
  const u64 base = (tile - tile_offset) * (u64) TILE;
  if constexpr (Cfg::PF > 0 && rs_src_in_memory<Src>::value) {
    const u64 pbase = base + (u64) Cfg::PF * TILE;
    if (pbase + TILE <= N) {
      const char *pk = reinterpret_cast<const char *>(src.keys + pbase) + 128u * threadIdx.x;
      if (128u * threadIdx.x < sizeof(u64) * TILE) asm volatile("prefetch.global.L2 [%0];" :: "l"(pk));
      const char *pv = reinterpret_cast<const char *>(src.vals + pbase) + 128u * threadIdx.x;
      if (128u * threadIdx.x < sizeof(u32) * TILE) asm volatile("prefetch.global.L2 [%0];" :: "l"(pv));
    }
  }
  const u32 count = (N - base) < (u64) TILE ? (u32) (N - base) : (u32) TILE;
  if constexpr (Src::ALWAYS_VALID) {
    if (count == (u32) TILE) {
      rs_tile<Src, Cfg, MODE, true>(src, okeys, ovals, base, count, tile, bsel, dmask, gbase, status, epoch, rs_smem, ow);
      return;
    }
  }
  rs_tile<Src, Cfg, MODE, false>(src, okeys, ovals, base, count, tile, bsel, dmask, gbase, status, epoch, rs_smem, ow);
}

// ---- host-side driver ------------------------------------------------------------
struct RadixWork {
  u64 *status = nullptr;            // [status_tiles][256]
  u64  status_tiles = 0;
  u32 *ticket = nullptr;
  u32  ticket_base = 0;
  u32  pending_tiles = 0;             // tiles of the first launch of a pass whose second launch is still to come
  u32  epoch = 0;
  unsigned long long *ghist = nullptr;   // [8][256] device
  u64 *gbase = nullptr;                  // [8][256] device
  unsigned long long *h_hist = nullptr;  // pinned host copy
  u64 *bounds = nullptr;                 // [RS_MAX_OWNERS] first keys of the owning ranges (partition passes)
  int nbounds = 0;
  u64 *binbase = nullptr;                // [RS_BINS] per range the address its values go to (peer stores)
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
  GTB_CUDA(cudaMalloc(&w.bounds, sizeof(u64) * RS_MAX_OWNERS));
  GTB_CUDA(cudaMalloc(&w.binbase, sizeof(u64) * RS_BINS));
  GTB_CUDA(cudaMemset(w.binbase, 0, sizeof(u64) * RS_BINS));
  for (int i = 0; i < 4; i++) GTB_CUDA(cudaEventCreate(&w.ev[i]));
  return 0;
}

static inline void radix_work_free(RadixWork &w)
{
  cudaFree(w.status); cudaFree(w.ticket); cudaFree(w.ghist); cudaFree(w.gbase);
  cudaFree(w.bounds); cudaFree(w.binbase);
  if (w.h_hist) cudaFreeHost(w.h_hist);
  for (int i = 0; i < 4; i++) if (w.ev[i]) cudaEventDestroy(w.ev[i]);
  w = RadixWork();
}

// the status array is sized for the smallest tile any configuration uses (2048)
constexpr int RS_MIN_TILE = 2048;

static inline int radix_work_reserve(RadixWork &w, u64 nitems, ErrBuf &err)
{
  const u64 tiles = div_up(nitems, RS_MIN_TILE) + 1;
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

template <class Src, class Cfg, int MODE>
static int rs_launch_mode(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc, u64 tiles,
                          u64 *okeys, u32 *ovals, unsigned bsel, unsigned dmask, int passidx,
                          ErrBuf &err, bool peer_bins = false, u32 tile_offset = 0)
{
  // (the attribute is per device, instantiation and context: gtb_release_devices destroys contexts)
  static std::atomic<unsigned> attr_set[64];
  int dev = 0;
  GTB_CUDA(cudaGetDevice(&dev));
  const unsigned generation = g_context_generation[dev & 63].load(std::memory_order_relaxed) + 1;
  if (attr_set[dev & 63].load(std::memory_order_relaxed) != generation) {
    GTB_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<Src, Cfg, MODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int) Cfg::SMEM));
    attr_set[dev & 63].store(generation, std::memory_order_relaxed);
  }
  rs_onesweep_kernel<Src, Cfg, MODE><<<(unsigned) tiles, Cfg::NT, Cfg::SMEM, st>>>(
      src, okeys, ovals, nsrc, bsel, dmask, w.gbase + passidx * RS_BINS,
      w.status, w.epoch, w.ticket, w.ticket_base, w.bounds, MODE == 3 ? w.nbounds : 0,
      peer_bins ? w.binbase : nullptr, tile_offset);
  GTB_LAUNCH_CHECK();
  return 0;
}

// tiles_before > 0: this launch continues the pass a preceding launch (same pass, another source)
// began: same epoch, the tile numbers go on
template <class Src, class Cfg>
static int rs_launch_pass(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc,
                          u64 *okeys, u32 *ovals, int shift, int bits, bool padded, int passidx,
                          ErrBuf &err, u64 tiles_before = 0, bool more_follows = false)
{
  static_assert(Cfg::TILE >= RS_MIN_TILE, "status array sizing");
  const u64 tiles = div_up(nsrc, Cfg::TILE);
  if (tiles == 0) return 0;
  if (tiles_before + tiles > w.status_tiles) { err.set("radix: status array too small"); return -1; }
  if (tiles_before == 0 && ++w.epoch >= 0xffffu) {           // epoch space exhausted: start over
    GTB_CUDA(cudaMemsetAsync(w.status, 0, sizeof(u64) * RS_BINS * w.status_tiles, st));
    w.epoch = 1;
  }
  // a digit that is a whole byte of one key half (or a shorter one whose byte is zero
  // padded) is extracted with one PRMT
  const unsigned dmask = (1u << bits) - 1u;
  if (shift % 8 == 0 && (bits == 8 || padded)) {
    const unsigned bsel = 0x4440u + (unsigned) (shift % 32) / 8u;
    if (shift < 32) GTB_TRY((rs_launch_mode<Src, Cfg, 0>(w, st, src, nsrc, tiles, okeys, ovals, bsel, dmask, passidx, err, false, (u32) tiles_before)));
    else GTB_TRY((rs_launch_mode<Src, Cfg, 1>(w, st, src, nsrc, tiles, okeys, ovals, bsel, dmask, passidx, err, false, (u32) tiles_before)));
  } else {
    GTB_TRY((rs_launch_mode<Src, Cfg, 2>(w, st, src, nsrc, tiles, okeys, ovals, (unsigned) shift, dmask, passidx, err, false, (u32) tiles_before)));
  }
  // (the ticket counter runs on through both launches of a pass; its base moves when the pass is complete)
  w.pending_tiles = more_follows ? (u32) (tiles_before + tiles) : 0u;
  if (!more_follows) w.ticket_base += (u32) (tiles_before + tiles);
  if (tiles_before == 0) w.passes++;
  w.launches++;
  return 0;
}

// number of items of `src` per owning key range
template <class Src>
__global__ void __launch_bounds__(RH_NT)
rs_owner_hist_kernel(Src src, u64 N, unsigned long long *__restrict__ ghist,
                     const u64 *__restrict__ owner_bounds, int nb)
{
  __shared__ u32 s_h[RS_BINS];
  __shared__ u64 s_bnd[RS_MAX_OWNERS];
  s_h[threadIdx.x] = 0;
  if (threadIdx.x < (unsigned) nb) s_bnd[threadIdx.x] = owner_bounds[threadIdx.x];
  __syncthreads();
  if constexpr (Src::BLOCKED_GEN) {
    // a thread takes 16 consecutive items; with at most 16 ranges their counts fit two registers of
    // eight 8-bit fields, summed over the warp before they touch shared memory
    const u64 nchunks = (N + 15) >> 4;
    for (u64 c0 = (u64) blockIdx.x * RH_NT; c0 < nchunks; c0 += (u64) gridDim.x * RH_NT) {
      const u64 c = c0 + threadIdx.x;
      const u64 idx = c << 4;
      const u32 cnt = c < nchunks ? (N - idx < 16 ? (u32) (N - idx) : 16u) : 0u;
      unsigned long long packed[2] = {0ull, 0ull};
      if (cnt > 0 && src.can_gen_fast(idx, cnt) && nb <= 16) {
        u64 keys[16];
        src.gen_block_fast(idx, keys);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (keys[i] == RS_INVALID_KEY) continue;
          const unsigned d = rs_owner(keys[i], s_bnd, nb);
          if (d < 8u) packed[0] += 1ull << (8u * d); else packed[1] += 1ull << (8u * (d - 8u));
        }
      } else {
        for (u32 i = 0; i < cnt; i++) {
          u64 key;
          if (src.load_key(idx + i, key)) atomicAdd(&s_h[rs_owner(key, s_bnd, nb)], 1u);
        }
      }
      for (int g = 0; g < nb && g < 16; g++) {               // (uniform trip count)
        const u32 mine = (u32) ((g < 8 ? packed[0] : packed[1]) >> (8 * (g & 7))) & 0xffu;
        const u32 sum = __reduce_add_sync(FULL_MASK, mine);
        if (lane_id() == 0 && sum) atomicAdd(&s_h[g], sum);
      }
    }
  } else {
    const u64 ntiles = (N + RH_TILE - 1) / RH_TILE;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const u64 base = tile * RH_TILE;
#pragma unroll 4
      for (int k = 0; k < RH_IPT; k++) {
        const u64 idx = base + (u64) k * RH_NT + threadIdx.x;
        u64 key = 0;
        bool ok = idx < N;
        if (ok) ok = src.load_key(idx, key);
        const unsigned d = rs_owner(key, s_bnd, nb);
        // consecutive text positions mostly differ in their owner: aggregate what does coincide
        const unsigned act = __ballot_sync(FULL_MASK, ok);
        if (ok) {
          const unsigned peers = __match_any_sync(act, d);
          if ((int) lane_id() == __ffs(peers) - 1) atomicAdd(&s_h[d], (u32) __popc(peers));
        }
      }
    }
  }
  __syncthreads();
  if (s_h[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], (unsigned long long) s_h[threadIdx.x]);
}

// A stable partition of the valid items of `src` by owning key range, in two steps so that a
// caller can learn the group sizes of ALL slices (an all-gather) before anything is stored:
//   rs_owner_counts   counts_out[r] = items of range r (bounds[0..nranges) = first key of each
//                     range, bounds[0] is taken as 0); synchronises the stream
//   rs_owner_scatter  one onesweep pass whose digit is the owner.  binbase_host == null: the groups
//                     are written one after the other to okeys/ovals (okeys may be null: values
//                     only).  binbase_host[r] != 0: the values of range r are stored, in source
//                     order, from that device address on -- a buffer of the OWNER, mapped into this
//                     GPU's address space (peer access, or a mapping of the owner's allocation): the exchange of the partitioned
//                     positions is fused into the pass, the stores travel over NVLink.
template <class Src>
static int rs_owner_counts(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc, const u64 *bounds,
                           int nranges, u64 *counts_out, ErrBuf &err)
{
  if (nranges < 1 || nranges > RS_MAX_OWNERS) { err.set("partition: 1..%d ranges", RS_MAX_OWNERS); return -1; }
  for (int r = 0; r < nranges; r++) counts_out[r] = 0;
  u64 hb[RS_MAX_OWNERS];
  for (int r = 0; r < nranges; r++) hb[r] = r == 0 ? 0 : bounds[r];
  GTB_CUDA(cudaMemcpyAsync(w.bounds, hb, sizeof(u64) * nranges, cudaMemcpyHostToDevice, st));
  GTB_CUDA(cudaStreamSynchronize(st));            // (hb lives on this stack frame)
  w.nbounds = nranges;
  if (nsrc == 0) return 0;
  GTB_TRY(radix_work_reserve(w, nsrc, err));
  GTB_CUDA(cudaMemsetAsync(w.ghist, 0, sizeof(unsigned long long) * RS_BINS, st));
  u64 tiles = div_up(nsrc, RH_TILE);     // (a blocked source takes 16 items per thread: the same 4096 per CTA)
  unsigned grid = (unsigned) (tiles < 148ull * 8 ? tiles : 148ull * 8);
  rs_owner_hist_kernel<Src><<<grid, RH_NT, 0, st>>>(src, nsrc, w.ghist, w.bounds, nranges);
  GTB_LAUNCH_CHECK();
  w.launches += 1;
  GTB_CUDA(cudaMemcpyAsync(w.h_hist, w.ghist, sizeof(unsigned long long) * RS_BINS, cudaMemcpyDeviceToHost, st));
  GTB_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < nranges; r++) counts_out[r] = w.h_hist[r];
  return 0;
}

template <class Src, class Cfg = RsDefault>
static int rs_owner_scatter(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc, int nranges, u64 total,
                            u64 *okeys, u32 *ovals, const u64 *binbase_host, ErrBuf &err)
{
  if (nsrc == 0 || total == 0) return 0;
  if (binbase_host) {
    u64 bb[RS_BINS];
    for (int r = 0; r < RS_BINS; r++) bb[r] = r < nranges ? binbase_host[r] : 0ull;
    GTB_CUDA(cudaMemcpyAsync(w.binbase, bb, sizeof bb, cudaMemcpyHostToDevice, st));
    GTB_CUDA(cudaMemsetAsync(w.gbase, 0, sizeof(u64) * RS_BINS, st));     // offsets count inside each bin
    GTB_CUDA(cudaStreamSynchronize(st));          // (bb lives on this stack frame)
  } else {
    rs_scan_kernel<<<1, RS_BINS, 0, st>>>(w.ghist, w.gbase);
    GTB_LAUNCH_CHECK();
    w.launches += 1;
  }
  const u64 tiles = div_up(nsrc, Cfg::TILE);
  if (tiles > w.status_tiles) { err.set("radix: status array too small"); return -1; }
  if (++w.epoch >= 0xffffu) {
    GTB_CUDA(cudaMemsetAsync(w.status, 0, sizeof(u64) * RS_BINS * w.status_tiles, st));
    w.epoch = 1;
  }
  GTB_TRY((rs_launch_mode<Src, Cfg, 3>(w, st, src, nsrc, tiles, okeys, ovals, 0u, 0xffu, 0, err, binbase_host != nullptr)));
  w.ticket_base += (u32) tiles;
  w.passes++; w.launches++;
  w.pairs_moved += total;
  return 0;
}

// both steps for one caller-owned buffer.  Synchronises the stream.
template <class Src, class Cfg = RsDefault>
static int rs_partition_by_owner(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc, const u64 *bounds,
                                 int nranges, u64 *okeys, u32 *ovals, u64 capacity, u64 *counts_out, ErrBuf &err)
{
  GTB_TRY(rs_owner_counts(w, st, src, nsrc, bounds, nranges, counts_out, err));
  u64 total = 0;
  for (int r = 0; r < nranges; r++) total += counts_out[r];
  if (total > capacity) { err.set("partition: %llu items for a buffer of %llu", (unsigned long long) total, (unsigned long long) capacity); return -1; }
  GTB_TRY((rs_owner_scatter<Src, Cfg>(w, st, src, nsrc, nranges, total, okeys, ovals, nullptr, err)));
  GTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// Sort pairs by the digits of `plan` (LSD, stable).  The first executed pass reads
// from `src` (nsrc items, some possibly invalid); the sorted pairs end up in
// kbuf[*res]/vbuf[*res], *nout = number of valid pairs.  Passes whose digit is the
// same for every key are skipped.  Synchronises the stream once (histogram
// read-back).
//
// extra (optional): a second source whose items count as FOLLOWING all items of `src` in input order
// (the first pass is fed by two launches).  This is how the first-level suffix sort drops the pass over
// the tail field of its keys: the few keys that carry a tail are taken out of the text scan, sorted by
// their tail and appended -- a stable LSD sort of "full keys in text order, then tail keys in
// (tail, text) order" by the symbol digits alone ends exactly where the sort with the tail digit ends.
template <class Src, class Extra = PairSrc, class Cfg = RsDefault>
static int radix_sort(RadixWork &w, cudaStream_t st, const Src &src, u64 nsrc,
                      u64 *kbuf[2], u32 *vbuf[2], const PassPlan &plan,
                      int *res, u64 *nout, ErrBuf &err, const Extra *extra = nullptr, u64 nextra = 0)
{
  *res = 0; *nout = 0;
  if (!extra) nextra = 0;
  if (nsrc == 0 && nextra == 0) return 0;
  if (plan.npass < 1 || plan.npass > RS_MAXPASS) { err.set("radix: bad pass plan"); return -1; }
  GTB_TRY(radix_work_reserve(w, nsrc + nextra + 2 * RS_MIN_TILE, err));
  GTB_CUDA(cudaEventRecord(w.ev[0], st));
  GTB_CUDA(cudaMemsetAsync(w.ghist, 0, sizeof(unsigned long long) * RS_MAXPASS * RS_BINS, st));
  {
    if (nsrc > 0) {
      RsHistLauncher<Src>::launch(src, nsrc, plan, w.ghist, st);
      GTB_LAUNCH_CHECK();
      w.launches++;
    }
    if (nextra > 0) {
      RsHistLauncher<Extra>::launch(*extra, nextra, plan, w.ghist, st);
      GTB_LAUNCH_CHECK();
      w.launches++;
    }
    rs_scan_kernel<<<plan.npass, RS_BINS, 0, st>>>(w.ghist, w.gbase);
    GTB_LAUNCH_CHECK();
    w.launches++;
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
      const u64 tiles_a = div_up(nsrc, Cfg::TILE);
      if (nsrc > 0)
        GTB_TRY((rs_launch_pass<Src, Cfg>(w, st, src, nsrc, kbuf[0], vbuf[0], plan.shift[p], plan.bits[p], plan.padded, p, err, 0, nextra > 0)));
      if (nextra > 0)
        GTB_TRY((rs_launch_pass<Extra, Cfg>(w, st, *extra, nextra, kbuf[0], vbuf[0], plan.shift[p], plan.bits[p], plan.padded, p, err,
                                            tiles_a, false)));
      cur = 0;
    } else {
      PairSrc ps{kbuf[cur], vbuf[cur]};
      GTB_TRY((rs_launch_pass<PairSrc, Cfg>(w, st, ps, total, kbuf[cur ^ 1], vbuf[cur ^ 1], plan.shift[p],
                                            plan.bits[p], plan.padded, p, err)));
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
