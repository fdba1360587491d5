// gtb_esa.cu -- libgtb200.so: host orchestration + C-ABI (include/gtb200.h).
//
// Pipeline of gtb_esa_run (replaces gt_Sfxiterator_new/next + the GtOutlcpinfo side
// channel, /root/reference/src/match/sfx-suffixer.c:1363-2204,
// /root/reference/src/match/sfx-lcpvalues.c):
//   0  special mask in HBM (from the special ranges / from the byte symbols)
//   K1 code histogram (+ special k-mers)            -> leftborder counts, countspecialcodes,
//   K2 exclusive scan                                   distpfxidx  (.bck)
//   K3+K4 fused key generation + onesweep LSD radix sort of (key64, pos32)
//   A  group analysis: lcp of resolved neighbours, compaction of ties, inverse SA
//   K5 prefix doubling rounds on the ties (same radix engine)
//   K6 exact lcp of deep pairs, .llv list, stats
//   K7 special tail
#include <stdarg.h>
#include <stdlib.h>
#include <chrono>
#include <map>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/gtb200.h"
#include "gtb_common.cuh"
#include "gtb_radix.cuh"
#include "gtb_esa_kernels.cuh"
#include "gtb_hostio.cuh"
#include "gtb_vmm.cuh"
#include "gtb_shard.cuh"

namespace gtb {

void ErrBuf::set(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
}

// a device buffer that only grows (re-runs on one handle do not reallocate)
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool borrowed = false;     // memory owned by another handle (gtb_esa_share_input)
  // share_dev >= 0: the buffer is allocated with the virtual-memory API on that device so that the
  // other processes of a sharded job can map it (gtb_vmm.cuh); alloc_id names the allocation
  int share_dev = -1;
  bool vmm = false;
  CUmemGenericAllocationHandle mh = 0;
  u64 alloc_id = 0;
  int ensure(size_t bytes, ErrBuf &err)
  {
    if (bytes <= cap && !borrowed && (share_dev < 0 || vmm)) return 0;
    release();
    size_t want = bytes + (bytes >> 6) + 256;
    if (share_dev >= 0) {
      static std::atomic<unsigned long long> next_id{0};
      size_t got = 0;
      if (vmm_alloc(share_dev, want, &p, &mh, &got, err) != 0) { p = nullptr; return -1; }
      cap = got; vmm = true; alloc_id = ++next_id;
      return 0;
    }
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      err.set("cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
      p = nullptr; return -1;
    }
    cap = want;
    return 0;
  }
  void release()
  {
    if (p && !borrowed) { if (vmm) vmm_free(p, mh, cap); else cudaFree(p); }
    p = nullptr; cap = 0; borrowed = false; vmm = false; mh = 0; alloc_id = 0;
  }
  void borrow(const DevBuf &o) { release(); p = o.p; cap = o.cap; borrowed = o.p != nullptr; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

static u64 ipow_u64(u64 b, unsigned e) { u64 r = 1; while (e--) r *= b; return r; }

} // namespace gtb

using namespace gtb;

struct gtb_esa {
  ErrBuf err;
  int device = 0;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;   // second copy stream of gtb_esa_copy_suftab_u64
  cudaStream_t st3 = nullptr;   // lcp bytes beside the suffix table (gtb_esa_copy_tables)
  // input
  bool have_input = false, dna = true;
  int borrowers = 0;             // handles that use this handle's sequence in HBM (gtb_esa_share_input)
  gtb_esa *lender = nullptr;     // ... and the handle this one borrows from
  u64 n = 0, S = 0;
  unsigned K = 4;
  DevBuf words, bytes, spmask, ranges, sepbits, seppos;
  bool have_sep = false;
  u64 nmaskwords = 0;
  unsigned readmode = 0;    // GtReadmode of the sequence in HBM: 0 fwd, 1 rev, 2 cpl, 3 rcl
  // code range (shard)
  bool full_range = true;
  u64 mincode = 0, maxcode = 0;
  int emit_tail = 1;
  // bucket table
  unsigned pl = 0;
  bool counted = false;
  u64 ncodes = 0, nspecialcodes = 0, ndist = 0;
  DevBuf leftborder, csc, dist, distoff;
  // sort state
  DevBuf kbuf[2], vbuf[2], lcp8, hbits, ubits, tbits, tpre, trank, spre, tile_a, tile_b, tile_c, tile_d, scantmp, dstats, misc;
  DevBuf uidx0, ugrp0, uidx[2], ugrp[2], upos[2], dkeys, kd[2], vd[2], ulcp, llvflags, llv;
  RadixWork rw;
  int res = 0;              // which vbuf holds the suffix table
  u64 N = 0;                // sorted (non-special) suffixes of this shard
  u64 sa_offset = 0;        // global SA index of the shard's first entry
  bool range_given = false; // sa_offset / width of the code range come from the caller (coarse counts)
  u64 given_offset = 0, given_width = 0;
  bool lb_own = false;      // leftborder holds (only) this range's own codes, filled from its sorted keys
  DevBuf coarse; unsigned plc = 0; u32 ncoarse = 0;
  u64 entries = 0;          // N (+ S + 1 with tail)
  u64 nllv = 0;
  u64 first_key = 0, last_key = 0;
  bool ran = false;
  // staged execution (multi-range prefix doubling)
  unsigned flags = 0, round = 0;
  u64 M0 = 0, M = 0, atiles = 0, pending_send = 0;
  int cur = 0, bits_lo = 1, bits_hi = 1;
  bool isa_built = false, in_progress = false;
  KeyFmt fmt = byte_fmt();  // key format of this run (DNA: chosen from the text length)
  int opt_key_symbols = 0;  // test knobs (environment GTB200_KEY_SYMBOLS = 16..29,
  int opt_text_rounds = -1; // GTB200_TEXT_ROUNDS = 0..8,
  int opt_tail_last = -1;   // GTB200_TAIL_LAST = 0|1): override the automatic choices
  int opt_pairs_by_text = 1; // GTB200_PAIRS_BY_TEXT = r: tie groups of two are compared by text in the first r doubling rounds
  unsigned isa_round = 0;   // doubling rounds of this run so far (the first one compares pairs by text)
  u64 nspecialranges = 0;   // maximal special runs of the 2-bit input
  DevBuf nearbits, tailkeys[2], tailpos[2];   // the keys with a tail field, sorted apart (stage_begin)
  unsigned text_left = 0;   // text-driven rounds still allowed before ranks are built
  float ms_count_ext = 0;   // device time of gtb_esa_count_partial/_finish since the last run
  float ms_part_ext = 0;    // ... and of gtb_esa_slice_partition (its pass counts as a radix pass)
  u64 part_pairs_ext = 0; u32 part_launches_ext = 0;
  float ext_ms_radix = 0, ext_ms_keygen = 0; u64 ext_pairs = 0; u32 ext_launches = 0;   // carried into this run's stats
  u64 depth[64];            // depth[r] = common prefix of the groups entering round r
  DevBuf ranks, owner, sendidx, rcounts, rankwords;
  // sharded job (gtb_shard_host.cuh): peer mappings, the rank maps of all ranges as this GPU sees them
  DevBuf peertab;
  u64 peer_enabled = 0;     // devices this handle's device has peer access to (bit per device)
  // ranges in other processes (bench.py under torchrun): their buffers mapped here (gtb_vmm.cuh)
  struct PeerImport { u64 alloc_id; void *ptr; size_t size; CUmemGenericAllocationHandle mh; };
  std::map<std::pair<int, int>, PeerImport> imports;      // (rank, slot) -> mapping on this GPU
  struct PendingFd { FdMsg msg; int fd; };
  std::vector<PendingFd> pending_fds;                      // descriptors received before they were asked for
  std::thread ipc_thread;                                  // receives descriptors whenever they arrive: a sender
  std::mutex ipc_mu;                                       // never waits for this range to reach a receive call
  std::condition_variable ipc_cv;
  std::atomic<bool> ipc_stop{false};
  int ipc_sock = -1;
  std::string ipc_key;
  u64 sent_id[8] = {0, 0, 0, 0, 0, 0, 0, 0};              // allocation of each shareable buffer the peers hold
  u64 llv_before = 0;       // .llv pairs of the preceding ranges of the job
  int shard_np = 1;         // code ranges the last sharded run cut
  HostStage hstage;         // pinned staging + host threads of the result copies
  // gtb_esa_run_to_host: the suffix table leaves for the host while the refinement still runs
  struct EarlyCopy {
    uint64_t *dst = nullptr;        // where entries [0, N) go (set before the run; null: no early copy)
    bool started = false;
    std::thread th;
    ChunkDealer *dealer = nullptr;
    ErrBuf err;
    int rc = 0;
    u64 per = 0, count = 0;
  } ec;
  void *patch_host = nullptr;       // pinned: (index : 32 | position : 32) of the entries copied too early
  size_t patch_cap = 0;
  gtb_stats stats;
};

namespace {

struct PhaseTimer {
  gtb_esa *h;
  cudaEvent_t a, b;
  float *acc;
  PhaseTimer(gtb_esa *h_, float *acc_) : h(h_), acc(acc_)
  {
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, h->st);
  }
  void stop()
  {
    cudaEventRecord(b, h->st);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    *acc += ms;
  }
  ~PhaseTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

inline unsigned grid_for(u64 items, unsigned block, unsigned maxblocks = 148u * 16u)
{
  u64 g = div_up(items, block);
  if (g < 1) g = 1;
  return (unsigned) (g < maxblocks ? g : maxblocks);
}

int bitlen(u64 v) { int b = 0; while (v) { b++; v >>= 1; } return b < 1 ? 1 : b; }

// exclusive scan of a u32 array in place (popc_mode: scan popcounts, out may differ)
int device_scan_u32(gtb_esa *h, const u32 *in, u32 *out, u64 count, int popc_mode_tiles_only,
                    u32 **tileoff_out, u64 *total_out)
{
  ErrBuf &err = h->err;
  const u64 tiles = div_up(count, SC_TILE);
  if (tiles == 0) { if (total_out) *total_out = 0; return 0; }
  GTB_TRY(h->scantmp.ensure(sizeof(u32) * tiles + 64, err));
  GTB_TRY(h->misc.ensure(256, err));
  u32 *tilesum = h->scantmp.as<u32>();
  u64 *d_total = h->misc.as<u64>();
  k_scan_tilesums<<<(unsigned) tiles, SC_NT, 0, h->st>>>(in, count, tilesum, popc_mode_tiles_only);
  GTB_LAUNCH_CHECK();
  k_scan_single<<<1, 1024, 0, h->st>>>(tilesum, tiles, d_total);
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches += 2;
  if (!popc_mode_tiles_only) {
    k_scan_apply<<<(unsigned) tiles, SC_NT, 0, h->st>>>(in, out, count, tilesum);
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
  }
  if (tileoff_out) *tileoff_out = tilesum;
  if (total_out) {
    GTB_CUDA(cudaMemcpyAsync(total_out, d_total, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    GTB_CUDA(cudaStreamSynchronize(h->st));
  }
  return 0;
}

// exclusive sum of a[] and exclusive running maximum of b[] over the tiles, in place; total -> misc[0]
int scan_tiles_sum_max(gtb_esa *h, u32 *a, u32 *b, u64 ntiles)
{
  ErrBuf &err = h->err;
  GTB_TRY(h->misc.ensure(256, err));
  if (ntiles <= 4ull * STM_BLOCK) {
    k_scan_tiles_sum_max<<<1, 1024, 0, h->st>>>(a, b, ntiles, h->misc.as<u64>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
    return 0;
  }
  const u64 nblk = div_up(ntiles, STM_BLOCK);
  GTB_TRY(h->scantmp.ensure(sizeof(u32) * 2 * nblk + 64, err));
  u32 *bs = h->scantmp.as<u32>(), *bm = bs + nblk;
  k_scan_tiles_blocks<<<(unsigned) nblk, 1024, 0, h->st>>>(a, b, ntiles, bs, bm, 0);
  GTB_LAUNCH_CHECK();
  k_scan_tiles_sum_max<<<1, 1024, 0, h->st>>>(bs, bm, nblk, h->misc.as<u64>());
  GTB_LAUNCH_CHECK();
  k_scan_tiles_blocks<<<(unsigned) nblk, 1024, 0, h->st>>>(a, b, ntiles, bs, bm, 1);
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches += 3;
  return 0;
}

template <bool DNA>
TextSrc<DNA> make_src(gtb_esa *h, u64 klo, u64 khi)
{
  TextSrc<DNA> s;
  s.words = h->words.as<u64>();
  s.bytes = h->bytes.as<u8>();
  s.spmask = h->spmask.as<u32>();
  s.klo = klo; s.khi = khi;
  s.pos0 = 0;
  s.f = h->fmt;
  s.skip_near = false;
  s.ranges = h->nspecialranges > 0 ? h->ranges.as<u64>() : nullptr;
  s.nranges = h->nspecialranges;
  s.hist4 = false;
  return s;
}

// key range of the code interval [mincode, maxcode]
void code_range_to_keys(const gtb_esa *h, u64 *klo, u64 *khi)
{
  if (h->full_range || h->pl == 0) { *klo = 0; *khi = ~0ull; return; }
  if (h->dna) {
    const unsigned sh = 64 - 2 * h->pl;
    *klo = h->mincode << sh;
    *khi = (h->maxcode << sh) | ((sh < 64 ? (1ull << sh) : 0ull) - 1ull);
  } else {
    const KeyFmt f = byte_fmt();
    u64 lo = 0, hi = 0, a = h->mincode, b = h->maxcode;
    u64 dl[16], dh[16];
    for (int k = (int) h->pl - 1; k >= 0; k--) { dl[k] = a % h->K; a /= h->K; dh[k] = b % h->K; b /= h->K; }
    for (unsigned k = 0; k < h->pl; k++) {
      // the largest symbol of a code may stand for the filler (31) of a special k-mer
      lo = (lo << f.b) | dl[k];
      hi = (hi << f.b) | (dh[k] == h->K - 1 ? 31ull : dh[k]);
    }
    const unsigned sh = 64 - f.b * h->pl;
    *klo = lo << sh;
    *khi = (hi << sh) | ((1ull << sh) - 1ull);
    // a min code ending in K-1 symbols must not exclude nothing below it: fine, dl kept as is
  }
}

int build_mask(gtb_esa *h, const gtb_range *specials, u64 nranges)
{
  ErrBuf &err = h->err;
  h->nmaskwords = (h->n >> 5) + 4;
  GTB_TRY(h->spmask.ensure(sizeof(u32) * h->nmaskwords, err));
  GTB_CUDA(cudaMemsetAsync(h->spmask.p, 0, sizeof(u32) * h->nmaskwords, h->st));
  k_mask_tail<<<grid_for(h->nmaskwords - (h->n >> 5), 256, 64), 256, 0, h->st>>>(
      h->spmask.as<u32>(), h->n, h->nmaskwords);
  GTB_LAUNCH_CHECK();
  if (h->dna) {
    if (nranges > 0) {
      GTB_TRY(h->ranges.ensure(sizeof(u64) * 2 * nranges, err));
      GTB_CUDA(cudaMemcpyAsync(h->ranges.p, specials, sizeof(u64) * 2 * nranges,
                               cudaMemcpyHostToDevice, h->st));
      k_mask_ranges<<<grid_for(nranges * 32, 256), 256, 0, h->st>>>(
          h->spmask.as<u32>(), h->ranges.as<u64>(), nranges);
      GTB_LAUNCH_CHECK();
    }
  } else if (h->n > 0) {
    k_mask_from_bytes<<<grid_for((h->n + 31) >> 5, 256), 256, 0, h->st>>>(
        h->spmask.as<u32>(), h->bytes.as<u8>(), h->n);
    GTB_LAUNCH_CHECK();
  }
  // S = popcount of the mask over positions < n
  const u64 nw = (h->n + 31) >> 5;
  u64 total = 0;
  if (nw > 0) {
    GTB_TRY(device_scan_u32(h, h->spmask.as<u32>(), nullptr, nw, 1, nullptr, &total));
    if (h->n & 31u) total -= 32 - (h->n & 31u);
  }
  h->S = total;
  return 0;
}

// all = true: the three tables by counting (one atomic per suffix), then the scan.
// all = false: the tables are only zeroed; k_analyze_keys fills all three from the sorted
// keys.
int count_codes(gtb_esa *h, unsigned pl, bool all, u64 first = 0, u64 end = ~0ull, bool finish = true)
{
  ErrBuf &err = h->err;
  if (h->counted && h->pl == pl) return 0;
  h->pl = pl;
  h->ncodes = ipow_u64(h->K, pl);
  h->nspecialcodes = ipow_u64(h->K, pl - 1);
  h->ndist = 0;
  u64 distoff[32] = {0};
  for (unsigned i = 1; i + 2 <= pl; i++) { distoff[i] = h->ndist; h->ndist += ipow_u64(h->K, i); }
  GTB_TRY(h->leftborder.ensure(sizeof(u32) * (h->ncodes + 2), err));
  GTB_TRY(h->csc.ensure(sizeof(u32) * (h->nspecialcodes + 1), err));
  GTB_TRY(h->dist.ensure(sizeof(u32) * (h->ndist + 1), err));
  GTB_TRY(h->distoff.ensure(sizeof(u64) * 32, err));
  GTB_CUDA(cudaMemsetAsync(h->leftborder.p, 0, sizeof(u32) * (h->ncodes + 2), h->st));
  GTB_CUDA(cudaMemsetAsync(h->csc.p, 0, sizeof(u32) * (h->nspecialcodes + 1), h->st));
  GTB_CUDA(cudaMemsetAsync(h->dist.p, 0, sizeof(u32) * (h->ndist + 1), h->st));
  GTB_CUDA(cudaMemcpyAsync(h->distoff.p, distoff, sizeof distoff, cudaMemcpyHostToDevice, h->st));
  if (end > h->n) end = h->n;
  if (h->n > 0 && all && first < end) {
    const unsigned grid = grid_for(end - first, 256, 148u * 8u);
    u32 *lb = h->leftborder.as<u32>(), *cs = h->csc.as<u32>(), *di = h->dist.as<u32>();
    const u64 *dof = h->distoff.as<u64>();
    if (h->dna)
      k_count_codes<true, true><<<grid, 256, 0, h->st>>>(make_src<true>(h, 0, ~0ull), first, end, pl, h->K, lb, cs, di, dof);
    else
      k_count_codes<false, true><<<grid, 256, 0, h->st>>>(make_src<false>(h, 0, ~0ull), first, end, pl, h->K, lb, cs, di, dof);
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
  }
  if (all && finish) {
    // counts -> bucket starts; entry ncodes becomes the number of non-special suffixes
    GTB_TRY(device_scan_u32(h, h->leftborder.as<u32>(), h->leftborder.as<u32>(), h->ncodes + 1, 0,
                            nullptr, nullptr));
    h->counted = true;
  }
  return 0;
}

// Key format of a run.  DNA: when only few suffixes meet a special within a key length (genomes: N
// runs and sequence ends are rare; reads are the opposite) the key length is a multiple of four
// symbols -- 16 / 20 / 24 / 28, the smallest with an expected share of chance ties n / 4^m below 1 % --
// so that the lowest radix digit holds nothing but the tail field: the first-level sort then runs
// over the symbol digits only and the keys with a tail enter it as a second, pre-sorted source
// (stage_begin, radix_sort): c4 sorts with 5 passes instead of 6.  Otherwise 17 / 21 / 25 / 29
// symbols (dna_fmt_for).  Every function that makes keys of a run takes its format from here.
// Is dropping the pass over the tail digit worth it?  It replaces one pass over all N pairs by the
// separate treatment of the keys with a tail: at most (special runs + 1) * (m - 1) of them (byte path: every
// special counts as a run).  Yes when those are at most a sixth of the suffixes.
bool tails_few(const gtb_esa *h, int m)
{
  if (h->opt_tail_last == 0) return false;
  if (h->opt_tail_last == 1) return true;
  const u64 runs = h->dna ? h->nspecialranges : h->S;
  const u64 regular = h->n - h->S;
  return (runs + 1) * (u64) (m - 1) <= regular / 6;
}
// ... and the digit histograms from one table of 4-mers (k_hist_4mer) when the corrections at the run ends
// (m - 4 windows per run and digit) are few against the text
bool hist4_pays(const gtb_esa *h, int m) { return h->dna && (h->nspecialranges + 1) * (u64) m * 8 <= h->n / 16; }

KeyFmt choose_fmt(const gtb_esa *h, unsigned pl)
{
  const int k = h->opt_key_symbols;
  if (!h->dna) {
    // byte path, 5 bits per symbol.  Tail-last formats: the boundary between the tail field and the symbols
    // lies on a byte (m = 8: 40 symbol bits; m = 9: 45 symbol bits + 3 zero bits), so the symbols take 5 / 6
    // passes and the tail digit none; otherwise m = 8 / 10 / 12 with the tail directly below the symbols
    const bool forced = (k == 8 || k == 9 || k == 10 || k == 12) && k >= (int) pl;
    if (forced && !(k <= 9 && tails_few(h, k))) { if (k != 9) return make_fmt(k, 5, 4); }
    const int cand[2] = {8, 9};
    for (int i = 0; i < 2; i++) {
      const int m = forced ? k : cand[i];
      if (m > 9 || (unsigned) m < pl || !tails_few(h, m)) { if (forced) break; continue; }
      double km = 1.0;
      for (int j = 0; j < m; j++) km *= (double) (h->K < 2 ? 2 : h->K);
      if (forced || (double) h->n / km <= 0.001) {
        const int boundary = (64 - 5 * m) & ~7;
        return KeyFmt{m, 5, 4, boundary - 4};
      }
    }
    return byte_fmt_for(h->n, h->K, pl);
  }
  if (k >= (int) pl && k >= 1 && k <= 29) return make_fmt(k, 2, k > 15 ? (k == 29 ? 6 : 5) : 4);
  {
    const int cand[4] = {16, 20, 24, 28};
    for (int i = 0; i < 4; i++) {
      const int m = cand[i];
      if ((unsigned) m < pl) continue;
      if ((double) h->n / (double) (1ull << (2 * m)) <= 0.01 || m == 28) {
        if (tails_few(h, m)) return make_fmt(m, 2, 5);
        break;
      }
    }
  }
  return dna_fmt_for(h->n, pl);
}
// the digit below the symbols holds only the tail field (and the pass over it can be replaced)
bool fmt_tail_digit_alone(const KeyFmt &f) { return f.tb <= 8 && (f.sh + f.tb) % 8 == 0 && f.sh + f.tb <= 64 - f.m * f.b && f.sh + f.tb < 64; }

template <bool DNA> int compact_ties(gtb_esa *h);
template <bool DNA> int build_ranks(gtb_esa *h);
template <bool DNA> int round_local(gtb_esa *h);

// ---- stage 1: bucket table, first-level sort, analysis, special tail, compaction of ties ----
// ext != null: the (key, position) pairs of this range already lie in device memory in text
// order (they were generated slice-wise on several GPUs and exchanged): no text scan here
// ext_tail: the pairs of `ext` whose keys carry a tail, sorted by that field (text order inside): the
// first-level sort then skips them in `ext` and runs over the symbol digits only
template <bool DNA>
int stage_begin(gtb_esa *h, unsigned flags, const PairSrc *ext = nullptr, u64 extcount = 0,
                const TailSrc *ext_tail = nullptr, u64 ntail = 0)
{
  ErrBuf &err = h->err;
  gtb_stats &S = h->stats;
  cudaStream_t st = h->st;
  h->flags = flags;
  h->isa_built = false;
  h->M0 = h->M = 0; h->cur = 0; h->round = 0; h->nllv = 0; h->text_left = 0; h->isa_round = 0;
  h->fmt = choose_fmt(h, h->pl);
  const KeyFmt f = h->fmt;
  h->depth[0] = (u64) f.m;

  // ---- K1/K2 bucket table ----
  // a single range takes the bucket starts from its sorted keys (k_analyze_keys) unless
  // the table is much larger than the text; code ranges need them before the sort
  const bool want_table = h->pl > 0 && ((flags & GTB_WANT_BCK) || !h->full_range);
  bool lb_from_keys = false;
  h->lb_own = false;
  {
    PhaseTimer t(h, &S.ms_count);
    if (want_table && !h->counted) {
      // a range whose offset and width the caller knows (coarse counts) also fills its own codes
      lb_from_keys = (h->full_range && ipow_u64(h->K, h->pl) <= 4 * (h->n - h->S) + 1024) ||
                     (!h->full_range && h->range_given);
      GTB_TRY(count_codes(h, h->pl, !lb_from_keys));
    }
    t.stop();
  }
  u64 klo, khi;
  code_range_to_keys(h, &klo, &khi);
  u64 Ncap = h->n - h->S;
  h->sa_offset = 0;
  if (!h->full_range && h->range_given && !h->counted) {
    h->sa_offset = h->given_offset;
    Ncap = h->given_width;
  } else if (!h->full_range) {
    u32 lb[2];
    GTB_CUDA(cudaMemcpyAsync(&lb[0], h->leftborder.as<u32>() + h->mincode, sizeof(u32),
                             cudaMemcpyDeviceToHost, st));
    GTB_CUDA(cudaMemcpyAsync(&lb[1], h->leftborder.as<u32>() + h->maxcode + 1, sizeof(u32),
                             cudaMemcpyDeviceToHost, st));
    GTB_CUDA(cudaStreamSynchronize(st));
    h->sa_offset = lb[0];
    Ncap = lb[1] - lb[0];
  }
  const u64 tailcnt = h->emit_tail ? h->S + 1 : 0;
  const u64 cap_entries = Ncap + tailcnt;

  // ---- buffers ----
  for (int i = 0; i < 2; i++) {
    GTB_TRY(h->kbuf[i].ensure(sizeof(u64) * (Ncap + 1), err));
    GTB_TRY(h->vbuf[i].ensure(sizeof(u32) * (cap_entries + 1), err));
  }
  GTB_TRY(h->lcp8.ensure(cap_entries + 16, err));
  GTB_TRY(h->dstats.ensure(sizeof(DevStats), err));
  GTB_TRY(h->misc.ensure(256, err));
  {
    DevStats z; memset(&z, 0, sizeof z); z.longest = ~0ull;
    GTB_CUDA(cudaMemcpyAsync(h->dstats.p, &z, sizeof z, cudaMemcpyHostToDevice, st));
  }
  DevStats *dstats = h->dstats.as<DevStats>();

  // ---- K3+K4: fused key generation + LSD radix sort over the significant key bytes ----
  TextSrc<DNA> src = make_src<DNA>(h, klo, khi);
  u64 *kb[2] = {h->kbuf[0].as<u64>(), h->kbuf[1].as<u64>()};
  u32 *vb[2] = {h->vbuf[0].as<u32>(), h->vbuf[1].as<u32>()};
  PassPlan plan; plan.npass = 0; plan.padded = true;
  u64 N = 0;
  h->rw.passes = 0; h->rw.pairs_moved = 0; h->rw.launches = 0; h->rw.ms_hist = 0; h->rw.ms_radix = 0;
  // The pass over the tail digit is dropped when that digit holds nothing else and tails are rare:
  // the keys with a tail leave the text scan, are sorted by their tail (one tiny pass) and follow
  // the full keys as a second source of the first symbol pass -- the result is the one of the sort
  // with the tail digit (equal symbols: full keys first, then by tail, then text order).
  float first_ms0 = 0; u32 first_p0 = 0; u64 first_m0 = 0;     // (the tiny sort of the tail keys is not "first level")
  const bool tail_last = !ext && h->n > 0 && fmt_tail_digit_alone(f) && tails_few(h, f.m);
  if (tail_last) {
    const u64 nw = (h->n + 31) >> 5;
    u64 nt = 0;
    GTB_TRY(h->nearbits.ensure(sizeof(u32) * (nw + 2), err));
    k_near_bits<<<grid_for(nw, 256), 256, 0, st>>>(h->spmask.as<u32>(), nw, (unsigned) f.m, h->nearbits.as<u32>());
    GTB_LAUNCH_CHECK();
    u32 *tileoff = nullptr;
    GTB_TRY(device_scan_u32(h, h->nearbits.as<u32>(), nullptr, nw, 1, &tileoff, &nt));
    S.kernel_launches++;
    TailSrc tsrc{nullptr, nullptr, klo, khi};
    if (nt > 0) {
      for (int i = 0; i < 2; i++) {
        GTB_TRY(h->tailkeys[i].ensure(sizeof(u64) * nt, err));
        GTB_TRY(h->tailpos[i].ensure(sizeof(u32) * nt, err));
      }
      // their positions ascending, their keys, then stably by the tail field
      k_emit_special_tail<<<(unsigned) div_up(nw, SC_TILE), SC_NT, 0, st>>>(h->nearbits.as<u32>(), nw, h->n, tileoff,
          h->tailpos[1].as<u32>(), nullptr, 0, reinterpret_cast<unsigned long long *>(h->misc.as<u64>() + 24) /* unused */);
      GTB_LAUNCH_CHECK();
      k_keys_from_positions<DNA><<<grid_for(nt, 256), 256, 0, st>>>(make_src<DNA>(h, 0, ~0ull), h->tailpos[1].as<u32>(), nt,
                                                                   h->tailkeys[1].as<u64>());
      GTB_LAUNCH_CHECK();
      S.kernel_launches += 2;
      PassPlan tp; tp.npass = 0; tp.padded = false;
      plan_add_bits(tp, f.sh, f.sh + f.tb);
      PairSrc ps{h->tailkeys[1].as<u64>(), h->tailpos[1].as<u32>()};
      u64 *tk[2] = {h->tailkeys[0].as<u64>(), h->tailkeys[1].as<u64>()};
      u32 *tv[2] = {h->tailpos[0].as<u32>(), h->tailpos[1].as<u32>()};
      int tres = 0; u64 tout = 0;
      GTB_TRY(radix_sort(h->rw, st, ps, nt, tk, tv, tp, &tres, &tout, err));
      if (tout != nt) { err.set("internal: the tail keys lost elements"); return -1; }
      tsrc.keys = tk[tres]; tsrc.vals = tv[tres];
    }
    first_ms0 = h->rw.ms_radix; first_p0 = h->rw.passes; first_m0 = h->rw.pairs_moved;
    src.skip_near = true;
    src.hist4 = hist4_pays(h, f.m) && (h->nspecialranges == 0 || h->ranges.p != nullptr) && getenv("GTB200_NO_HIST4") == nullptr;
    plan_add_bits(plan, f.sh + f.tb, 64);
    GTB_TRY((radix_sort<TextSrc<DNA>, TailSrc>(h->rw, st, src, h->n, kb, vb, plan, &h->res, &N, err, nt > 0 ? &tsrc : nullptr, nt)));
    src.skip_near = false; src.hist4 = false;
  } else if (ext && ext_tail) {
    plan_add_bits(plan, f.sh + f.tb, 64);
    PairSrcNoTail nt{ext->keys, ext->vals, f.tailmask()};
    GTB_TRY((radix_sort<PairSrcNoTail, TailSrc>(h->rw, st, nt, extcount, kb, vb, plan, &h->res, &N, err, ntail > 0 ? ext_tail : nullptr, ntail)));
  } else {
    plan_add_bits(plan, f.lowbit() & ~7, 64);
    if (ext) GTB_TRY(radix_sort(h->rw, st, *ext, extcount, kb, vb, plan, &h->res, &N, err));
    else GTB_TRY(radix_sort(h->rw, st, src, h->n, kb, vb, plan, &h->res, &N, err));
  }
  if (ext && N != Ncap) { err.set("exchanged pairs: %llu received, the code range holds %llu suffixes",
                                  (unsigned long long) N, (unsigned long long) Ncap); return -1; }
  if (N > Ncap) { err.set("internal: sorted %llu suffixes, expected at most %llu",
                          (unsigned long long) N, (unsigned long long) Ncap); return -1; }
  h->N = N;
  h->entries = N + tailcnt;
  // the first-level sort alone (plus the partition pass of a sharded scan that fed it)
  S.ms_radix_first = h->rw.ms_radix - first_ms0 + h->ext_ms_radix;
  S.radix_passes_first = h->rw.passes - first_p0 + (h->ext_pairs ? 1u : 0u);
  S.radix_pairs_first = h->rw.pairs_moved - first_m0 + h->ext_pairs;
  u64 *keys = kb[h->res];
  u32 *sa = vb[h->res];
  u8 *lcp8 = h->lcp8.as<u8>();
  // The first-level order is final for all but the tied suffixes (1.5 % of c4): when the caller wants the
  // table on the host (gtb_esa_run_to_host) it starts to cross PCIe NOW, on a stream of its own, while the
  // analysis, the refinement rounds and the lcp kernels run; the chunks that left before the refinement
  // was over are patched afterwards.
  if (h->ec.dst && !h->ec.started && N > 0) {
    GTB_TRY(h->hstage.ensure(err));
    if (!h->st2) GTB_CUDA(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
    h->ec.per = h->hstage.chunk / sizeof(u32);
    h->ec.count = N;
    if (div_up(N, h->ec.per) >= 64) {
      h->ec.dealer = new ChunkDealer(div_up(N, h->ec.per));
      h->ec.rc = 0;
      h->ec.started = true;
      gtb_esa *hh = h;
      const u32 *src_sa = sa;
      h->ec.th = std::thread([hh, src_sa] {
        cudaSetDevice(hh->device);
        hh->ec.rc = staged_d2h(hh->hstage, hh->st2, src_sa, hh->ec.dst, hh->ec.count, sizeof(u32), true, hh->ec.err, hh->ec.dealer);
        if (hh->ec.rc == 0 && cudaStreamSynchronize(hh->st2) != cudaSuccess) { hh->ec.err.set("early copy of the suffix table failed"); hh->ec.rc = -1; }
      });
    }
  }
  GTB_CUDA(cudaMemsetAsync(lcp8, 0, h->entries + 16, st));
  h->first_key = h->last_key = 0;
  if (N > 0) {
    GTB_CUDA(cudaMemcpyAsync(&h->first_key, keys, sizeof(u64), cudaMemcpyDeviceToHost, st));
    GTB_CUDA(cudaMemcpyAsync(&h->last_key, keys + N - 1, sizeof(u64), cudaMemcpyDeviceToHost, st));
  }

  // ---- A: analysis (+ bucket starts) ----
  u64 M0 = 0;
  h->atiles = div_up(N, AN_TILE);
  {
    PhaseTimer t(h, &S.ms_analyze);
    if (N > 0) {
      GTB_TRY(h->tile_a.ensure(sizeof(u32) * (h->atiles + 1), err));
      GTB_TRY(h->tile_b.ensure(sizeof(u32) * (h->atiles + 1), err));
      GTB_TRY(h->hbits.ensure(N / 8 + 16, err));
      GTB_TRY(h->ubits.ensure(N / 8 + 16, err));
      GTB_CUDA(cudaMemsetAsync(h->tile_a.p, 0, sizeof(u32) * (h->atiles + 1), st));   // the warps of a tile add / max
      GTB_CUDA(cudaMemsetAsync(h->tile_b.p, 0, sizeof(u32) * (h->atiles + 1), st));   // into these
      if (lb_from_keys)
        k_analyze_keys<DNA, true><<<grid_for(h->atiles, 1, 148u * 8u), AN_NT, 0, st>>>(keys, N, f, h->pl, h->K, lcp8,
            h->tile_a.as<u32>(), h->tile_b.as<u32>(), dstats, 0, 0ull, h->leftborder.as<u32>(), h->ncodes,
            h->csc.as<u32>(), h->dist.as<u32>(), h->distoff.as<u64>(), h->hbits.as<u8>(), h->ubits.as<u8>(),
            h->full_range ? 0 : h->mincode,
            (h->full_range || h->maxcode + 1 == h->ncodes) ? h->ncodes : h->maxcode, h->sa_offset);
      else
        k_analyze_keys<DNA, false><<<grid_for(h->atiles, 1, 148u * 8u), AN_NT, 0, st>>>(keys, N, f, h->pl, h->K, lcp8,
            h->tile_a.as<u32>(), h->tile_b.as<u32>(), dstats, 0, 0ull, nullptr, 0, nullptr, nullptr, nullptr,
            h->hbits.as<u8>(), h->ubits.as<u8>(), 0, 0, 0);
      GTB_LAUNCH_CHECK();
      GTB_TRY(scan_tiles_sum_max(h, h->tile_a.as<u32>(), h->tile_b.as<u32>(), h->atiles));
      k_find_longest<DNA><<<1, 1, 0, st>>>(src, keys, N, h->sa_offset, dstats);
      GTB_LAUNCH_CHECK();
      S.kernel_launches += 2;
      GTB_CUDA(cudaMemcpyAsync(&M0, h->misc.p, sizeof(u64), cudaMemcpyDeviceToHost, st));
      GTB_CUDA(cudaStreamSynchronize(st));
    }
    if (lb_from_keys && h->full_range) h->counted = true;   // (N = 0: the zeroed table is already right)
    if (lb_from_keys && !h->full_range) h->lb_own = true;   // own codes only: summed over the ranges by the caller
    t.stop();
  }
  S.unresolved_after_first_sort = M0;
  S.doubling_rounds = 0;
  h->M0 = h->M = M0;

  // ---- K7 special tail ----
  {
    PhaseTimer t(h, &S.ms_tail);
    const u64 nw = (h->n + 31) >> 5;
    if (h->emit_tail && h->S > 0 && nw > 0) {
      u32 *tileoff = nullptr;
      GTB_TRY(device_scan_u32(h, h->spmask.as<u32>(), nullptr, nw, 1, &tileoff, nullptr));
      k_emit_special_tail<<<(unsigned) div_up(nw, SC_TILE), SC_NT, 0, st>>>(
          h->spmask.as<u32>(), nw, h->n, tileoff, sa + N, nullptr, h->n - h->S, &dstats->longest);
      GTB_LAUNCH_CHECK();
      S.kernel_launches++;
    }
    if (h->emit_tail) {
      k_set_u32<<<1, 1, 0, st>>>(sa + N + h->S, (u32) h->n);
      GTB_LAUNCH_CHECK();
      S.kernel_launches++;
    }
    t.stop();
  }

  // ---- ties: compaction; a small tie set of a single range is first refined by the text
  // itself (no inverse suffix array needed), ranks are built when that does not finish ----
  if (M0 > 0) {
    PhaseTimer t(h, &S.ms_doubling);
    GTB_TRY(h->uidx0.ensure(sizeof(u32) * M0, err));
    GTB_TRY(h->ugrp0.ensure(sizeof(u32) * M0, err));
    for (int i = 0; i < 2; i++) {
      GTB_TRY(h->uidx[i].ensure(sizeof(u32) * M0, err));
      GTB_TRY(h->ugrp[i].ensure(sizeof(u32) * M0, err));
      GTB_TRY(h->upos[i].ensure(sizeof(u32) * M0, err));
      GTB_TRY(h->kd[i].ensure(sizeof(u64) * M0, err));
      GTB_TRY(h->vd[i].ensure(sizeof(u32) * M0, err));
    }
    GTB_TRY(h->dkeys.ensure(sizeof(u64) * M0, err));
    bool text_first = h->full_range && (M0 <= 65536 || M0 * 128 <= h->n);
    h->text_left = text_first ? 2 : 0;
    if (h->full_range && h->opt_text_rounds >= 0) { h->text_left = (unsigned) h->opt_text_rounds; text_first = h->text_left > 0; }
    GTB_TRY(compact_ties<DNA>(h));
    (void) text_first;
    GTB_CUDA(cudaMemcpyAsync(h->uidx0.p, h->uidx[0].p, sizeof(u32) * M0, cudaMemcpyDeviceToDevice, st));
    GTB_CUDA(cudaMemcpyAsync(h->ugrp0.p, h->ugrp[0].p, sizeof(u32) * M0, cudaMemcpyDeviceToDevice, st));
    t.stop();
  }
  h->bits_lo = bitlen(h->n);
  h->bits_hi = bitlen(N > 0 ? N - 1 : 0);
  h->in_progress = true;
  // the text-driven rounds are local to the range: run them here
  while (h->M > 0 && h->text_left > 0) GTB_TRY(round_local<DNA>(h));
  return 0;
}

// list of the tied suffixes (SA index, position, group head)
template <bool DNA>
int compact_ties(gtb_esa *h)
{
  ErrBuf &err = h->err;
  if (h->N > 0) {
    k_compact_keys<<<grid_for(h->atiles, 1, 148u * 8u), AN_NT, 0, h->st>>>(h->hbits.as<u8>(), h->ubits.as<u8>(),
        h->vbuf[h->res].as<u32>(), h->N, h->tile_a.as<u32>(), h->tile_b.as<u32>(), h->atiles, h->M0, h->uidx[0].as<u32>(),
        h->upos[0].as<u32>(), h->ugrp[0].as<u32>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
  }
  return 0;
}

template <bool DNA>
RankMap<DNA> make_rankmap(gtb_esa *h)
{
  RankMap<DNA> rm;
  rm.src = make_src<DNA>(h, 0, ~0ull);
  rm.keys = h->kbuf[h->res].as<u64>();
  rm.sa = h->vbuf[h->res].as<u32>();
  rm.N = h->N;
  rm.rw = h->rankwords.as<uint4>(); rm.trank = h->trank.as<u32>();
  rm.leftborder = ((h->counted || h->lb_own) && h->pl > 0) ? h->leftborder.as<u32>() : nullptr;
  rm.own_last = (h->lb_own && !h->counted) ? h->maxcode : ~0ull;
  rm.pl = h->pl; rm.K = h->K;
  rm.n = h->n; rm.nonspecials = h->n - h->S; rm.sa_offset = h->sa_offset;
  return rm;
}

// exclusive popcount prefix per word of a bitmap
int popcount_prefix(gtb_esa *h, const u32 *bits, u32 *pre, u64 nwords)
{
  ErrBuf &err = h->err;
  if (nwords == 0) return 0;
  u32 *tileoff = nullptr;
  GTB_TRY(device_scan_u32(h, bits, nullptr, nwords, 1, &tileoff, nullptr));
  k_scan_apply_popc<<<(unsigned) div_up(nwords, SC_TILE), SC_NT, 0, h->st>>>(bits, pre, nwords, tileoff);
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches++;
  return 0;
}

// the sparse rank map of this range in its current state of refinement (RankMap,
// gtb_esa_kernels.cuh): no inverse suffix array is ever built
template <bool DNA>
int build_ranks(gtb_esa *h)
{
  ErrBuf &err = h->err;
  cudaStream_t st = h->st;
  if (h->isa_built) return 0;
  const u64 nw = (h->n >> 5) + 2;
  if (h->S > 0) {
    GTB_TRY(h->spre.ensure(sizeof(u32) * nw, err));
    GTB_TRY(popcount_prefix(h, h->spmask.as<u32>(), h->spre.as<u32>(), nw));
  }
  GTB_TRY(h->tbits.ensure(sizeof(u32) * nw, err));
  GTB_CUDA(cudaMemsetAsync(h->tbits.p, 0, sizeof(u32) * nw, st));
  if (h->M0 > 0) {
    GTB_TRY(h->tpre.ensure(sizeof(u32) * nw, err));
    GTB_TRY(h->trank.ensure(sizeof(u32) * h->M0, err));
    k_tied_bits<<<grid_for(h->M0, 256), 256, 0, st>>>(h->uidx0.as<u32>(), h->M0, h->vbuf[h->res].as<u32>(),
                                                       h->tbits.as<u32>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
    GTB_TRY(popcount_prefix(h, h->tbits.as<u32>(), h->tpre.as<u32>(), nw));
  }
  GTB_TRY(h->rankwords.ensure(sizeof(uint4) * nw, err));
  k_pack_rankwords<<<grid_for(nw, 256), 256, 0, st>>>(h->spmask.as<u32>(), h->S > 0 ? h->spre.as<u32>() : nullptr,
      h->tbits.as<u32>(), h->M0 > 0 ? h->tpre.as<u32>() : nullptr, nw, h->rankwords.as<uint4>());
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches++;
  if (h->M0 > 0) {
    RankMap<DNA> rm = make_rankmap<DNA>(h);
    if (h->round > 0) {          // text-driven rounds have resolved some of the initial ties
      k_trank_resolved<DNA><<<grid_for(h->M0, 256), 256, 0, st>>>(rm, h->uidx0.as<u32>(), h->M0);
      GTB_LAUNCH_CHECK();
      h->stats.kernel_launches++;
    }
    if (h->M > 0) {
      k_trank_tied<DNA><<<grid_for(h->M, 256), 256, 0, st>>>(rm, h->upos[h->cur].as<u32>(),
                                                             h->ugrp[h->cur].as<u32>(), h->M);
      GTB_LAUNCH_CHECK();
      h->stats.kernel_launches++;
    }
  }
  h->isa_built = true;
  return 0;
}

// sort the ties by the keys in h->dkeys (group head : 32, refinement key : 32), write the
// order back, re-compact.  tm: bits that mark a key as resolved (text-driven rounds);
// keybits: significant bits of the refinement key
template <bool DNA>
int round_sort_apply(gtb_esa *h, u64 tm, int keybits, bool text_round)
{
  ErrBuf &err = h->err;
  cudaStream_t st = h->st;
  gtb_stats &S = h->stats;
  const u64 M = h->M;
  const int cur = h->cur;
  if (h->round >= 62) { err.set("internal: refinement did not converge"); return -1; }
  PassPlan dp; dp.npass = 0; dp.padded = true;
  if (text_round) plan_add_bits(dp, 32 - ((keybits + 7) & ~7), 32);   // the key is top-aligned in its half
  else plan_add_bits(dp, 0, keybits);
  plan_add_bits(dp, 32, 32 + h->bits_hi);
  PairSrc ps{h->dkeys.as<u64>(), h->upos[cur].as<u32>()};
  u64 *kk[2] = {h->kd[0].as<u64>(), h->kd[1].as<u64>()};
  u32 *vv[2] = {h->vd[0].as<u32>(), h->vd[1].as<u32>()};
  int r2 = 0; u64 nout = 0;
  GTB_TRY(radix_sort(h->rw, st, ps, M, kk, vv, dp, &r2, &nout, err));
  if (nout != M) { err.set("internal: refinement sort lost elements"); return -1; }
  const u64 dt = div_up(M, AN_TILE);
  GTB_TRY(h->tile_c.ensure(sizeof(u32) * (dt + 1), err));
  GTB_TRY(h->tile_d.ensure(sizeof(u32) * (dt + 1), err));
  k_analyze_dkeys<<<(unsigned) dt, AN_NT, 0, st>>>(kk[r2], M, tm, h->tile_c.as<u32>(), h->tile_d.as<u32>());
  GTB_LAUNCH_CHECK();
  GTB_TRY(scan_tiles_sum_max(h, h->tile_c.as<u32>(), h->tile_d.as<u32>(), dt));
  k_apply_dkeys<DNA><<<(unsigned) dt, AN_NT, 0, st>>>(kk[r2], vv[r2], h->uidx[cur].as<u32>(), M, tm,
      h->tile_c.as<u32>(), h->tile_d.as<u32>(), h->vbuf[h->res].as<u32>(),
      make_rankmap<DNA>(h), h->isa_built ? 1 : 0,
      h->lcp8.as<u8>(), (u8) h->round, h->sa_offset,
      h->uidx[cur ^ 1].as<u32>(), h->upos[cur ^ 1].as<u32>(), h->ugrp[cur ^ 1].as<u32>(),
      h->dstats.as<DevStats>());
  GTB_LAUNCH_CHECK();
  S.kernel_launches += 2;
  u64 Mnext = 0;
  GTB_CUDA(cudaMemcpyAsync(&Mnext, h->misc.p, sizeof(u64), cudaMemcpyDeviceToHost, st));
  GTB_CUDA(cudaStreamSynchronize(st));
  h->M = Mnext;
  h->cur ^= 1;
  h->round++;
  S.doubling_rounds = h->round;
  return 0;
}

// one local refinement round: by the next symbols of the text while text rounds are
// allowed, else by the ranks of the suffixes depth[round] further (prefix doubling)
template <bool DNA>
int round_local(gtb_esa *h)
{
  ErrBuf &err = h->err;
  if (h->M == 0) { h->depth[h->round + 1] = h->depth[h->round]; h->round++; return 0; }
  if (h->round >= 62) { err.set("internal: refinement did not converge"); return -1; }
  PhaseTimer t(h, &h->stats.ms_doubling);
  const u64 hlen = h->depth[h->round];
  if (h->text_left > 0) {
    const KeyFmt g = DNA ? dna_tfmt() : byte_tfmt();
    k_build_tkeys<DNA><<<grid_for(h->M, 256), 256, 0, h->st>>>(make_src<DNA>(h, 0, ~0ull), g,
        h->upos[h->cur].as<u32>(), h->ugrp[h->cur].as<u32>(), h->M, hlen, h->dkeys.as<u64>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
    h->depth[h->round + 1] = hlen + (u64) g.m;
    h->text_left--;
    GTB_TRY(round_sort_apply<DNA>(h, g.tailmask() >> 32, g.m * g.b + g.tb, true));
  } else {
    GTB_TRY(build_ranks<DNA>(h));
    GTB_TRY(h->sendidx.ensure(sizeof(u32) * h->M, err));       // queue of the partners that need a search
    unsigned int *qcount = reinterpret_cast<unsigned int *>(h->misc.as<u64>() + 20);
    GTB_CUDA(cudaMemsetAsync(qcount, 0, sizeof(unsigned int), h->st));
    k_build_dkeys<DNA><<<grid_for(h->M, 256), 256, 0, h->st>>>(make_rankmap<DNA>(h), h->upos[h->cur].as<u32>(),
        h->ugrp[h->cur].as<u32>(), h->M, hlen, h->dkeys.as<u64>(), h->sendidx.as<u32>(), qcount,
        (h->isa_round < (unsigned) h->opt_pairs_by_text) ? 1 : 0);
    GTB_LAUNCH_CHECK();
    h->isa_round++;
    k_build_dkeys_search<DNA><<<grid_for(h->M, 256, 148u * 8u), 256, 0, h->st>>>(make_rankmap<DNA>(h), h->upos[h->cur].as<u32>(),
        h->ugrp[h->cur].as<u32>(), hlen, h->dkeys.as<u64>(), h->sendidx.as<u32>(), qcount);
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches += 2;
    h->depth[h->round + 1] = 2 * hlen;
    GTB_TRY(round_sort_apply<DNA>(h, 0ull, h->bits_lo, false));
  }
  t.stop();
  return 0;
}

// multi-range round, part 1: ranks this range can read itself; positions to ask the other ranges
template <bool DNA>
int round_prepare(gtb_esa *h, const u64 *first_keys, int nranges, int mine, u32 *dev_send,
                  u64 capacity, u64 *counts_out)
{
  ErrBuf &err = h->err;
  cudaStream_t st = h->st;
  for (int i = 0; i < nranges; i++) counts_out[i] = 0;
  h->pending_send = 0;
  if (h->M == 0) return 0;
  if (nranges < 1 || nranges > MAX_RANGES || mine < 0 || mine >= nranges) { err.set("bad range table"); return -1; }
  if (h->round >= 62) { err.set("internal: prefix doubling did not converge"); return -1; }
  PhaseTimer t(h, &h->stats.ms_doubling);
  const u64 M = h->M;
  GTB_TRY(build_ranks<DNA>(h));
  GTB_TRY(h->ranks.ensure(sizeof(u32) * M, err));
  GTB_TRY(h->owner.ensure(M, err));
  GTB_TRY(h->sendidx.ensure(sizeof(u32) * M, err));
  GTB_TRY(h->rcounts.ensure(sizeof(unsigned int) * 3 * MAX_RANGES, err));
  GTB_CUDA(cudaMemsetAsync(h->rcounts.p, 0, sizeof(unsigned int) * 3 * MAX_RANGES, st));
  RangeBounds rb; rb.n = nranges; rb.mine = mine;
  for (int i = 0; i < nranges; i++) rb.first_key[i] = first_keys[i];
  rb.first_key[0] = 0;
  const u64 hlen = h->depth[h->round];
  unsigned int *cnt = h->rcounts.as<unsigned int>();
  k_round_classify<DNA><<<grid_for(M, 256), 256, 0, st>>>(make_src<DNA>(h, 0, ~0ull), h->upos[h->cur].as<u32>(),
      M, hlen, rb, make_rankmap<DNA>(h), h->ranks.as<u32>(), h->owner.as<u8>(), cnt);
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches++;
  unsigned int hc[MAX_RANGES], off[MAX_RANGES];
  GTB_CUDA(cudaMemcpyAsync(hc, cnt, sizeof(unsigned int) * nranges, cudaMemcpyDeviceToHost, st));
  GTB_CUDA(cudaStreamSynchronize(st));
  u64 total = 0;
  for (int i = 0; i < nranges; i++) { off[i] = (unsigned int) total; total += hc[i]; counts_out[i] = hc[i]; }
  if (total > capacity) { err.set("send buffer too small: %llu positions", (unsigned long long) total); return -1; }
  if (total > 0) {
    GTB_CUDA(cudaMemcpyAsync(cnt + MAX_RANGES, off, sizeof(unsigned int) * nranges, cudaMemcpyHostToDevice, st));
    k_round_fill<<<grid_for(M, 256), 256, 0, st>>>(h->upos[h->cur].as<u32>(), h->owner.as<u8>(), M, hlen,
        cnt + MAX_RANGES, cnt + 2 * MAX_RANGES, dev_send, h->sendidx.as<u32>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
  }
  GTB_CUDA(cudaStreamSynchronize(st));
  h->pending_send = total;
  t.stop();
  return 0;
}

// multi-range round, part 2: the answers arrived (same order as the send buffer)
template <bool DNA>
int round_finish(gtb_esa *h, const u32 *dev_answers)
{
  ErrBuf &err = h->err;
  if (h->M == 0) { h->depth[h->round + 1] = 2 * h->depth[h->round]; h->round++; return 0; }
  PhaseTimer t(h, &h->stats.ms_doubling);
  if (h->pending_send > 0) {
    k_round_scatter<<<grid_for(h->pending_send, 256), 256, 0, h->st>>>(dev_answers, h->sendidx.as<u32>(),
        h->pending_send, h->ranks.as<u32>());
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
  }
  k_build_dkeys_ranks<<<grid_for(h->M, 256), 256, 0, h->st>>>(h->ugrp[h->cur].as<u32>(), h->ranks.as<u32>(),
      h->M, h->dkeys.as<u64>());
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches++;
  h->depth[h->round + 1] = 2 * h->depth[h->round];
  GTB_TRY(round_sort_apply<DNA>(h, 0ull, h->bits_lo, false));
  t.stop();
  return 0;
}

// ---- last stage: exact lcp of the deep pairs, .llv list, stats ----
template <bool DNA>
int stage_end(gtb_esa *h)
{
  ErrBuf &err = h->err;
  gtb_stats &S = h->stats;
  cudaStream_t st = h->st;
  const u64 M0 = h->M0, N = h->N;
  DevStats *dstats = h->dstats.as<DevStats>();
  u32 *sa = h->vbuf[h->res].as<u32>();
  u8 *lcp8 = h->lcp8.as<u8>();
  if (h->M != 0) { err.set("prefix doubling not finished: %llu suffixes still tied", (unsigned long long) h->M); return -1; }
  if (M0 > 0 && (h->flags & GTB_WANT_LCP)) {
    PhaseTimer t(h, &S.ms_lcp);
    GTB_TRY(h->ulcp.ensure(sizeof(u32) * (M0 > h->S + 1 ? M0 : h->S + 1), err));
    DepthTab dt;
    for (int i = 0; i < 64; i++) dt.d[i] = i <= (int) h->round ? h->depth[i] : 0;
    // (the refinement keys are dead: their buffer queues the pairs with long extensions)
    GTB_TRY(h->dkeys.ensure(sizeof(u64) * M0, err));
    unsigned long long *qcount = reinterpret_cast<unsigned long long *>(h->misc.as<u64>() + 8);
    GTB_CUDA(cudaMemsetAsync(qcount, 0, sizeof(unsigned long long), st));
    k_deep_lcp<DNA><<<grid_for(M0, 128), 128, 0, st>>>(h->uidx0.as<u32>(), h->ugrp0.as<u32>(), M0, sa,
        h->words.as<u64>(), h->bytes.as<u8>(), h->spmask.as<u32>(), h->n, dt, lcp8,
        h->ulcp.as<u32>(), dstats, h->dkeys.as<u64>(), qcount);
    GTB_LAUNCH_CHECK();
    S.kernel_launches++;
    if (DNA) {
      k_deep_lcp_long<<<148 * 8, 256, 0, st>>>(h->dkeys.as<u64>(), qcount, h->uidx0.as<u32>(), sa,
          h->words.as<u64>(), h->spmask.as<u32>(), h->n, lcp8, h->ulcp.as<u32>(), dstats);
      GTB_LAUNCH_CHECK();
      S.kernel_launches++;
    }
    DevStats hs;
    GTB_CUDA(cudaMemcpyAsync(&hs, dstats, sizeof hs, cudaMemcpyDeviceToHost, st));
    GTB_CUDA(cudaStreamSynchronize(st));
    if (hs.numlarge > 0) {
      GTB_TRY(h->llvflags.ensure(sizeof(u32) * M0, err));
      GTB_TRY(h->llv.ensure(sizeof(u64) * 2 * hs.numlarge, err));
      k_llv_flags<<<grid_for(M0, 256), 256, 0, st>>>(h->ulcp.as<u32>(), M0, h->llvflags.as<u32>());
      GTB_LAUNCH_CHECK();
      u64 tot = 0;
      GTB_TRY(device_scan_u32(h, h->llvflags.as<u32>(), h->llvflags.as<u32>(), M0, 0, nullptr, &tot));
      if (tot != hs.numlarge) { err.set("internal: llv count mismatch"); return -1; }
      k_llv_emit<<<grid_for(M0, 256), 256, 0, st>>>(h->ulcp.as<u32>(), h->uidx0.as<u32>(), M0,
          h->llvflags.as<u32>(), h->sa_offset, h->llv.as<u64>());
      GTB_LAUNCH_CHECK();
      S.kernel_launches += 2;
      h->nllv = hs.numlarge;
    }
    t.stop();
  }
  DevStats hs;
  GTB_CUDA(cudaMemcpyAsync(&hs, dstats, sizeof hs, cudaMemcpyDeviceToHost, st));
  GTB_CUDA(cudaStreamSynchronize(st));
  S.totallength = h->n; S.specialcharacters = h->S; S.nonspecials = N;
  S.sa_offset = h->sa_offset;
  S.longest = hs.longest;
  if (h->n == 0 && h->emit_tail) S.longest = 0;   // the empty text: suffix 0 is the final entry n
  S.numoflargelcpvalues = hs.numlarge;
  S.maxbranchdepth = hs.maxlcp;
  S.lcptabsum = (double) hs.lcpsum;
  S.prefixlength = h->pl; S.numofchars = h->K;
  // (a slice partition that preceded this run is one more pass of the same kernel)
  S.radix_passes = h->rw.passes + (h->ext_pairs ? 1u : 0u);
  S.radix_pairs_moved = h->rw.pairs_moved + h->ext_pairs;
  S.kernel_launches += h->rw.launches + h->ext_launches;
  S.ms_hist = h->rw.ms_hist + h->ext_ms_keygen; S.ms_radix = h->rw.ms_radix + h->ext_ms_radix;
  h->in_progress = false;
  return 0;
}

} // namespace

static int check_run_args(gtb_esa *h, unsigned prefixlength, unsigned flags)
{
  ErrBuf &err = h->err;
  if (!h->have_input) { err.set("no input set"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  const unsigned maxpl = h->dna ? 15u : 7u;     // gt_maxbasepower, initbasepower.c:23
  if (prefixlength > maxpl || (!h->dna && ipow_u64(h->K, prefixlength) > 0xffffffffull)) {
    err.set("prefix length %u is too large for alphabet size %u", prefixlength, h->K); return -1;
  }
  if (prefixlength == 0 && ((flags & GTB_WANT_BCK) || !h->full_range)) {
    err.set("prefixlength 0 cannot be combined with a bucket table or a code range"); return -1;
  }
  if (!h->full_range && h->maxcode >= ipow_u64(h->K, prefixlength)) {
    err.set("code range beyond numofchars^prefixlength"); return -1;
  }
  const float up = h->stats.ms_upload;
  memset(&h->stats, 0, sizeof h->stats);
  h->stats.ms_upload = up;
  if (flags & GTB_REUSE_COUNTS) { h->stats.ms_count = h->ms_count_ext; h->stats.ms_total = h->ms_count_ext + h->ms_part_ext; }
  h->ext_ms_keygen = 0;
  h->ext_ms_radix = (flags & GTB_REUSE_COUNTS) ? h->ms_part_ext : 0;
  h->ext_pairs = (flags & GTB_REUSE_COUNTS) ? h->part_pairs_ext : 0;
  h->ext_launches = (flags & GTB_REUSE_COUNTS) ? h->part_launches_ext : 0;
  h->ms_count_ext = 0; h->ms_part_ext = 0; h->part_pairs_ext = 0; h->part_launches_ext = 0;
  if (h->pl != prefixlength || !(flags & GTB_REUSE_COUNTS)) h->counted = false;
  h->pl = prefixlength;
  h->ran = false;
  return 0;
}

// no C++ exception crosses the C-ABI: an entry point that allocates host memory or starts threads runs
// its body through this
template <class F>
static int no_throw(ErrBuf &err, F body)
{
  try { return body(); }
  catch (const std::exception &e) { err.set("libgtb200: %s", e.what()); return -1; }
  catch (...) { err.set("libgtb200: unexpected exception"); return -1; }
}

// wraps one stage: device time of the stage is added to ms_total
template <typename F>
static int timed_stage(gtb_esa *h, F body)
{
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  cudaEvent_t e0, e1;
  GTB_CUDA(cudaEventCreate(&e0)); GTB_CUDA(cudaEventCreate(&e1));
  GTB_CUDA(cudaEventRecord(e0, h->st));
  int rc = body();
  cudaEventRecord(e1, h->st);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  h->stats.ms_total += ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc == 0) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err.set("CUDA error: %s", cudaGetErrorString(e)); rc = -1; }
  }
  return rc;
}

// checksum of `count` entries at device pointer v (global index of the first = index_base), added to *acc
template <typename T>
static int hash_table(gtb_esa *h, const T *v, u64 count, u64 index_base, u64 *acc)
{
  ErrBuf &err = h->err;
  if (count == 0) return 0;
  GTB_TRY(h->misc.ensure(256, err));
  unsigned long long *d = reinterpret_cast<unsigned long long *>(h->misc.as<u64>() + 16);
  GTB_CUDA(cudaMemsetAsync(d, 0, sizeof *d, h->st));
  k_mixhash<T><<<grid_for(count, 256, 148u * 8u), 256, 0, h->st>>>(v, count, index_base, d);
  GTB_LAUNCH_CHECK();
  unsigned long long r = 0;
  GTB_CUDA(cudaMemcpyAsync(&r, d, sizeof r, cudaMemcpyDeviceToHost, h->st));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  *acc += (u64) r;
  return 0;
}

// a handle that borrowed its input gives it back (before it takes another input or goes away)
static void return_borrowed_input(gtb_esa *h)
{
  if (h->lender) {
    h->lender->borrowers--;
    h->lender = nullptr;
    h->words.release(); h->bytes.release(); h->spmask.release(); h->sepbits.release(); h->ranges.release();
    h->have_input = false;
  }
}
static int refuse_while_lent(gtb_esa *h, const char *what)
{
  if (h->borrowers > 0) {
    h->err.set("%s: %d other handle(s) still use this handle's sequence in HBM (gtb_esa_share_input); delete them or give them another input first", what, h->borrowers);
    return -1;
  }
  return 0;
}

#include "gtb_shard_host.cuh"

// =============================== C-ABI =================================================
extern "C" {

int gtb_abi_version(void) { return GTB200_ABI_VERSION; }

int gtb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// live handles per device (gtb_esa_new / gtb_esa_delete), for gtb_release_devices
static std::mutex g_live_mutex;
static int g_live[64];
static bool g_used[64];

static void live_handles(int device, int delta)
{
  if (device < 0 || device >= 64) return;
  std::lock_guard<std::mutex> lock(g_live_mutex);
  g_live[device] += delta;
  g_used[device] = true;
}

int gtb_release_devices(void)
{
  int released = 0;
  for (int d = 0; d < 64; d++) {
    {
      std::lock_guard<std::mutex> lock(g_live_mutex);
      if (!g_used[d] || g_live[d] > 0) continue;
      g_used[d] = false;
    }
    if (cudaSetDevice(d) == cudaSuccess && cudaDeviceReset() == cudaSuccess) released++;
    else cudaGetLastError();
    g_context_generation[d].fetch_add(1);
  }
  return released;
}

void gtb_bck_sizes(unsigned K, unsigned pl, uint64_t *nall, uint64_t *nspecial, uint64_t *ndist)
{
  u64 d = 0;
  for (unsigned i = 1; i + 2 <= pl; i++) d += ipow_u64(K, i);
  if (nall) *nall = pl ? ipow_u64(K, pl) : 0;
  if (nspecial) *nspecial = pl ? ipow_u64(K, pl - 1) : 0;
  if (ndist) *ndist = d;
}

gtb_esa *gtb_esa_new(int device, char *errbuf, size_t errlen)
{
  auto fail = [&](const char *what, cudaError_t e) -> gtb_esa * {
    if (errbuf && errlen) snprintf(errbuf, errlen, "%s: %s", what, cudaGetErrorString(e));
    return nullptr;
  };
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    if (errbuf && errlen)
      snprintf(errbuf, errlen, "libgtb200: no CUDA device available (%s); there is no CPU fallback",
               e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    if (errbuf && errlen) snprintf(errbuf, errlen, "libgtb200: device %d out of range (%d devices)", device, ndev);
    return nullptr;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  gtb_esa *h = new (std::nothrow) gtb_esa();
  if (!h) { if (errbuf && errlen) snprintf(errbuf, errlen, "out of host memory"); return nullptr; }
  h->device = device;
  live_handles(device, +1);
  if (const char *e = getenv("GTB200_KEY_SYMBOLS")) h->opt_key_symbols = atoi(e);
  if (const char *e = getenv("GTB200_TEXT_ROUNDS")) h->opt_text_rounds = atoi(e) > 8 ? 8 : atoi(e);
  if (const char *e = getenv("GTB200_TAIL_LAST")) h->opt_tail_last = atoi(e) ? 1 : 0;
  if (const char *e = getenv("GTB200_PAIRS_BY_TEXT")) h->opt_pairs_by_text = atoi(e) < 0 ? 0 : atoi(e);
  memset(&h->stats, 0, sizeof h->stats);
  if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess) {
    live_handles(device, -1); delete h; return fail("cudaStreamCreate", e);
  }
  if (radix_work_init(h->rw, h->err) != 0) {
    if (errbuf && errlen) snprintf(errbuf, errlen, "%s", h->err.msg);
    cudaStreamDestroy(h->st); live_handles(device, -1); delete h; return nullptr;
  }
  return h;
}

void gtb_esa_delete(gtb_esa *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  return_borrowed_input(h);
  if (h->borrowers > 0) {
    // other handles still read this sequence: it is left in HBM (and the handle with it) rather than
    // freed under them
    fprintf(stderr, "libgtb200: gtb_esa_delete of a handle whose sequence %d other handle(s) still use -- kept\n", h->borrowers);
    return;
  }
  DevBuf *all[] = {&h->words, &h->bytes, &h->spmask, &h->ranges, &h->sepbits, &h->seppos, &h->leftborder, &h->csc, &h->dist,
                   &h->distoff, &h->kbuf[0], &h->kbuf[1], &h->vbuf[0], &h->vbuf[1], &h->lcp8, &h->coarse, &h->hbits, &h->ubits, &h->tbits, &h->tpre, &h->trank, &h->spre,
                   &h->tile_a, &h->tile_b, &h->tile_c, &h->tile_d, &h->scantmp, &h->dstats, &h->misc, &h->uidx0, &h->ugrp0,
                   &h->uidx[0], &h->uidx[1], &h->ugrp[0], &h->ugrp[1], &h->upos[0], &h->upos[1],
                   &h->dkeys, &h->kd[0], &h->kd[1], &h->vd[0], &h->vd[1], &h->ulcp, &h->llvflags, &h->llv,
                   &h->ranks, &h->owner, &h->sendidx, &h->rcounts, &h->rankwords, &h->peertab,
                   &h->nearbits, &h->tailkeys[0], &h->tailkeys[1], &h->tailpos[0], &h->tailpos[1]};
  for (DevBuf *b : all) b->release();
  for (auto &m : h->imports) vmm_free(m.second.ptr, m.second.mh, m.second.size);
  h->imports.clear();
  if (h->ec.started && h->ec.th.joinable()) h->ec.th.join();
  delete h->ec.dealer;
  if (h->patch_host) cudaFreeHost(h->patch_host);
  stop_ipc_receiver(h);
  for (auto &f : h->pending_fds) close(f.fd);
  h->pending_fds.clear();
  if (h->ipc_sock >= 0) close(h->ipc_sock);
  radix_work_free(h->rw);
  h->hstage.release();
  if (h->st2) cudaStreamDestroy(h->st2);
  if (h->st3) cudaStreamDestroy(h->st3);
  cudaStreamDestroy(h->st);
  live_handles(h->device, -1);
  delete h;
}

const char *gtb_esa_error(const gtb_esa *h) { return h ? h->err.msg : "null handle"; }

int gtb_esa_set_readmode(gtb_esa *h, unsigned readmode)
{
  if (!h) return -1;
  if (readmode > 3) { h->err.set("unknown readmode, must be fwd or rev or cpl or rcl"); return -1; }
  if (readmode != h->readmode) {
    if (h->lender) return_borrowed_input(h);
    h->have_input = false; h->counted = false; h->ran = false;
  }
  h->readmode = readmode;
  return 0;
}

int gtb_esa_set_input_2bit(gtb_esa *h, const uint64_t *twobitenc, uint64_t nwords,
                           uint64_t n, const gtb_range *specials, uint64_t nranges)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  GTB_TRY(refuse_while_lent(h, "gtb_esa_set_input_2bit"));
  return_borrowed_input(h);
  if (n + 1 >= 0xffffffffull) { err.set("totallength %llu needs 64-bit suffix tables: not supported by this build (u32 positions)", (unsigned long long) n); return -1; }
  if (nwords < (n + 31) / 32) { err.set("twobitencoding too short: %llu words for %llu bases", (unsigned long long) nwords, (unsigned long long) n); return -1; }
  PhaseTimer t(h, &h->stats.ms_upload);
  h->dna = true; h->K = 4; h->n = n; h->counted = false; h->ran = false; h->have_sep = false;
  h->nspecialranges = nranges;
  const u64 need = (n >> 5) + 4;
  GTB_TRY(h->words.ensure(sizeof(u64) * need, err));
  const u64 ncopy = nwords < need ? nwords : need;
  const bool rev = (h->readmode & 1u) != 0, cpl = (h->readmode & 2u) != 0;
  if (h->readmode == 0) {
    GTB_CUDA(cudaMemsetAsync(h->words.p, 0, sizeof(u64) * need, h->st));
    if (ncopy) GTB_CUDA(cudaMemcpyAsync(h->words.p, twobitenc, sizeof(u64) * ncopy, cudaMemcpyHostToDevice, h->st));
  } else {
    // -dir rev|cpl|rcl: the words go to scratch memory and are rewritten in the read direction
    GTB_TRY(h->kbuf[0].ensure(sizeof(u64) * need, err));
    GTB_CUDA(cudaMemsetAsync(h->kbuf[0].p, 0, sizeof(u64) * need, h->st));
    if (ncopy) GTB_CUDA(cudaMemcpyAsync(h->kbuf[0].p, twobitenc, sizeof(u64) * ncopy, cudaMemcpyHostToDevice, h->st));
    k_readmode_words<<<grid_for(need, 256), 256, 0, h->st>>>(h->kbuf[0].as<u64>(), h->words.as<u64>(), n, need,
                                                             rev ? 1 : 0, cpl ? 1 : 0);
    GTB_LAUNCH_CHECK();
  }
  for (u64 r = 0; r < nranges; r++) {
    if (specials[r].start >= specials[r].end || specials[r].end > n ||
        (r > 0 && specials[r].start < specials[r - 1].end)) {
      err.set("special range %llu [%llu,%llu) is empty, unordered or beyond the text", (unsigned long long) r,
              (unsigned long long) specials[r].start, (unsigned long long) specials[r].end);
      return -1;
    }
  }
  if (rev && nranges > 0) {
    // the special runs in read direction: mirrored and in reverse order
    gtb_range *mir = static_cast<gtb_range *>(malloc(sizeof(gtb_range) * nranges));
    if (!mir) { err.set("out of host memory for %llu special ranges", (unsigned long long) nranges); return -1; }
    for (u64 r = 0; r < nranges; r++) {
      mir[r].start = n - specials[nranges - 1 - r].end;
      mir[r].end = n - specials[nranges - 1 - r].start;
    }
    const int rc = build_mask(h, mir, nranges);
    if (rc == 0) cudaStreamSynchronize(h->st);      // (the upload of `mir` is asynchronous)
    free(mir);
    GTB_TRY(rc);
  } else {
    GTB_TRY(build_mask(h, specials, nranges));
  }
  t.stop();
  h->have_input = true;
  return 0;
}

int gtb_esa_set_input_bytes(gtb_esa *h, const uint8_t *symbols, uint64_t n, unsigned K)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  GTB_TRY(refuse_while_lent(h, "gtb_esa_set_input_bytes"));
  return_borrowed_input(h);
  if (n + 1 >= 0xffffffffull) { err.set("totallength %llu needs 64-bit suffix tables: not supported by this build", (unsigned long long) n); return -1; }
  if (K < 1 || K > 31) { err.set("numofchars %u not supported by the byte path (1..31)", K); return -1; }
  if ((h->readmode & 2u) && K != 4) {     // sfx-run.c:541-549
    err.set("option -%s only can be used for DNA alphabets", h->readmode == 2 ? "cpl" : "rcl");
    return -1;
  }
  PhaseTimer t(h, &h->stats.ms_upload);
  h->dna = false; h->K = K; h->n = n; h->counted = false; h->ran = false;
  const u64 need = ((n + 31) / 32) * 32 + 64;
  GTB_TRY(h->bytes.ensure(need, err));
  GTB_CUDA(cudaMemsetAsync(h->bytes.p, 0xff, need, h->st));
  if (n && h->readmode == 0) GTB_CUDA(cudaMemcpyAsync(h->bytes.p, symbols, n, cudaMemcpyHostToDevice, h->st));
  if (n && h->readmode != 0) {
    GTB_TRY(h->kbuf[0].ensure(n, err));
    GTB_CUDA(cudaMemcpyAsync(h->kbuf[0].p, symbols, n, cudaMemcpyHostToDevice, h->st));
    k_readmode_bytes<<<grid_for(n, 256), 256, 0, h->st>>>(h->kbuf[0].as<u8>(), h->bytes.as<u8>(), n,
                                                          (h->readmode & 1u) ? 1 : 0, (h->readmode & 2u) ? 1 : 0);
    GTB_LAUNCH_CHECK();
  }
  GTB_TRY(build_mask(h, nullptr, 0));
  t.stop();
  h->have_input = true;
  return 0;
}

int gtb_esa_share_input(gtb_esa *h, const gtb_esa *src)
{
  if (!h || !src) return -1;
  if (!src->have_input) { h->err.set("gtb_esa_share_input: source has no input"); return -1; }
  if (h->device != src->device) { h->err.set("gtb_esa_share_input: handles live on different devices"); return -1; }
  if (h == src || src->lender) { h->err.set("gtb_esa_share_input: the source must own its sequence"); return -1; }
  GTB_TRY(refuse_while_lent(h, "gtb_esa_share_input"));
  return_borrowed_input(h);
  const_cast<gtb_esa *>(src)->borrowers++;
  h->lender = const_cast<gtb_esa *>(src);
  h->dna = src->dna; h->K = src->K; h->n = src->n; h->S = src->S; h->nmaskwords = src->nmaskwords;
  h->nspecialranges = src->nspecialranges;
  h->readmode = src->readmode;
  h->words.borrow(src->words); h->bytes.borrow(src->bytes); h->spmask.borrow(src->spmask);
  h->ranges.borrow(src->ranges);
  h->sepbits.borrow(src->sepbits); h->have_sep = src->have_sep;
  h->counted = false; h->ran = false; h->have_input = true;
  return 0;
}

int gtb_esa_set_code_range(gtb_esa *h, uint64_t mincode, uint64_t maxcode, uint64_t sa_offset,
                           int emit_special_tail)
{
  if (!h) return -1;
  // (the offset is derived from the bucket table when the handle has counted one)
  if (mincode > maxcode) { h->err.set("empty code range"); return -1; }
  h->full_range = false; h->mincode = mincode; h->maxcode = maxcode;
  h->given_offset = sa_offset; h->range_given = false;
  h->emit_tail = emit_special_tail ? 1 : 0;
  return 0;
}

int gtb_esa_run(gtb_esa *h, unsigned prefixlength, unsigned flags)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(h ? h->err : no_handle_err, [&]() -> int {
  if (!h) return -1;
  GTB_TRY(check_run_args(h, prefixlength, flags));
  int rc = timed_stage(h, [&]() -> int {
    GTB_TRY(h->dna ? stage_begin<true>(h, flags) : stage_begin<false>(h, flags));
    while (h->M > 0) GTB_TRY(h->dna ? round_local<true>(h) : round_local<false>(h));
    return h->dna ? stage_end<true>(h) : stage_end<false>(h);
  });
  h->ran = rc == 0;
  return rc;
  });
}

int gtb_esa_sort_begin(gtb_esa *h, unsigned prefixlength, unsigned flags)
{
  if (!h) return -1;
  GTB_TRY(check_run_args(h, prefixlength, flags));
  return timed_stage(h, [&]() -> int { return h->dna ? stage_begin<true>(h, flags) : stage_begin<false>(h, flags); });
}

int gtb_esa_slice_partition(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos,
                            const uint64_t *range_first_keys, int nranges, uint64_t *dev_keys,
                            uint32_t *dev_positions, uint64_t capacity, uint64_t *counts_out)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input) { err.set("gtb_esa_slice_partition: no input set"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  if (end_pos > h->n) end_pos = h->n;
  if (first_pos > end_pos) { err.set("gtb_esa_slice_partition: empty or reversed slice"); return -1; }
  h->fmt = choose_fmt(h, prefixlength);
  cudaEvent_t e0, e1;
  GTB_CUDA(cudaEventCreate(&e0)); GTB_CUDA(cudaEventCreate(&e1));
  GTB_CUDA(cudaEventRecord(e0, h->st));
  int rc;
  h->rw.passes = 0; h->rw.pairs_moved = 0; h->rw.launches = 0;
  if (h->dna) {
    TextSrc<true> src = make_src<true>(h, 0, ~0ull); src.pos0 = first_pos;
    rc = rs_partition_by_owner(h->rw, h->st, src, end_pos - first_pos, range_first_keys, nranges, dev_keys,
                               dev_positions, capacity, counts_out, err);
  } else {
    TextSrc<false> src = make_src<false>(h, 0, ~0ull); src.pos0 = first_pos;
    rc = rs_partition_by_owner(h->rw, h->st, src, end_pos - first_pos, range_first_keys, nranges, dev_keys,
                               dev_positions, capacity, counts_out, err);
  }
  cudaEventRecord(e1, h->st);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  h->ms_part_ext = ms; h->part_pairs_ext = h->rw.pairs_moved; h->part_launches_ext = h->rw.launches;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

int gtb_esa_sort_begin_pairs(gtb_esa *h, unsigned prefixlength, unsigned flags, const uint64_t *dev_keys,
                             const uint32_t *dev_positions, uint64_t count)
{
  if (!h) return -1;
  if (h->full_range) { h->err.set("gtb_esa_sort_begin_pairs needs a code range (gtb_esa_set_code_range)"); return -1; }
  GTB_TRY(check_run_args(h, prefixlength, flags));
  PairSrc ext{dev_keys, dev_positions};
  return timed_stage(h, [&]() -> int {
    return h->dna ? stage_begin<true>(h, flags, &ext, count) : stage_begin<false>(h, flags, &ext, count);
  });
}

int gtb_esa_sort_begin_positions(gtb_esa *h, unsigned prefixlength, unsigned flags,
                                 const uint32_t *dev_positions, uint64_t count)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (h->full_range) { err.set("gtb_esa_sort_begin_positions needs a code range"); return -1; }
  GTB_TRY(check_run_args(h, prefixlength, flags));
  // the keys are regenerated once into the second key buffer: the first pass reads them
  // there and writes the first buffer
  for (int i = 0; i < 2; i++) GTB_TRY(h->kbuf[i].ensure(sizeof(u64) * (count + 1), err));
  h->fmt = choose_fmt(h, prefixlength);
  PairSrc ext{h->kbuf[1].as<u64>(), dev_positions};
  return timed_stage(h, [&]() -> int {
    if (count > 0) {
      PhaseTimer t(h, &h->ext_ms_keygen);       // (key generation: accounted with the histogram phase)
      if (h->dna) k_keys_from_positions<true><<<grid_for(count, 256), 256, 0, h->st>>>(make_src<true>(h, 0, ~0ull), dev_positions, count, h->kbuf[1].as<u64>());
      else k_keys_from_positions<false><<<grid_for(count, 256), 256, 0, h->st>>>(make_src<false>(h, 0, ~0ull), dev_positions, count, h->kbuf[1].as<u64>());
      GTB_LAUNCH_CHECK();
      h->stats.kernel_launches++;
      t.stop();
    }
    return h->dna ? stage_begin<true>(h, flags, &ext, count) : stage_begin<false>(h, flags, &ext, count);
  });
}

uint64_t gtb_esa_unresolved(const gtb_esa *h) { return h ? h->M : 0; }

int gtb_esa_ensure_ranks(gtb_esa *h)
{
  if (!h) return -1;
  if (!h->in_progress) { h->err.set("gtb_esa_ensure_ranks outside gtb_esa_sort_begin/_end"); return -1; }
  return timed_stage(h, [&]() -> int { return h->dna ? build_ranks<true>(h) : build_ranks<false>(h); });
}

int gtb_esa_round_local(gtb_esa *h)
{
  if (!h) return -1;
  if (!h->in_progress) { h->err.set("gtb_esa_round_local outside gtb_esa_sort_begin/_end"); return -1; }
  return timed_stage(h, [&]() -> int { return h->dna ? round_local<true>(h) : round_local<false>(h); });
}

int gtb_esa_round_prepare(gtb_esa *h, const uint64_t *range_first_keys, int nranges, int my_range,
                          uint32_t *dev_send_positions, uint64_t send_capacity, uint64_t *counts_out)
{
  if (!h) return -1;
  if (!h->in_progress) { h->err.set("gtb_esa_round_prepare outside gtb_esa_sort_begin/_end"); return -1; }
  return timed_stage(h, [&]() -> int {
    return h->dna ? round_prepare<true>(h, range_first_keys, nranges, my_range, dev_send_positions, send_capacity, counts_out)
                  : round_prepare<false>(h, range_first_keys, nranges, my_range, dev_send_positions, send_capacity, counts_out);
  });
}

int gtb_esa_rank_lookup(gtb_esa *h, const uint32_t *dev_positions, uint64_t count, uint32_t *dev_ranks)
{
  if (!h) return -1;
  if (!h->in_progress || !h->isa_built) { h->err.set("gtb_esa_rank_lookup: ranks not built (gtb_esa_ensure_ranks)"); return -1; }
  return timed_stage(h, [&]() -> int {
    ErrBuf &err = h->err;
    if (count == 0) return 0;
    if (h->dna) k_rank_lookup<true><<<grid_for(count, 256), 256, 0, h->st>>>(make_rankmap<true>(h), dev_positions, count, dev_ranks);
    else k_rank_lookup<false><<<grid_for(count, 256), 256, 0, h->st>>>(make_rankmap<false>(h), dev_positions, count, dev_ranks);
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
    GTB_CUDA(cudaStreamSynchronize(h->st));
    return 0;
  });
}

int gtb_esa_round_finish(gtb_esa *h, const uint32_t *dev_answers)
{
  if (!h) return -1;
  if (!h->in_progress) { h->err.set("gtb_esa_round_finish outside gtb_esa_sort_begin/_end"); return -1; }
  return timed_stage(h, [&]() -> int { return h->dna ? round_finish<true>(h, dev_answers) : round_finish<false>(h, dev_answers); });
}

int gtb_esa_sort_end(gtb_esa *h)
{
  if (!h) return -1;
  if (!h->in_progress) { h->err.set("gtb_esa_sort_end without gtb_esa_sort_begin"); return -1; }
  int rc = timed_stage(h, [&]() -> int { return h->dna ? stage_end<true>(h) : stage_end<false>(h); });
  h->ran = rc == 0;
  return rc;
}

// the smallest filled key of bucket `code`: the first key of a code range
uint64_t gtb_code_first_key(unsigned numofchars, unsigned prefixlength, uint64_t code)
{
  if (prefixlength == 0) return 0;
  if (numofchars == 4) return code << (64 - 2 * prefixlength);
  const KeyFmt f = byte_fmt();
  u64 d[16], k = 0;
  for (int i = (int) prefixlength - 1; i >= 0; i--) { d[i] = code % numofchars; code /= numofchars; }
  for (unsigned i = 0; i < prefixlength; i++) k = (k << f.b) | d[i];
  return k << (64 - f.b * prefixlength);
}

int gtb_esa_count(gtb_esa *h, unsigned prefixlength)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input) { err.set("gtb_esa_count: no input set"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  if (prefixlength == 0 || prefixlength > (h->dna ? 15u : 7u)) { err.set("bad prefixlength %u", prefixlength); return -1; }
  if (h->pl != prefixlength) h->counted = false;
  h->fmt = choose_fmt(h, prefixlength);
  GTB_TRY(count_codes(h, prefixlength, true));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

int gtb_esa_count_partial(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input) { err.set("gtb_esa_count_partial: no input set"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  if (prefixlength == 0 || prefixlength > (h->dna ? 15u : 7u)) { err.set("bad prefixlength %u", prefixlength); return -1; }
  h->counted = false;
  h->fmt = choose_fmt(h, prefixlength);
  cudaEvent_t e0, e1;
  GTB_CUDA(cudaEventCreate(&e0)); GTB_CUDA(cudaEventCreate(&e1));
  GTB_CUDA(cudaEventRecord(e0, h->st));
  int rc = count_codes(h, prefixlength, true, first_pos, end_pos, false);
  cudaEventRecord(e1, h->st);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  h->ms_count_ext = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

int gtb_esa_count_finish(gtb_esa *h)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input || h->pl == 0 || h->ncodes == 0) { err.set("gtb_esa_count_finish without gtb_esa_count_partial"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  cudaEvent_t e0, e1;
  GTB_CUDA(cudaEventCreate(&e0)); GTB_CUDA(cudaEventCreate(&e1));
  GTB_CUDA(cudaEventRecord(e0, h->st));
  int rc = device_scan_u32(h, h->leftborder.as<u32>(), h->leftborder.as<u32>(), h->ncodes + 1, 0, nullptr, nullptr);
  cudaEventRecord(e1, h->st);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  h->ms_count_ext += ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc == 0) h->counted = true;
  return rc;
}

// ---- coarse counts: code ranges for several GPUs without a fine-grained counting pass ----
static unsigned coarse_pl(const gtb_esa *h, unsigned pl)
{
  unsigned plc = 0;
  u64 c = 1;
  while (plc < pl && c * h->K <= (u64) CC_MAXCODES) { c *= h->K; plc++; }
  return plc;
}

int gtb_esa_coarse_partial(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos,
                           uint32_t **dev_counts, uint64_t *ncounts)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input) { err.set("gtb_esa_coarse_partial: no input set"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  if (prefixlength == 0 || prefixlength > (h->dna ? 15u : 7u)) { err.set("bad prefixlength %u", prefixlength); return -1; }
  h->counted = false; h->lb_own = false;
  h->pl = prefixlength;
  h->ncodes = ipow_u64(h->K, prefixlength);
  h->fmt = choose_fmt(h, prefixlength);
  h->plc = coarse_pl(h, prefixlength);
  h->ncoarse = (u32) ipow_u64(h->K, h->plc);
  GTB_TRY(h->coarse.ensure(sizeof(u32) * (h->ncoarse + 1), err));
  cudaEvent_t e0, e1;
  GTB_CUDA(cudaEventCreate(&e0)); GTB_CUDA(cudaEventCreate(&e1));
  GTB_CUDA(cudaEventRecord(e0, h->st));
  GTB_CUDA(cudaMemsetAsync(h->coarse.p, 0, sizeof(u32) * (h->ncoarse + 1), h->st));
  if (end_pos > h->n) end_pos = h->n;
  if (first_pos < end_pos) {
    const unsigned grid = grid_for(end_pos - first_pos, 256, 148u * 8u);
    if (h->dna) k_count_coarse<true><<<grid, 256, 0, h->st>>>(make_src<true>(h, 0, ~0ull), first_pos, end_pos, h->plc, h->K, h->ncoarse, h->coarse.as<u32>());
    else k_count_coarse<false><<<grid, 256, 0, h->st>>>(make_src<false>(h, 0, ~0ull), first_pos, end_pos, h->plc, h->K, h->ncoarse, h->coarse.as<u32>());
    GTB_LAUNCH_CHECK();
  }
  cudaEventRecord(e1, h->st);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  h->ms_count_ext = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (dev_counts) *dev_counts = h->coarse.as<u32>();
  if (ncounts) *ncounts = h->ncoarse;
  return 0;
}

int gtb_esa_coarse_split(gtb_esa *h, unsigned numofparts, uint64_t *out4, unsigned *nparts)
{
  if (!h || !out4 || !nparts) return -1;
  ErrBuf &err = h->err;
  if (h->ncoarse == 0 || !h->coarse.p) { err.set("gtb_esa_coarse_split without gtb_esa_coarse_partial"); return -1; }
  if (numofparts < 1 || numofparts > (unsigned) MAX_RANGES) { err.set("gtb_esa_coarse_split: 1..%d parts", MAX_RANGES); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  // the (summed) coarse table is tiny: cut on the host, gt_suftabparts_new on coarse buckets
  std::vector<u32> cnt((size_t) h->ncoarse + 1);
  GTB_CUDA(cudaMemcpyAsync(cnt.data(), h->coarse.p, sizeof(u32) * h->ncoarse, cudaMemcpyDeviceToHost, h->st));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  std::vector<u64> cnt64(cnt.begin(), cnt.end());
  *nparts = coarse_cut(cnt64.data(), h->ncoarse, h->ncodes, numofparts, out4);
  return 0;
}

int gtb_esa_set_code_range_known(gtb_esa *h, uint64_t mincode, uint64_t maxcode, uint64_t sa_offset,
                                 uint64_t width, int emit_special_tail)
{
  if (gtb_esa_set_code_range(h, mincode, maxcode, sa_offset, emit_special_tail) != 0) return -1;
  h->range_given = true; h->given_offset = sa_offset; h->given_width = width;
  return 0;
}

int gtb_esa_dev_bcktab(const gtb_esa *h, uint32_t **leftborder, uint32_t **countspecialcodes, uint32_t **distpfxidx)
{
  if (!h || h->ncodes == 0) return -1;
  if (leftborder) *leftborder = h->leftborder.as<u32>();
  if (countspecialcodes) *countspecialcodes = h->csc.as<u32>();
  if (distpfxidx) *distpfxidx = h->dist.as<u32>();
  return 0;
}

int gtb_esa_split_ranges(gtb_esa *h, unsigned numofparts, uint64_t *out4, unsigned *nparts)
{
  if (!h || !out4 || !nparts) return -1;
  ErrBuf &err = h->err;
  if (!h->counted) { err.set("gtb_esa_split_ranges: no bucket table"); return -1; }
  if (numofparts < 1 || numofparts > (unsigned) MAX_RANGES) { err.set("gtb_esa_split_ranges: 1..%d parts", MAX_RANGES); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  GTB_TRY(h->misc.ensure(sizeof(u64) * (4 * MAX_RANGES + 8), err));
  unsigned long long *d = (unsigned long long *) h->misc.as<u64>();
  unsigned *dn = (unsigned *) (d + 4 * MAX_RANGES);
  k_split_ranges<<<1, 1, 0, h->st>>>(h->leftborder.as<u32>(), h->ncodes, numofparts, d, dn);
  GTB_LAUNCH_CHECK();
  h->stats.kernel_launches++;
  unsigned np = 0;
  GTB_CUDA(cudaMemcpyAsync(&np, dn, sizeof np, cudaMemcpyDeviceToHost, h->st));
  GTB_CUDA(cudaMemcpyAsync(out4, d, sizeof(u64) * 4 * numofparts, cudaMemcpyDeviceToHost, h->st));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  *nparts = np;
  return 0;
}

int gtb_esa_get_stats(const gtb_esa *h, gtb_stats *st)
{
  if (!h || !st) return -1;
  *st = h->stats;
  return 0;
}

uint64_t gtb_esa_num_entries(const gtb_esa *h) { return h && h->ran ? h->entries : 0; }
uint64_t gtb_esa_num_llv(const gtb_esa *h) { return h && h->ran ? h->nllv : 0; }

int gtb_esa_boundary_keys(const gtb_esa *h, uint64_t *first_key, uint64_t *last_key)
{
  if (!h || !h->ran) return -1;
  if (first_key) *first_key = h->first_key;
  if (last_key) *last_key = h->last_key;
  return 0;
}

// lcp between the last suffix of the previous shard and the first of this one
// (computelocallcpvalue of two codes, sfx-lcpvalues.c:91-111: shards meet at bucket
// borders so the filled keys decide); patches lcptab[0] and the stats.
int gtb_esa_fix_seam(gtb_esa *h, uint64_t prev_last_key)
{
  if (!h || !h->ran) return -1;
  ErrBuf &err = h->err;
  if (h->N == 0) return 0;
  GTB_CUDA(cudaSetDevice(h->device));
  const KeyFmt f = h->fmt;
  const u64 x = (prev_last_key ^ h->first_key) & f.symmask();
  u32 l = x ? (u32) (__builtin_clzll(x) / f.b) : (u32) f.m;
  const u32 ua = f.m - f.tail(prev_last_key), ub = f.m - f.tail(h->first_key);
  l = l < ua ? l : ua; l = l < ub ? l : ub;
  u8 v = (u8) l;
  GTB_CUDA(cudaMemcpyAsync(h->lcp8.p, &v, 1, cudaMemcpyHostToDevice, h->st));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  if (ub >= h->pl) h->stats.lcptabsum += l;
  if (l > h->stats.maxbranchdepth) h->stats.maxbranchdepth = l;
  return 0;
}

static int check_range(gtb_esa *h, uint64_t first, uint64_t count)
{
  if (!h) return -1;
  if (!h->ran) { h->err.set("no results: gtb_esa_run has not succeeded"); return -1; }
  if (first > h->entries || count > h->entries - first) { h->err.set("copy range out of bounds"); return -1; }
  return 0;
}

int gtb_esa_copy_suftab_u32(gtb_esa *h, uint32_t *dst, uint64_t first, uint64_t count)
{
  GTB_TRY(check_range(h, first, count));
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  const u32 *src = h->vbuf[h->res].as<u32>() + first;
  if (count && host_pointer_is_pinned(dst))
    GTB_CUDA(cudaMemcpyAsync(dst, src, sizeof(u32) * count, cudaMemcpyDeviceToHost, h->st));
  else
    GTB_TRY(staged_d2h(h->hstage, h->st, src, dst, count, sizeof(u32), false, err));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

// the wide half of gtb_esa_copy_suftab_u64: chunks from the back of the dealer are widened on the
// device (into the dead key buffers) and copied as uint64 straight into the caller's pinned buffer.
// grace_ms > 0: first give the narrow path that long; if it has dealt itself at least min_front chunks
// by then it is faster than this path could be and keeps the whole table
static int suftab_wide_chunks(gtb_esa *h, cudaStream_t st2, uint64_t *dst, uint64_t first, uint64_t count,
                              u64 per, ChunkDealer *dealer, ErrBuf &err, int grace_ms, u64 min_front)
{
  GTB_CUDA(cudaSetDevice(h->device));
  if (grace_ms > 0) {
    std::this_thread::sleep_for(std::chrono::milliseconds(grace_ms));
    if (dealer->front_taken() >= min_front) return 0;
  }
  u64 *stage[2];
  cudaEvent_t done[2];
  for (int i = 0; i < 2; i++) {
    stage[i] = h->kbuf[i].as<u64>();
    GTB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  int b = 0, rc = 0;
  u64 k;
  while (rc == 0 && dealer->take_back(&k)) {
    const u64 off = k * per, c = count - off < per ? count - off : per;
    cudaError_t e = cudaEventSynchronize(done[b]);          // (a fresh event is complete)
    if (e == cudaSuccess) {
      k_widen_u32_u64<<<grid_for(c, 256), 256, 0, st2>>>(h->vbuf[h->res].as<u32>() + first + off, stage[b], c);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst + off, stage[b], sizeof(u64) * c, cudaMemcpyDeviceToHost, st2);
    if (e == cudaSuccess) e = cudaEventRecord(done[b], st2);
    if (e != cudaSuccess) { err.set("wide suffix-table copy failed: %s", cudaGetErrorString(e)); rc = -1; }
    b ^= 1;
  }
  if (cudaStreamSynchronize(st2) != cudaSuccess && rc == 0) { err.set("wide suffix-table copy failed"); rc = -1; }
  for (int i = 0; i < 2; i++) cudaEventDestroy(done[i]);
  return rc;
}

int gtb_esa_copy_suftab_u64(gtb_esa *h, uint64_t *dst, uint64_t first, uint64_t count)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(h ? h->err : no_handle_err, [&]() -> int {
  GTB_TRY(check_range(h, first, count));
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  // The .suf entries are uint64, the table in HBM uint32.  Two ways to the host: "narrow" -- 4 bytes
  // per entry cross PCIe into small pinned staging buffers (they stay in the host's last level cache)
  // and host threads widen them into the destination: bound by the host's cores (measured: 334-400 ms
  // for c4's 24.8 GB with 14 threads; 1025 ms on a host whose cores stream slowly; 8 ranks sharing one
  // host starve each other); "wide" -- widened on the device, 8 bytes per entry by DMA straight into
  // a pinned destination: bound by PCIe (463 ms for c4 on one GPU, but it scales with the GPUs).
  // Default (GTB200_SUF_COPY unset or "auto"): the narrow path starts; after a few milliseconds the
  // wide path looks at how far it got -- slower than PCIe would be, and the wide path takes chunks from
  // the other end of the table until the two meet.  GTB200_SUF_COPY = narrow | wide | both force a way.
  const char *mode = getenv("GTB200_SUF_COPY");
  const bool pinned = host_pointer_is_pinned(dst);
  const bool m_wide = mode && strcmp(mode, "wide") == 0, m_both = mode && strcmp(mode, "both") == 0;
  const bool m_narrow = mode && strcmp(mode, "narrow") == 0;
  GTB_TRY(h->hstage.ensure(err));
  const u64 per = h->hstage.chunk / sizeof(u32);
  const u64 nchunks = div_up(count, per);
  const bool want_wide = pinned && !m_narrow && nchunks >= (m_wide || m_both ? 4u : 128u);
  const bool want_narrow = !(pinned && m_wide);
  if (!want_wide) {
    GTB_TRY(staged_d2h(h->hstage, h->st, h->vbuf[h->res].as<u32>() + first, dst, count, sizeof(u32), true, err));
    GTB_CUDA(cudaStreamSynchronize(h->st));
    return 0;
  }
  for (int i = 0; i < 2; i++) GTB_TRY(h->kbuf[i].ensure(sizeof(u64) * per, err));
  if (!h->st2) GTB_CUDA(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
  GTB_CUDA(cudaStreamSynchronize(h->st));                  // the results are final; both streams may read them
  ChunkDealer dealer(nchunks);
  ErrBuf err2;
  int rc2 = 0;
  // auto: 12 ms of grace; PCIe moves about 6.6 G entries/s the wide way -- the narrow path must beat 85 % of it
  const int grace_ms = (m_wide || m_both) ? 0 : 12;
  const u64 min_front = (u64) (0.85 * 6.6e9 * 0.012 / (double) per);
  std::thread wide([&] { rc2 = suftab_wide_chunks(h, h->st2, dst, first, count, per, &dealer, err2, grace_ms, min_front); });
  int rc = 0;
  if (want_narrow)
    rc = staged_d2h(h->hstage, h->st, h->vbuf[h->res].as<u32>() + first, dst, count, sizeof(u32), true, err, &dealer);
  wide.join();
  if (rc == 0 && rc2 != 0) { err = err2; rc = -1; }
  GTB_TRY(rc);
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
  });
}

int gtb_esa_copy_lcptab(gtb_esa *h, uint8_t *dst, uint64_t first, uint64_t count)
{
  GTB_TRY(check_range(h, first, count));
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  const u8 *src = h->lcp8.as<u8>() + first;
  if (count && host_pointer_is_pinned(dst))
    GTB_CUDA(cudaMemcpyAsync(dst, src, count, cudaMemcpyDeviceToHost, h->st));
  else
    GTB_TRY(staged_d2h(h->hstage, h->st, src, dst, count, 1, false, err));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

// suffix table and lcp table together: the suffix-table copy is bound by the host's memory system,
// not by the bus, so the lcp bytes travel on a second stream at the same time
int gtb_esa_copy_tables(gtb_esa *h, uint64_t *suftab, uint8_t *lcptab, uint64_t first, uint64_t count)
{
  GTB_TRY(check_range(h, first, count));
  if (!suftab || !lcptab || count == 0 || !host_pointer_is_pinned(lcptab)) {
    if (suftab) GTB_TRY(gtb_esa_copy_suftab_u64(h, suftab, first, count));
    if (lcptab) GTB_TRY(gtb_esa_copy_lcptab(h, lcptab, first, count));
    return 0;
  }
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  if (!h->st3) GTB_CUDA(cudaStreamCreateWithFlags(&h->st3, cudaStreamNonBlocking));
  GTB_CUDA(cudaStreamSynchronize(h->st));                  // the results are final
  GTB_CUDA(cudaMemcpyAsync(lcptab, h->lcp8.as<u8>() + first, count, cudaMemcpyDeviceToHost, h->st3));
  const int rc = gtb_esa_copy_suftab_u64(h, suftab, first, count);
  GTB_CUDA(cudaStreamSynchronize(h->st3));
  return rc;
}

// everything suffixeratorwithoutput() writes, in one call: the small tables (lcp, llv, bucket table)
// cross the bus on a second stream while the suffix table is copied and widened.  Any pointer may be
// NULL; pageable destinations are served one after the other.
int gtb_esa_copy_results(gtb_esa *h, uint64_t *suftab, uint8_t *lcptab, uint64_t *llv,
                         uint32_t *leftborder, uint32_t *countspecialcodes, uint32_t *distpfxidx)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(h ? h->err : no_handle_err, [&]() -> int {
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->ran) { err.set("no results: gtb_esa_run has not succeeded"); return -1; }
  const u64 e = h->entries;
  const bool want_bck = leftborder || countspecialcodes || distpfxidx;
  if (want_bck && !h->counted && !h->lb_own) { err.set("no bucket table: run with GTB_WANT_BCK or call gtb_esa_count first"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  if (!h->st3) GTB_CUDA(cudaStreamCreateWithFlags(&h->st3, cudaStreamNonBlocking));
  GTB_CUDA(cudaStreamSynchronize(h->st));                  // the results are final
  // what the copy engine can write directly goes to the second stream now ...
  bool later_lcp = false, later_llv = false, later_lb = false, later_csc = false, later_dist = false;
  auto side = [&](void *dst, const void *src, size_t bytes, bool *later) -> int {
    if (!dst || bytes == 0) return 0;
    if (!host_pointer_is_pinned(dst)) { *later = true; return 0; }
    GTB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st3));
    return 0;
  };
  GTB_TRY(side(lcptab, h->lcp8.p, e, &later_lcp));
  GTB_TRY(side(llv, h->llv.p, sizeof(u64) * 2 * h->nllv, &later_llv));
  if (want_bck) {
    GTB_TRY(side(leftborder, h->leftborder.p, sizeof(u32) * (h->ncodes + 1), &later_lb));
    GTB_TRY(side(countspecialcodes, h->csc.p, sizeof(u32) * h->nspecialcodes, &later_csc));
    GTB_TRY(side(distpfxidx, h->dist.p, sizeof(u32) * h->ndist, &later_dist));
  }
  // ... while the suffix table takes the main path
  int rc = suftab ? gtb_esa_copy_suftab_u64(h, suftab, 0, e) : 0;
  if (cudaStreamSynchronize(h->st3) != cudaSuccess && rc == 0) { err.set("result copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
  GTB_TRY(rc);
  if (later_lcp) GTB_TRY(gtb_esa_copy_lcptab(h, lcptab, 0, e));
  if (later_llv) GTB_TRY(gtb_esa_copy_llv(h, llv));
  if (later_lb || later_csc || later_dist)
    GTB_TRY(gtb_esa_copy_bcktab(h, later_lb ? leftborder : nullptr, later_csc ? countspecialcodes : nullptr,
                                later_dist ? distpfxidx : nullptr));
  return 0;
  });
}

// the early copy of gtb_esa_run_to_host is over (or never began): join, free
static int finish_early_copy(gtb_esa *h)
{
  int rc = 0;
  if (h->ec.started) {
    if (h->ec.th.joinable()) h->ec.th.join();
    rc = h->ec.rc;
    if (rc != 0) h->err = h->ec.err;
    delete h->ec.dealer;
    h->ec.dealer = nullptr;
    h->ec.started = false;
  }
  h->ec.dst = nullptr;
  return rc;
}

// gtb_esa_run + gtb_esa_copy_results in one call, overlapped: what suffixeratorwithoutput() (sfx-run.c:212-317)
// does with the iterator -- sort, then write every table -- with the suffix table already on its way to the
// host while the ties are refined.  Host buffers as gtb_esa_copy_results; llv must hold 2 * llv_capacity uint64.
int gtb_esa_run_to_host(gtb_esa *h, unsigned prefixlength, unsigned flags, uint64_t *suftab, uint8_t *lcptab,
                        uint64_t *llv, uint64_t llv_capacity, uint64_t *nllv, uint32_t *leftborder,
                        uint32_t *countspecialcodes, uint32_t *distpfxidx)
{
  if (!h) return -1;
  static thread_local ErrBuf no_handle_err;
  return no_throw(h->err, [&]() -> int {
    ErrBuf &err = h->err;
    const char *off = getenv("GTB200_NO_EARLY_COPY");
    h->ec.dst = (suftab && (flags & GTB_WANT_SUF) && !off) ? suftab : nullptr;
    int rc = gtb_esa_run(h, prefixlength, flags);
    const bool early = h->ec.started;
    if (rc != 0) { finish_early_copy(h); return -1; }
    GTB_CUDA(cudaSetDevice(h->device));
    if (nllv) *nllv = h->nllv;
    if (llv && h->nllv > llv_capacity) {
      finish_early_copy(h);
      err.set("llv buffer too small: %llu entries needed", (unsigned long long) h->nllv);
      return -1;
    }
    if (!early) {
      h->ec.dst = nullptr;
      return gtb_esa_copy_results(h, suftab, lcptab, h->nllv ? llv : nullptr, leftborder, countspecialcodes, distpfxidx);
    }
    // every chunk that has not been handed out yet will read the final table; the others are patched
    const u64 N = h->N, per = h->ec.per;
    u64 limit = h->ec.dealer->front_taken() * per;
    if (limit > N) limit = N;
    unsigned int npatch = 0;
    if (h->M0 > 0 && limit > 0) {
      GTB_TRY(h->dkeys.ensure(sizeof(u64) * h->M0, err));
      unsigned int *dcount = reinterpret_cast<unsigned int *>(h->misc.as<u64>() + 22);
      GTB_CUDA(cudaMemsetAsync(dcount, 0, sizeof(unsigned int), h->st));
      k_patch_gather<<<grid_for(h->M0, 256), 256, 0, h->st>>>(h->uidx0.as<u32>(), h->M0, h->vbuf[h->res].as<u32>(), limit,
                                                             h->dkeys.as<u64>(), dcount);
      if (cudaGetLastError() != cudaSuccess) { finish_early_copy(h); err.set("k_patch_gather failed to launch"); return -1; }
      GTB_CUDA(cudaMemcpyAsync(&npatch, dcount, sizeof npatch, cudaMemcpyDeviceToHost, h->st));
      GTB_CUDA(cudaStreamSynchronize(h->st));
      if (npatch > 0) {
        if (h->patch_cap < sizeof(u64) * (size_t) npatch) {
          if (h->patch_host) cudaFreeHost(h->patch_host);
          h->patch_host = nullptr; h->patch_cap = 0;
          const size_t want = sizeof(u64) * (size_t) npatch + (sizeof(u64) * (size_t) npatch >> 3) + 4096;
          if (cudaHostAlloc(&h->patch_host, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError(); finish_early_copy(h); err.set("no pinned memory for %u patches", npatch); return -1;
          }
          h->patch_cap = want;
        }
        GTB_CUDA(cudaMemcpyAsync(h->patch_host, h->dkeys.p, sizeof(u64) * (size_t) npatch, cudaMemcpyDeviceToHost, h->st));
      }
    }
    rc = finish_early_copy(h);                       // the table [0, N) is on the host
    if (rc != 0) return -1;
    GTB_CUDA(cudaStreamSynchronize(h->st));          // ... and so are the patches
    if (npatch > 0) {
      const u64 *pt = static_cast<const u64 *>(h->patch_host);
      int T = h->hstage.nthreads > 0 ? h->hstage.nthreads : 1;
      if (npatch < 65536u) T = 1;
      std::vector<std::thread> pool;
      auto work = [&](int t) {
        const u64 lo = (u64) npatch * (u64) t / (u64) T, hi = (u64) npatch * (u64) (t + 1) / (u64) T;
        for (u64 i = lo; i < hi; i++) suftab[pt[i] >> 32] = pt[i] & 0xffffffffull;
      };
      pool.reserve((size_t) T);
      for (int t = 1; t < T; t++) pool.emplace_back(work, t);
      work(0);
      for (auto &th : pool) th.join();
    }
    // the special tail and the other tables
    if (h->entries > N) GTB_TRY(gtb_esa_copy_suftab_u64(h, suftab + N, N, h->entries - N));
    return gtb_esa_copy_results(h, nullptr, lcptab, h->nllv ? llv : nullptr, leftborder, countspecialcodes, distpfxidx);
  });
}

int gtb_esa_set_separators(gtb_esa *h, const uint64_t *positions, uint64_t count)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->have_input || !h->dna) { err.set("gtb_esa_set_separators: needs a 2-bit input (the byte path carries its separators)"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  for (u64 i = 0; i < count; i++)
    if (positions[i] >= h->n || (i > 0 && positions[i] <= positions[i - 1])) {
      err.set("separator %llu at %llu is unordered or beyond the text", (unsigned long long) i, (unsigned long long) positions[i]);
      return -1;
    }
  u64 *mirrored = nullptr;
  if ((h->readmode & 1u) && count > 0) {          // forward coordinates -> read direction
    mirrored = static_cast<u64 *>(malloc(sizeof(u64) * count));
    if (!mirrored) { err.set("out of host memory for %llu separators", (unsigned long long) count); return -1; }
    for (u64 i = 0; i < count; i++) mirrored[i] = h->n - 1 - positions[count - 1 - i];
    positions = mirrored;
  }
  struct FreeLater { u64 *p; ~FreeLater() { free(p); } } free_later{mirrored};
  const u64 nw = (h->n >> 5) + 2;
  GTB_TRY(h->sepbits.ensure(sizeof(u32) * nw, err));
  GTB_CUDA(cudaMemsetAsync(h->sepbits.p, 0, sizeof(u32) * nw, h->st));
  if (count > 0) {
    GTB_TRY(h->seppos.ensure(sizeof(u64) * count, err));
    GTB_CUDA(cudaMemcpyAsync(h->seppos.p, positions, sizeof(u64) * count, cudaMemcpyHostToDevice, h->st));
    k_set_bits<<<grid_for(count, 256), 256, 0, h->st>>>(h->seppos.as<u64>(), count, h->sepbits.as<u32>());
    GTB_LAUNCH_CHECK();
  }
  GTB_CUDA(cudaStreamSynchronize(h->st));
  h->have_sep = true;
  return 0;
}

int gtb_esa_copy_bwttab(gtb_esa *h, uint8_t *dst, uint64_t first, uint64_t count)
{
  GTB_TRY(check_range(h, first, count));
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  if (count == 0) return 0;
  // staged in the (dead) key buffer, copied in chunks
  const u64 chunk = 1ull << 28;
  GTB_TRY(h->kbuf[h->res ^ 1].ensure(chunk < count ? chunk : count, err));
  u8 *stage = h->kbuf[h->res ^ 1].as<u8>();
  for (u64 off = 0; off < count; off += chunk) {
    const u64 c = count - off < chunk ? count - off : chunk;
    const u32 *sa = h->vbuf[h->res].as<u32>() + first + off;
    const u32 *sep = h->have_sep ? h->sepbits.as<u32>() : nullptr;
    if (h->dna) k_bwt<true><<<grid_for(c, 256), 256, 0, h->st>>>(sa, c, h->words.as<u64>(), h->bytes.as<u8>(), h->spmask.as<u32>(), sep, stage);
    else k_bwt<false><<<grid_for(c, 256), 256, 0, h->st>>>(sa, c, h->words.as<u64>(), h->bytes.as<u8>(), h->spmask.as<u32>(), sep, stage);
    GTB_LAUNCH_CHECK();
    h->stats.kernel_launches++;
    GTB_CUDA(cudaMemcpyAsync(dst + off, stage, c, cudaMemcpyDeviceToHost, h->st));
    GTB_CUDA(cudaStreamSynchronize(h->st));
  }
  return 0;
}

int gtb_esa_copy_llv(gtb_esa *h, uint64_t *dst)
{
  if (!h || !h->ran) return -1;
  ErrBuf &err = h->err;
  GTB_CUDA(cudaSetDevice(h->device));
  if (h->nllv) GTB_CUDA(cudaMemcpyAsync(dst, h->llv.p, sizeof(u64) * 2 * h->nllv, cudaMemcpyDeviceToHost, h->st));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

int gtb_esa_copy_bcktab(gtb_esa *h, uint32_t *leftborder, uint32_t *countspecialcodes, uint32_t *distpfxidx)
{
  if (!h) return -1;
  ErrBuf &err = h->err;
  if (!h->counted && !h->lb_own) { err.set("no bucket table: run with GTB_WANT_BCK or call gtb_esa_count first"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  // (pageable destinations go through the pinned staging buffers and host threads)
  auto table = [&](uint32_t *dst, const void *src, u64 cnt) -> int {
    if (!dst || cnt == 0) return 0;
    if (host_pointer_is_pinned(dst)) GTB_CUDA(cudaMemcpyAsync(dst, src, sizeof(u32) * cnt, cudaMemcpyDeviceToHost, h->st));
    else GTB_TRY(staged_d2h(h->hstage, h->st, src, dst, cnt, sizeof(u32), false, err));
    return 0;
  };
  GTB_TRY(table(leftborder, h->leftborder.p, h->ncodes + 1));
  GTB_TRY(table(countspecialcodes, h->csc.p, h->nspecialcodes));
  GTB_TRY(table(distpfxidx, h->dist.p, h->ndist));
  GTB_CUDA(cudaStreamSynchronize(h->st));
  return 0;
}

int gtb_esa_hash_results(gtb_esa *h, uint64_t llv_pairs_before, uint64_t out3[3])
{
  if (!h || !out3) return -1;
  ErrBuf &err = h->err;
  if (!h->ran) { err.set("no results: gtb_esa_run has not succeeded"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  out3[0] = out3[1] = out3[2] = 0;
  GTB_TRY(hash_table(h, h->vbuf[h->res].as<u32>(), h->entries, h->sa_offset, &out3[0]));
  if (h->flags & GTB_WANT_LCP) {
    GTB_TRY(hash_table(h, h->lcp8.as<u8>(), h->entries, h->sa_offset, &out3[1]));
    GTB_TRY(hash_table(h, h->llv.as<u64>(), 2 * h->nllv, 2 * llv_pairs_before, &out3[2]));
  }
  return 0;
}

int gtb_esa_hash_bcktab(gtb_esa *h, uint64_t *out)
{
  if (!h || !out) return -1;
  ErrBuf &err = h->err;
  if (!h->counted && !h->lb_own) { err.set("no bucket table: run with GTB_WANT_BCK or call gtb_esa_count first"); return -1; }
  GTB_CUDA(cudaSetDevice(h->device));
  // the .bck file: three uint32 tables, each padded to 8 bytes (gt_mapspec_write, mapspec.c:350-365)
  u64 acc = 0, word = 0;
  const u64 counts[3] = {h->ncodes + 1, h->nspecialcodes, h->ndist};
  const u32 *tabs[3] = {h->leftborder.as<u32>(), h->csc.as<u32>(), h->dist.as<u32>()};
  for (int t = 0; t < 3; t++) {
    GTB_TRY(hash_table(h, tabs[t], counts[t], word, &acc));
    word += counts[t];
    if (word & 1ull) { acc += mh_term(word, 0); word++; }
  }
  *out = acc;
  return 0;
}

// ---- one job sharded over several code ranges (gtb_shard_host.cuh) ----
int gtb_esa_run_sharded(gtb_esa *h, unsigned prefixlength, unsigned flags, int rank, int world,
                        gtb_allgather_fn allgather, void *ctx, int separate_processes)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(h ? h->err : no_handle_err, [&]() -> int {
  if (!h) return -1;
  if (!allgather) { h->err.set("gtb_esa_run_sharded needs an all-gather function"); return -1; }
  if (!h->have_input) { h->err.set("no input set"); return -1; }
  ShardComm c;
  c.me = rank; c.world = world; c.ag = allgather; c.ctx = ctx; c.ipc = separate_processes != 0;
  h->ran = false;
  return run_sharded(h, c, prefixlength, flags);
  });
}

uint64_t gtb_esa_llv_before(const gtb_esa *h) { return h && h->ran ? h->llv_before : 0; }

gtb_group *gtb_group_new(const int *devices, int nranges, char *errbuf, size_t errlen)
{
  if (nranges < 1 || nranges > MAX_RANGES || !devices) {
    if (errbuf && errlen) snprintf(errbuf, errlen, "gtb_group_new: 1..%d code ranges", MAX_RANGES);
    return nullptr;
  }
  gtb_group *g = new (std::nothrow) gtb_group();
  if (!g) { if (errbuf && errlen) snprintf(errbuf, errlen, "out of host memory"); return nullptr; }
  memset(&g->stats, 0, sizeof g->stats);
  for (int i = 0; i < nranges; i++) {
    gtb_esa *h = gtb_esa_new(devices[i], errbuf, errlen);
    if (!h) { gtb_group_delete(g); return nullptr; }
    h->hstage.thread_div = nranges;       // the host threads of the result copies are shared by the ranges
    g->hs.push_back(h);
  }
  g->comm.world = nranges;
  return g;
}

void gtb_group_delete(gtb_group *g)
{
  if (!g) return;
  // borrowers first: they only drop their references to the input of the first handle of their device
  for (size_t i = g->hs.size(); i-- > 0;) gtb_esa_delete(g->hs[i]);
  delete g;
}

const char *gtb_group_error(const gtb_group *g) { return g ? g->err.msg : "null group"; }
int gtb_group_size(const gtb_group *g) { return g ? (int) g->hs.size() : 0; }
gtb_esa *gtb_group_range(gtb_group *g, int i) { return g && i >= 0 && i < (int) g->hs.size() ? g->hs[(size_t) i] : nullptr; }

int gtb_group_set_readmode(gtb_group *g, unsigned readmode)
{
  if (!g) return -1;
  for (gtb_esa *h : g->hs)
    if (gtb_esa_set_readmode(h, readmode) != 0) { snprintf(g->err.msg, sizeof g->err.msg, "%s", h->err.msg); return -1; }
  return 0;
}

int gtb_group_set_input_2bit(gtb_group *g, const uint64_t *twobitenc, uint64_t nwords, uint64_t n,
                             const gtb_range *specials, uint64_t nranges)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g) return -1;
  return group_set_input(g, [&](gtb_esa *h) { return gtb_esa_set_input_2bit(h, twobitenc, nwords, n, specials, nranges); });
  });
}

int gtb_group_set_input_bytes(gtb_group *g, const uint8_t *symbols, uint64_t n, unsigned K)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g) return -1;
  return group_set_input(g, [&](gtb_esa *h) { return gtb_esa_set_input_bytes(h, symbols, n, K); });
  });
}

int gtb_group_set_separators(gtb_group *g, const uint64_t *positions, uint64_t count)
{
  if (!g) return -1;
  for (size_t i = 0; i < g->hs.size(); i++) {
    gtb_esa *h = g->hs[i], *owner = nullptr;
    for (size_t j = 0; j < i; j++) if (g->hs[j]->device == h->device) { owner = g->hs[j]; break; }
    int rc;
    if (!owner) rc = gtb_esa_set_separators(h, positions, count);
    else { h->sepbits.borrow(owner->sepbits); h->have_sep = owner->have_sep; rc = 0; }
    if (rc != 0) { snprintf(g->err.msg, sizeof g->err.msg, "%s", h->err.msg); return -1; }
  }
  return 0;
}

int gtb_group_run(gtb_group *g, unsigned prefixlength, unsigned flags)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g) return -1;
  g->ran = false; g->bck_merged = false;
  g->pl = prefixlength; g->flags = flags;
  const int n = (int) g->hs.size();
  std::vector<LocalCtx> ctx((size_t) n);
  for (int i = 0; i < n; i++) ctx[(size_t) i] = LocalCtx{&g->comm, i};
  GTB_TRY(group_parallel(g, [&](int i) -> int {
    return gtb_esa_run_sharded(g->hs[(size_t) i], prefixlength, flags, i, n, local_allgather, &ctx[(size_t) i], 0);
  }));
  // the job's numbers: what gt_Sfxiterator_longest and the GtOutlcpinfo getters report (sfx-run.c:300,676-680)
  gtb_stats &S = g->stats;
  memset(&S, 0, sizeof S);
  S.longest = ~0ull;
  for (gtb_esa *h : g->hs) {
    const gtb_stats &t = h->stats;
    S.totallength = t.totallength; S.specialcharacters = t.specialcharacters;
    S.prefixlength = prefixlength; S.numofchars = h->K;
    S.nonspecials += t.nonspecials;
    if (t.longest != ~0ull) S.longest = t.longest;
    S.numoflargelcpvalues += t.numoflargelcpvalues;
    if (t.maxbranchdepth > S.maxbranchdepth) S.maxbranchdepth = t.maxbranchdepth;
    S.lcptabsum += t.lcptabsum;
    S.unresolved_after_first_sort += t.unresolved_after_first_sort;
    if (t.doubling_rounds > S.doubling_rounds) S.doubling_rounds = t.doubling_rounds;
    S.radix_passes += t.radix_passes; S.radix_pairs_moved += t.radix_pairs_moved;
    S.radix_passes_first += t.radix_passes_first; S.radix_pairs_first += t.radix_pairs_first;
    if (t.ms_radix_first > S.ms_radix_first) S.ms_radix_first = t.ms_radix_first;
    S.kernel_launches += t.kernel_launches;
    float *dst[] = {&S.ms_total, &S.ms_upload, &S.ms_count, &S.ms_hist, &S.ms_radix, &S.ms_analyze, &S.ms_doubling, &S.ms_lcp, &S.ms_tail};
    const float src[] = {t.ms_total, t.ms_upload, t.ms_count, t.ms_hist, t.ms_radix, t.ms_analyze, t.ms_doubling, t.ms_lcp, t.ms_tail};
    for (int k = 0; k < 9; k++) if (src[k] > *dst[k]) *dst[k] = src[k];     // device times: the slowest range
  }
  g->ran = true;
  return 0;
  });
}

int gtb_group_get_stats(const gtb_group *g, gtb_stats *st)
{
  if (!g || !st || !g->ran) return -1;
  *st = g->stats;
  return 0;
}

uint64_t gtb_group_num_entries(const gtb_group *g)
{
  u64 e = 0;
  if (g && g->ran) for (gtb_esa *h : g->hs) e += h->entries;
  return e;
}
uint64_t gtb_group_num_llv(const gtb_group *g)
{
  u64 e = 0;
  if (g && g->ran) for (gtb_esa *h : g->hs) e += h->nllv;
  return e;
}

// the result gather: every range copies its shard straight to its place in the caller's tables
// (offset sa_offset of the global suffix / lcp table, its pairs of the .llv list), all at once
int gtb_group_copy_results(gtb_group *g, uint64_t *suftab, uint8_t *lcptab, uint64_t *llv,
                           uint32_t *leftborder, uint32_t *countspecialcodes, uint32_t *distpfxidx)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g) return -1;
  if (!g->ran) { g->err.set("no results: gtb_group_run has not succeeded"); return -1; }
  if (g->hs.size() == 1)
    return group_parallel(g, [&](int) { return gtb_esa_copy_results(g->hs[0], suftab, lcptab, llv, leftborder, countspecialcodes, distpfxidx); });
  GTB_TRY(group_parallel(g, [&](int i) -> int {
    gtb_esa *h = g->hs[(size_t) i];
    if (h->entries == 0) return 0;
    return gtb_esa_copy_results(h, suftab ? suftab + h->sa_offset : nullptr, lcptab ? lcptab + h->sa_offset : nullptr,
                                (llv && h->nllv) ? llv + 2 * h->llv_before : nullptr, nullptr, nullptr, nullptr);
  }));
  if (leftborder || countspecialcodes || distpfxidx) {
    GTB_TRY(group_merge_bck(g));
    if (gtb_esa_copy_bcktab(g->hs[0], leftborder, countspecialcodes, distpfxidx) != 0) {
      snprintf(g->err.msg, sizeof g->err.msg, "%s", g->hs[0]->err.msg); return -1;
    }
  }
  return 0;
  });
}

int gtb_group_copy_bwttab(gtb_group *g, uint8_t *dst)
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g || !dst) return -1;
  if (!g->ran) { g->err.set("no results: gtb_group_run has not succeeded"); return -1; }
  return group_parallel(g, [&](int i) -> int {
    gtb_esa *h = g->hs[(size_t) i];
    return h->entries ? gtb_esa_copy_bwttab(h, dst + (g->hs.size() == 1 ? 0 : h->sa_offset), 0, h->entries) : 0;
  });
  });
}

int gtb_group_hash_results(gtb_group *g, uint64_t out4[4])
{
  static thread_local ErrBuf no_handle_err;
  return no_throw(g ? g->err : no_handle_err, [&]() -> int {
  if (!g || !out4) return -1;
  if (!g->ran) { g->err.set("no results: gtb_group_run has not succeeded"); return -1; }
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  for (gtb_esa *h : g->hs) {
    u64 o[3];
    if (gtb_esa_hash_results(h, g->hs.size() == 1 ? 0 : h->llv_before, o) != 0) { snprintf(g->err.msg, sizeof g->err.msg, "%s", h->err.msg); return -1; }
    for (int k = 0; k < 3; k++) out4[k] += o[k];
  }
  GTB_TRY(group_merge_bck(g));
  if (gtb_esa_hash_bcktab(g->hs[0], &out4[3]) != 0) { snprintf(g->err.msg, sizeof g->err.msg, "%s", g->hs[0]->err.msg); return -1; }
  return 0;
  });
}

void *gtb_esa_stream(const gtb_esa *h) { return h ? (void *) h->st : nullptr; }

const uint32_t *gtb_esa_dev_suftab(const gtb_esa *h) { return h && h->ran ? h->vbuf[h->res].as<u32>() : nullptr; }
const uint8_t *gtb_esa_dev_lcptab(const gtb_esa *h) { return h && h->ran ? h->lcp8.as<u8>() : nullptr; }
const uint32_t *gtb_esa_dev_leftborder(const gtb_esa *h) { return h && h->counted ? h->leftborder.as<u32>() : nullptr; }

static int one_shot(gtb_esa *h, unsigned pl, uint64_t *suftab, uint8_t *lcptab, uint64_t *llv,
                    uint64_t llv_capacity, uint64_t *nllv, uint32_t *leftborder, uint32_t *csc,
                    uint32_t *dist, gtb_stats *stats)
{
  unsigned flags = 0;
  if (suftab) flags |= GTB_WANT_SUF;
  if (lcptab || llv) flags |= GTB_WANT_LCP;
  if (leftborder || csc || dist) flags |= GTB_WANT_BCK;
  GTB_TRY(gtb_esa_run(h, pl, flags));
  const u64 e = gtb_esa_num_entries(h);
  GTB_TRY(gtb_esa_copy_tables(h, suftab, lcptab, 0, e));
  if (nllv) *nllv = h->nllv;
  if (llv) {
    if (h->nllv > llv_capacity) { h->err.set("llv buffer too small: %llu entries needed", (unsigned long long) h->nllv); return -1; }
    GTB_TRY(gtb_esa_copy_llv(h, llv));
  }
  if (flags & GTB_WANT_BCK) GTB_TRY(gtb_esa_copy_bcktab(h, leftborder, csc, dist));
  if (stats) *stats = h->stats;
  return 0;
}

int gtb_esa_build_2bit(int device, const uint64_t *twobitenc, uint64_t nwords, uint64_t n,
                       const gtb_range *specials, uint64_t nranges, unsigned pl,
                       uint64_t *suftab, uint8_t *lcptab, uint64_t *llv, uint64_t llv_capacity,
                       uint64_t *nllv, uint32_t *leftborder, uint32_t *csc, uint32_t *dist,
                       gtb_stats *stats, char *errbuf, size_t errlen)
{
  gtb_esa *h = gtb_esa_new(device, errbuf, errlen);
  if (!h) return -1;
  int rc = gtb_esa_set_input_2bit(h, twobitenc, nwords, n, specials, nranges);
  if (rc == 0) rc = one_shot(h, pl, suftab, lcptab, llv, llv_capacity, nllv, leftborder, csc, dist, stats);
  if (rc != 0 && errbuf && errlen) snprintf(errbuf, errlen, "%s", h->err.msg);
  gtb_esa_delete(h);
  return rc;
}

int gtb_esa_build_bytes(int device, const uint8_t *symbols, uint64_t n, unsigned K, unsigned pl,
                        uint64_t *suftab, uint8_t *lcptab, uint64_t *llv, uint64_t llv_capacity,
                        uint64_t *nllv, uint32_t *leftborder, uint32_t *csc, uint32_t *dist,
                        gtb_stats *stats, char *errbuf, size_t errlen)
{
  gtb_esa *h = gtb_esa_new(device, errbuf, errlen);
  if (!h) return -1;
  int rc = gtb_esa_set_input_bytes(h, symbols, n, K);
  if (rc == 0) rc = one_shot(h, pl, suftab, lcptab, llv, llv_capacity, nllv, leftborder, csc, dist, stats);
  if (rc != 0 && errbuf && errlen) snprintf(errbuf, errlen, "%s", h->err.msg);
  gtb_esa_delete(h);
  return rc;
}

int gtb_radixsort_pairs_u64_u32(int device, uint64_t *keys, uint32_t *values, uint64_t count,
                                unsigned begin_bit, unsigned end_bit, char *errbuf, size_t errlen)
{
  ErrBuf err;
  auto body = [&]() -> int {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      err.set("libgtb200: no CUDA device available; there is no CPU fallback"); return -1;
    }
    if (begin_bit >= end_bit || end_bit > 64) { err.set("bad bit range"); return -1; }
    GTB_CUDA(cudaSetDevice(device));
    if (count == 0) return 0;
    cudaStream_t st;
    GTB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    RadixWork rw;
    GTB_TRY(radix_work_init(rw, err));
    DevBuf k[2], v[2], kin, vin;
    int rc = 0;
    for (int i = 0; i < 2 && rc == 0; i++) { rc |= k[i].ensure(sizeof(u64) * count, err); rc |= v[i].ensure(sizeof(u32) * count, err); }
    if (rc == 0) { rc |= kin.ensure(sizeof(u64) * count, err); rc |= vin.ensure(sizeof(u32) * count, err); }
    if (rc == 0) {
      if (cudaMemcpyAsync(kin.p, keys, sizeof(u64) * count, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(vin.p, values, sizeof(u32) * count, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        err.set("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1;
      }
    }
    if (rc == 0) {
      PassPlan plan; plan.npass = 0;
      plan_add_bits(plan, (int) begin_bit, (int) end_bit);
      PairSrc ps{kin.as<u64>(), vin.as<u32>()};    // the first pass reads the staged input
      u64 *kk[2] = {k[0].as<u64>(), k[1].as<u64>()};
      u32 *vv[2] = {v[0].as<u32>(), v[1].as<u32>()};
      int res = 0; u64 nout = 0;
      rc = radix_sort(rw, st, ps, count, kk, vv, plan, &res, &nout, err);
      if (rc == 0) {
        if (cudaMemcpyAsync(keys, kk[res], sizeof(u64) * count, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(values, vv[res], sizeof(u32) * count, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { err.set("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
      }
    }
    kin.release(); vin.release();
    for (int i = 0; i < 2; i++) { k[i].release(); v[i].release(); }
    radix_work_free(rw);
    cudaStreamDestroy(st);
    return rc;
  };
  int rc = body();
  if (rc != 0 && errbuf && errlen) snprintf(errbuf, errlen, "%s", err.msg);
  return rc;
}

// records of `width` uint64 (1: plain keys; 2: pairs), sorted by component 0 (nkeys = 1) or by
// (component 0, component 1) (nkeys = 2): LSD over the components with the onesweep engine on
// (key, record index) pairs, then one gather of the records.  Stable.
static int sort_u64_records(int device, uint64_t *rec, uint64_t count, unsigned width, unsigned nkeys, ErrBuf &err)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    err.set("libgtb200: no CUDA device available; there is no CPU fallback"); return -1;
  }
  if (count >= 0xffffffffull) { err.set("more than 2^32-2 records are not supported"); return -1; }
  GTB_CUDA(cudaSetDevice(device));
  if (count <= 1) return 0;
  cudaStream_t st;
  GTB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  RadixWork rw;
  DevBuf k[2], v[2], kin, vin, drec, dout;
  int rc = radix_work_init(rw, err);
  for (int i = 0; i < 2 && rc == 0; i++) { rc |= k[i].ensure(sizeof(u64) * count, err); rc |= v[i].ensure(sizeof(u32) * count, err); }
  if (rc == 0) { rc |= kin.ensure(sizeof(u64) * count, err); rc |= vin.ensure(sizeof(u32) * count, err); }
  if (rc == 0) rc |= drec.ensure(sizeof(u64) * width * count, err);
  if (rc == 0) rc |= dout.ensure(sizeof(u64) * width * count, err);
  if (rc == 0 && cudaMemcpyAsync(drec.p, rec, sizeof(u64) * width * count, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    err.set("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1;
  }
  u64 *kk[2] = {k[0].as<u64>(), k[1].as<u64>()};
  u32 *vv[2] = {v[0].as<u32>(), v[1].as<u32>()};
  const u32 *perm = nullptr;
  for (int comp = (int) nkeys - 1; comp >= 0 && rc == 0; comp--) {     // least significant component first
    k_records_to_pairs<<<grid_for(count, 256), 256, 0, st>>>(drec.as<u64>(), width, (unsigned) comp, perm, count,
                                                           kin.as<u64>(), vin.as<u32>());
    if (cudaGetLastError() != cudaSuccess) { err.set("kernel launch failed"); rc = -1; break; }
    PassPlan plan; plan.npass = 0;
    plan_add_bits(plan, 0, 64);
    PairSrc ps{kin.as<u64>(), vin.as<u32>()};
    int res = 0; u64 nout = 0;
    rc = radix_sort(rw, st, ps, count, kk, vv, plan, &res, &nout, err);
    perm = vv[res];
    if (rc == 0 && comp > 0) {
      // the next sort overwrites both value buffers: keep this order aside
      if (cudaMemcpyAsync(dout.p, perm, sizeof(u32) * count, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { err.set("copy failed"); rc = -1; }
      perm = dout.as<u32>();
    }
  }
  if (rc == 0) {
    // (the order of the last sort lies in a value buffer; dout is free again)
    k_gather_records<<<grid_for(count, 256), 256, 0, st>>>(drec.as<u64>(), width, perm, count, dout.as<u64>());
    if (cudaGetLastError() != cudaSuccess) { err.set("kernel launch failed"); rc = -1; }
  }
  if (rc == 0 && (cudaMemcpyAsync(rec, dout.p, sizeof(u64) * width * count, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                  cudaStreamSynchronize(st) != cudaSuccess)) {
    err.set("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1;
  }
  kin.release(); vin.release(); drec.release(); dout.release();
  for (int i = 0; i < 2; i++) { k[i].release(); v[i].release(); }
  radix_work_free(rw);
  cudaStreamDestroy(st);
  return rc;
}

static int sort_u64_records_entry(int device, uint64_t *rec, uint64_t count, unsigned width, unsigned nkeys,
                                  char *errbuf, size_t errlen)
{
  ErrBuf err;
  const int rc = sort_u64_records(device, rec, count, width, nkeys, err);
  if (rc != 0 && errbuf && errlen) snprintf(errbuf, errlen, "%s", err.msg);
  return rc;
}

int gtb_radixsort_u64(int device, uint64_t *keys, uint64_t count, char *errbuf, size_t errlen)
{ return sort_u64_records_entry(device, keys, count, 1, 1, errbuf, errlen); }
int gtb_radixsort_u64pair(int device, uint64_t *pairs, uint64_t count, char *errbuf, size_t errlen)
{ return sort_u64_records_entry(device, pairs, count, 2, 1, errbuf, errlen); }
int gtb_radixsort_u64keypair(int device, uint64_t *pairs, uint64_t count, char *errbuf, size_t errlen)
{ return sort_u64_records_entry(device, pairs, count, 2, 2, errbuf, errlen); }

} // extern "C"
