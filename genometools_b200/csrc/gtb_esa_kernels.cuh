// gtb_esa_kernels.cuh -- device code of the enhanced-suffix-array pipeline.
//
// Key layout ("filled key"), top-aligned in 64 bits:
//   bits [63 : 64-m*b]      the first m symbols of the suffix, b bits each, MSB first;
//                           symbols at and after the first special are replaced by the
//                           all-ones filler (for DNA that is T, the filler GenomeTools
//                           uses for its special k-mers, sfx-mapped4.gen:180-228)
//   bits [sh+tb-1 : sh]     m - u, where u = number of regular symbols before the first
//                           special (capped at m);  0 for a "full" suffix.  sh = 64-m*b-tb
//   bits [sh-1 : 0]         zero
//   DNA: b = 2, m = 17 / 21 / 25 (tb = 5) or 29 (tb = 6), chosen from the text length so
//   that chance ties stay rare -- 39 / 47 / 55 / 64 significant bits = 5 / 6 / 7 / 8 radix
//   passes;  protein/bytes: m = 12, b = 5, tb = 4
// Sorting these keys with a stable sort from text order reproduces rule R
// (/root/reference/src/core/encseq.c:6449-6528,7371-7460)
// for every pair except full suffixes with identical keys, which are refined by
// text-driven rounds (the next symbols) or prefix doubling (ranks).
#pragma once
#include "gtb_common.cuh"

namespace gtb {

struct KeyFmt {
  int m;        // symbols per key
  int b;        // bits per symbol
  int tb;       // tail bits
  int sh;       // position of the tail field
  __host__ __device__ u64 tailmask() const { return ((1ull << tb) - 1ull) << sh; }
  __host__ __device__ u64 symmask() const { return ~0ull << (64 - m * b); }
  __host__ __device__ unsigned tail(u64 key) const { return (unsigned) (key >> sh) & ((1u << tb) - 1u); }
  __host__ __device__ int lowbit() const { return sh; }
};
__host__ __device__ inline KeyFmt make_fmt(int m, int b, int tb) { return KeyFmt{m, b, tb, 64 - m * b - tb}; }
__host__ __device__ inline KeyFmt byte_fmt() { return make_fmt(12, 5, 4); }
// byte path (5 bits per symbol): the smallest m in {8, 10, 12} (44 / 54 / 64 significant
// bits = 6 / 7 / 8 radix passes) with an expected share of chance ties n / K^m below 0.1 %
// (the margin covers skewed residue frequencies)
static inline KeyFmt byte_fmt_for(u64 n, unsigned K, unsigned pl)
{
  const int cand[3] = {8, 10, 12};
  for (int i = 0; i < 3; i++) {
    const int m = cand[i];
    if ((unsigned) m < pl) continue;
    double km = 1.0;
    for (int j = 0; j < m; j++) km *= (double) (K < 2 ? 2 : K);
    if ((double) n / km <= 0.001 || m == 12) return make_fmt(m, 5, 4);
  }
  return make_fmt(12, 5, 4);
}
// DNA key length for a text of n symbols: the smallest m in {17, 21, 25, 29} (not below
// the prefix length) for which the expected share of chance ties n / 4^m is below 1 %
static inline KeyFmt dna_fmt_for(u64 n, unsigned pl)
{
  const int cand[4] = {17, 21, 25, 29};
  for (int i = 0; i < 4; i++) {
    const int m = cand[i];
    if ((unsigned) m < pl) continue;
    const double ties = (double) n / (double) (1ull << (2 * m));
    if (ties <= 0.01 || m == 29) return make_fmt(m, 2, m == 29 ? 6 : 5);
  }
  return make_fmt(29, 2, 6);
}
// key of the text-driven refinement rounds: 13 symbols + 4 tail bits (DNA), 5 symbols +
// 3 tail bits (bytes), top-aligned in 32 bits
__host__ __device__ inline KeyFmt dna_tfmt()  { return make_fmt(13, 2, 4); }
__host__ __device__ inline KeyFmt byte_tfmt() { return make_fmt(5, 5, 3); }

// window of 32 mask bits starting at bit `pos`
__device__ __forceinline__ u32 mask_window(const u32 *__restrict__ spmask, u64 pos)
{
  const u64 w = pos >> 5;
  return __funnelshift_r(spmask[w], spmask[w + 1], (u32) (pos & 31u));
}

// 32 bases (64 bits, MSB first) starting at base `pos`
__device__ __forceinline__ u64 dna_window(const u64 *__restrict__ words, u64 pos)
{
  const u64 w = pos >> 5;
  const unsigned sh = (unsigned) (pos & 31u) * 2u;
  const u64 w0 = words[w], w1 = words[w + 1];
  return sh ? (w0 << sh) | (w1 >> (64u - sh)) : w0;
}

template <bool DNA>
struct TextSrc {
  static constexpr bool ALWAYS_VALID = false;
  static constexpr bool BLOCKED_GEN = true;  // keys of 16 consecutive positions at once (gen_block)
  const u64 *words;      // DNA: packed 2-bit words (padded by >= 2 words)
  const u8  *bytes;      // bytes path: symbols (padded by >= 16 bytes of 255)
  const u32 *spmask;     // bit i set <=> position i is special; bits >= n all set
  u64 klo, khi;          // inclusive key range of this shard
  u64 pos0;              // item i of the source is text position pos0 + i
  KeyFmt f;
  bool skip_near;        // positions that meet a special within their m symbols are no items (they are
                         // sorted from a list of their own, k_near_bits / TailSrc)
  const u64 *ranges;     // the special runs [start, end) ascending (device), for the 4-mer histogram; may be
  u64 nranges;           // null when there are none
  bool hist4;            // the digit histograms may be derived from one table of 4-mers (k_hist_4mer)

  // filled key of position pos in format g (any m <= 29 for DNA, m*b + tb <= 64)
  __device__ __forceinline__ bool make_key_fmt(u64 pos, u64 &key, const KeyFmt &g) const
  {
    const u32 win = mask_window(spmask, pos);
    if (win & 1u) return false;                         // special position
    const unsigned z = win ? (unsigned) __ffs(win) - 1u : 32u;
    const unsigned u = z < (unsigned) g.m ? z : (unsigned) g.m;
    u64 sym;
    if (DNA) {
      sym = dna_window(words, pos) >> (64 - 2 * g.m);   // m symbols in the low 2m bits
      if (u < (unsigned) g.m) sym |= (1ull << (2u * (g.m - u))) - 1ull;
    } else {
      sym = 0;
      for (int k = 0; k < g.m; k++) {
        const u64 c = (unsigned) k < u ? (u64) bytes[pos + k] : 31ull;
        sym = (sym << g.b) | c;
      }
    }
    key = (sym << (64 - g.m * g.b)) | ((u64) (g.m - u) << g.sh);
    return true;
  }
  __device__ __forceinline__ bool make_key(u64 pos, u64 &key) const
  {
    if (!make_key_fmt(pos, key, f)) return false;
    if (skip_near && (key & f.tailmask()) != 0) return false;
    return key >= klo && key <= khi;
  }
  __device__ __forceinline__ bool load_key(u64 idx, u64 &k) const
  { return make_key(pos0 + idx, k); }
  __device__ __forceinline__ u32 load_val(u64 idx) const { return (u32) (pos0 + idx); }
  // keys of the `cnt` <= 16 consecutive items idx .. idx + cnt - 1 into out[0 .. 15] (the rest
  // and every item that is special or outside [klo, khi] becomes ~0, which no filled key equals).
  // The 16 + m - 1 symbols a thread needs are loaded once and laid out as one big-endian bit
  // string in 32-bit registers (DNA: three words of the packed text; bytes: 27 symbols of 5
  // bits from two 128-bit loads); key i is then two funnel shifts by the constant i*b and two
  // ANDs, its distance to the next special one funnel shift of the mask -- about 8 instructions
  // per key unless a special lies within its m symbols.
  __device__ __forceinline__ bool can_gen_fast(u64 idx, u32 cnt) const
  { return cnt >= 16u && (DNA || ((pos0 + idx) & 15u) == 0); }
  __device__ __forceinline__ void gen_block(u64 idx, u32 cnt, u64 *out) const
  {
    if (!can_gen_fast(idx, cnt)) {                      // the ragged end of the source; unaligned slices
#pragma unroll 1
      for (u32 i = 0; i < 16u; i++) {
        u64 key = ~0ull;
        if (i < cnt && !make_key(pos0 + idx + i, key)) key = ~0ull;
        out[i] = key;
      }
      return;
    }
    gen_block_fast(idx, out);
  }
  // (only constant indices into out: it may live in registers)
  __device__ __forceinline__ void gen_block_fast(u64 idx, u64 *out) const
  {
    constexpr unsigned B = DNA ? 2u : 5u;
    const u64 p0 = pos0 + idx;
    const unsigned m = (unsigned) f.m;
    u32 S[DNA ? 3 : 5];
    if (DNA) {
      const u64 h64 = dna_window(words, p0);            // bases p0 .. p0+31
      S[0] = (u32) (h64 >> 32); S[1] = (u32) h64;
      S[2] = (u32) (dna_window(words, p0 + 32) >> 32);  // bases p0+32 .. p0+47
    } else {
      const uint4 a = *reinterpret_cast<const uint4 *>(bytes + p0);
      const uint4 c = *reinterpret_cast<const uint4 *>(bytes + p0 + 16);
      const u32 wv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int w = 0; w < 5; w++) S[w] = 0;
#pragma unroll
      for (int j = 0; j < 27; j++) {                    // symbol j at bits [160-5(j+1), 160-5j) of S
        const u32 sy = (wv[j >> 2] >> (8 * (j & 3))) & 31u;
        const int pos = 160 - 5 * (j + 1), word = 4 - pos / 32, off = pos % 32;
        S[word] |= sy << off;
        if (off > 27) S[word - 1] |= sy >> (32 - off);
      }
    }
    const u32 mlo = mask_window(spmask, p0), mhi = mask_window(spmask, p0 + 32);
    const u64 sm = f.symmask();
    const u32 smhi = (u32) (sm >> 32), smlo = (u32) sm;
    const u32 mmask = (1u << m) - 1u;                   // m <= 29
    const bool filter = klo != 0ull || khi != ~0ull;
#pragma unroll
    for (u32 i = 0; i < 16u; i++) {
      const u32 w = (B * i) >> 5, sh = (B * i) & 31u;   // (compile-time after unrolling)
      const u32 khi32 = __funnelshift_l(S[w + 1], S[w], sh) & smhi;
      const u32 klo32 = __funnelshift_l(S[w + 2], S[w + 1], sh) & smlo;
      u64 key = ((u64) khi32 << 32) | klo32;
      const u32 win = __funnelshift_r(mlo, mhi, i) & mmask;     // specials among positions i .. i+m-1
      if (win != 0u) {
        if ((win & 1u) || skip_near) key = ~0ull;       // a special position is no item
        else {
          const unsigned u = (unsigned) __ffs(win) - 1u;        // regular symbols before the special
          key |= (1ull << (64u - B * u)) - (1ull << (64u - B * m));   // filler over symbols u .. m-1
          key |= (u64) (m - u) << f.sh;
        }
      }
      if (filter && (key < klo || key > khi)) key = ~0ull;
      out[i] = key;
    }
  }
};

// ---- digit histograms of the whole DNA text with rolling keys -------------------------------
// The generic histogram kernel rebuilds every key from the packed text (~135 instructions
// per position).  Here a thread walks 32 consecutive positions (one 2-bit word): the key
// of the next position is the 128-bit text window shifted by one base, the distance to the
// next special the 64-bit mask window shifted by one bit.
__global__ void __launch_bounds__(256)
k_hist_text_dna(TextSrc<true> src, u64 n, PassPlan plan, unsigned long long *__restrict__ ghist)
{
  __shared__ u32 s_h[RS_MAXPASS * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAXPASS * RS_BINS; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const KeyFmt f = src.f;
  const unsigned m = (unsigned) f.m;
  const u64 nchunks = (n + 31) >> 5;
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < nchunks; c += (u64) gridDim.x * blockDim.x) {
    u64 hi = src.words[c], lo = src.words[c + 1];                       // bases c*32 .. c*32+63
    u64 mw = (u64) src.spmask[c] | ((u64) src.spmask[c + 1] << 32);     // their special bits
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
      if (!(mw & 1ull)) {                                               // (positions >= n are marked special)
        const unsigned z = mw ? (unsigned) __ffsll((long long) mw) - 1u : 64u;
        const unsigned u = z < m ? z : m;
        u64 sym = hi >> (64 - 2 * m);
        if (u < m) sym |= (1ull << (2u * (m - u))) - 1ull;
        const u64 key = (sym << (64 - 2 * m)) | ((u64) (m - u) << f.sh);
        if (key >= src.klo && key <= src.khi && !(src.skip_near && u < m)) {
#pragma unroll
          for (int p = 0; p < RS_MAXPASS; p++)
            if (p < plan.npass)
              atomicAdd(&s_h[p * RS_BINS + ((unsigned) (key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
        }
      }
      hi = (hi << 2) | (lo >> 62);
      lo <<= 2;
      mw >>= 1;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += blockDim.x)
    if (s_h[i]) atomicAdd(&ghist[i], (unsigned long long) s_h[i]);
}

// the same for any text source, keys from gen_block_fast (used for the byte path: ~135
// instructions per position in the generic kernel, mostly byte loads)
template <bool DNA>
__global__ void __launch_bounds__(256)
k_hist_text_blocked(TextSrc<DNA> src, u64 n, PassPlan plan, unsigned long long *__restrict__ ghist)
{
  __shared__ u32 s_h[RS_MAXPASS * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAXPASS * RS_BINS; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const u64 nchunks = (n + 15) >> 4;
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < nchunks; c += (u64) gridDim.x * blockDim.x) {
    const u64 idx = c << 4;
    const u32 cnt = n - idx < 16 ? (u32) (n - idx) : 16u;
    if (src.can_gen_fast(idx, cnt)) {
      u64 keys[16];
      src.gen_block_fast(idx, keys);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (keys[i] == ~0ull) continue;
#pragma unroll
        for (int p = 0; p < RS_MAXPASS; p++)
          if (p < plan.npass)
            atomicAdd(&s_h[p * RS_BINS + ((unsigned) (keys[i] >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
      }
    } else {
      for (u32 i = 0; i < cnt; i++) {
        u64 key;
        if (!src.load_key(idx + i, key)) continue;
        for (int p = 0; p < plan.npass; p++)
          atomicAdd(&s_h[p * RS_BINS + ((unsigned) (key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += blockDim.x)
    if (s_h[i]) atomicAdd(&ghist[i], (unsigned long long) s_h[i]);
}

template <>
struct RsHistLauncher<TextSrc<false>> {
  static void launch(const TextSrc<false> &src, u64 nsrc, const PassPlan &plan, unsigned long long *ghist,
                     cudaStream_t st)
  {
    const u64 chunks = (nsrc + 15) >> 4;
    u64 g = div_up(chunks, 256);
    if (g > 148ull * 8) g = 148ull * 8;
    if (g < 1) g = 1;
    k_hist_text_blocked<false><<<(unsigned) g, 256, 0, st>>>(src, nsrc, plan, ghist);
  }
};

// ---- all digit histograms of the full keys from ONE table of 4-mers ----------------------------
// With keys of m = 4k symbols and the tail keys sorted apart (skip_near), digit j of the key of a full
// position i (no special in [i, i + m)) is the plain 4-mer at i + 4j.  So every digit histogram is the
// 4-mer table T of the text minus the windows that no full position reaches: for a run [a, b) of regular
// symbols the full positions are [a, b - m], digit j reads the windows [a + 4j, b - m + 4j] -- the first
// 4j and the last m - 4 - 4j windows of the run drop out (all of them if the run is shorter than m).
// One shared-memory atomic per position instead of one per position and digit (5 on c4: 11.1 -> about
// 3 ms), and no key is built.  ghist row p holds the digit at bit plan.shift[p].
__global__ void __launch_bounds__(256)
k_hist_4mer(const u64 *__restrict__ words, const u32 *__restrict__ spmask, u64 n, int npass,
            unsigned long long *__restrict__ ghist)
{
  __shared__ u32 s_h[RS_BINS];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const u64 nchunks = (n + 31) >> 5;
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < nchunks; c += (u64) gridDim.x * blockDim.x) {
    u64 hi = words[c], lo = words[c + 1];                     // bases 32c .. 32c + 63
    const u64 mw = (u64) spmask[c] | ((u64) spmask[c + 1] << 32);
    u64 bad = mw | (mw >> 1) | (mw >> 2) | (mw >> 3);         // a special among the 4 symbols of the window
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
      if (!(bad & 1ull)) atomicAdd(&s_h[(u32) (hi >> 56)], 1u);
      hi = (hi << 2) | (lo >> 62);
      lo <<= 2;
      bad >>= 1;
    }
  }
  __syncthreads();
  const u32 v = s_h[threadIdx.x];
  if (v) for (int p = 0; p < npass; p++) atomicAdd(&ghist[p * RS_BINS + threadIdx.x], (unsigned long long) v);
}

// the windows a digit does not read: one thread per (run of regular symbols, digit)
__global__ void k_hist_4mer_fix(const u64 *__restrict__ words, const u64 *__restrict__ ranges, u64 nranges, u64 n,
                                unsigned m, PassPlan plan, unsigned long long *__restrict__ ghist)
{
  const u64 total = (nranges + 1) * (u64) plan.npass;
  for (u64 t = blockIdx.x * (u64) blockDim.x + threadIdx.x; t < total; t += (u64) gridDim.x * blockDim.x) {
    const u64 k = t / (u64) plan.npass;
    const int p = (int) (t % (u64) plan.npass);
    const unsigned j = (unsigned) (64 - plan.shift[p]) / 8u - 1u;     // digit j from the top: symbols 4j .. 4j+3
    const u64 a = k == 0 ? 0 : ranges[2 * (k - 1) + 1], b = k == nranges ? n : ranges[2 * k];
    if (b < a + 4) continue;                                   // no window of 4 regular symbols
    const u64 last = b - 4;                                    // windows a .. last
    unsigned long long *row = ghist + p * RS_BINS;
    if (b - a < m) {
      for (u64 q = a; q <= last; q++) atomicAdd(&row[(u32) (dna_window(words, q) >> 56)], ~0ull);
    } else {
      for (u64 q = a; q < a + 4 * j; q++) atomicAdd(&row[(u32) (dna_window(words, q) >> 56)], ~0ull);
      for (u64 q = b - m + 4 * j + 1; q <= last; q++) atomicAdd(&row[(u32) (dna_window(words, q) >> 56)], ~0ull);
    }
  }
}

template <>
struct RsHistLauncher<TextSrc<true>> {
  static void launch(const TextSrc<true> &src, u64 nsrc, const PassPlan &plan, unsigned long long *ghist,
                     cudaStream_t st)
  {
    bool bytes_from_top = src.hist4 && src.skip_near && src.pos0 == 0 && src.klo == 0 && src.khi == ~0ull &&
                          src.f.m % 4 == 0 && (src.nranges == 0 || src.ranges != nullptr);
    for (int p = 0; p < plan.npass && bytes_from_top; p++)
      bytes_from_top = plan.bits[p] == 8 && plan.shift[p] % 8 == 0 && plan.shift[p] >= 64 - 2 * src.f.m;
    if (bytes_from_top) {                // every digit is a 4-mer of the key's symbols
      const u64 chunks = (nsrc + 31) >> 5;
      u64 g = div_up(chunks, 256);
      if (g > 148ull * 8) g = 148ull * 8;
      if (g < 1) g = 1;
      k_hist_4mer<<<(unsigned) g, 256, 0, st>>>(src.words, src.spmask, nsrc, plan.npass, ghist);
      const u64 fix = (src.nranges + 1) * (u64) plan.npass;
      u64 g2 = div_up(fix, 128);
      if (g2 > 148ull * 16) g2 = 148ull * 16;
      k_hist_4mer_fix<<<(unsigned) g2, 128, 0, st>>>(src.words, src.ranges, src.nranges, nsrc, (unsigned) src.f.m, plan, ghist);
      return;
    }
    if (src.pos0 == 0) {                 // the whole text: rolling keys
      const u64 chunks = (nsrc + 31) >> 5;
      u64 g = div_up(chunks, 256);
      if (g > 148ull * 8) g = 148ull * 8;
      if (g < 1) g = 1;
      k_hist_text_dna<<<(unsigned) g, 256, 0, st>>>(src, nsrc, plan, ghist);
    } else {
      u64 tiles = div_up(nsrc, RH_TILE);
      unsigned grid = (unsigned) (tiles < 148ull * 8 ? tiles : 148ull * 8);
      rs_hist_kernel<TextSrc<true>, false><<<grid, RH_NT, 0, st>>>(src, nsrc, plan, ghist);
    }
  }
};

// ---- ranks without an inverse suffix array -----------------------------------------------
// Prefix doubling needs rank(q) = suffix-array index of q (of its group head while q is
// tied) for the positions q = p + h of the tied suffixes p only -- about 1 % of a genome.
// Instead of an n-sized inverse suffix array (one random 4-byte write per suffix) the
// rank is derived on demand:
//   q special      -> (n - S) + number of specials before q   (popcount prefix of the mask)
//   q tied         -> trank[dense id of q]   (bitmap of the tied positions + popcount
//                     prefix; trank holds the current group head, updated every round)
//   q anything else-> binary search of (key(q), q) in the sorted first-level keys: its
//                     key is unique, or shared only by suffixes that meet a special at
//                     the same depth, which are stored in text order
template <bool DNA>
__device__ __forceinline__ u64 key_code(u64 key, unsigned pl, unsigned K, const KeyFmt &f);

template <bool DNA>
struct RankMap {
  TextSrc<DNA> src;
  const u64 *keys;           // first-level sorted keys of this range
  const u32 *sa;             // its suffix table
  u64 N;
  // per 32 text positions one 16-byte word {special bits, specials before, tied bits, tied
  // before}: a rank query is one random 16-byte load (+ one into trank) instead of four loads
  // from four arrays
  const uint4 *rw;
  u32 *trank;
  const u32 *leftborder;     // bucket starts (global indices) or null: narrows the search
  u64 own_last;              // last code whose successor entry is not in this table (a code range
                             // fills its own codes only), ~0 if the table is complete
  unsigned pl, K;
  u64 n, nonspecials, sa_offset;

  __device__ __forceinline__ void set(u64 p, u32 r) const
  {
    const uint4 w = __ldg(rw + (p >> 5));
    trank[w.w + (u32) __popc(w.z & ((1u << (p & 31u)) - 1u))] = r;
  }
  // the part of a query that is one or two loads: end of text, special, tied.  false: q is an untied
  // regular suffix -- its rank is its place among the sorted keys (search)
  __device__ __forceinline__ bool get_fast(u64 q, u32 &r) const
  {
    if (q >= n) { r = (u32) n; return true; }
    const uint4 w = __ldg(rw + (q >> 5));
    const u32 bit = 1u << (q & 31u), below = bit - 1u;
    if (w.x & bit) { r = (u32) (nonspecials + w.y + (u32) __popc(w.x & below)); return true; }
    if (w.z & bit) { r = trank[w.w + (u32) __popc(w.z & below)]; return true; }
    return false;
  }
  __device__ u32 get(u64 q) const
  {
    u32 r;
    if (get_fast(q, r)) return r;
    return search(q);
  }
  __device__ u32 search(u64 q) const
  {
    u64 kq;
    src.make_key_fmt(q, kq, src.f);
    u64 lo = 0, hi = N;
    if (leftborder) {          // the suffix lies in its bucket: a few dozen entries
      const u64 c = key_code<DNA>(kq, pl, K, src.f);
      lo = (u64) leftborder[c] - sa_offset;
      hi = c == own_last ? N : (u64) leftborder[c + 1] - sa_offset;
    }
    while (lo < hi) {
      const u64 mid = (lo + hi) >> 1;
      const u64 km = keys[mid];
      const bool less = km < kq || (km == kq && (u64) sa[mid] < q);
      if (less) lo = mid + 1; else hi = mid;
    }
    return (u32) (sa_offset + lo);
  }
};

// ---- the keys that carry a tail, sorted apart (radix_sort with a second source) ---------------
// bit i of word w: position 32w + i is regular and meets a special (or the end of the text) within
// its next m symbols -- its filled key has a non-zero tail field
__global__ void k_near_bits(const u32 *__restrict__ spmask, u64 nwords, unsigned m, u32 *__restrict__ near)
{
  for (u64 w = blockIdx.x * (u64) blockDim.x + threadIdx.x; w < nwords; w += (u64) gridDim.x * blockDim.x) {
    const u64 sp = (u64) spmask[w] | ((u64) spmask[w + 1] << 32);
    u64 sm = sp >> 1;                         // special among positions +1 .. +(m-1): OR of m-1 shifts
    unsigned have = 1;
    while (have < m - 1) {
      const unsigned step = have < m - 1 - have ? have : m - 1 - have;
      sm |= sm >> step;
      have += step;
    }
    near[w] = (u32) (~sp & sm);
  }
}

struct TailSrc {                 // (key, position) pairs in memory, only those inside the key range
  static constexpr bool ALWAYS_VALID = false;
  static constexpr bool BLOCKED_GEN = false;
  const u64 *keys;
  const u32 *vals;
  u64 klo, khi;
  __device__ __forceinline__ bool load_key(u64 idx, u64 &k) const
  { k = keys[idx]; return k >= klo && k <= khi; }
  __device__ __forceinline__ u32 load_val(u64 idx) const { return vals[idx]; }
};

// ---- special mask ------------------------------------------------------------------
// bits >= n are set (the end of the text acts as a special, sfx-enumcodes.c:141-154)
__global__ void k_mask_tail(u32 *spmask, u64 n, u64 nmaskwords)
{
  const u64 first = n >> 5;
  for (u64 w = first + blockIdx.x * (u64) blockDim.x + threadIdx.x; w < nmaskwords;
       w += (u64) gridDim.x * blockDim.x) {
    if (w == first) atomicOr(&spmask[w], ~0u << (n & 31u));
    else spmask[w] = ~0u;
  }
}

// one warp per special range (ranges from gt_specialrangeiterator, encseq.h:127)
__global__ void k_mask_ranges(u32 *spmask, const u64 *__restrict__ ranges, u64 nranges)
{
  const u64 warps = ((u64) gridDim.x * blockDim.x) >> 5;
  const u64 wid = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  for (u64 r = wid; r < nranges; r += warps) {
    const u64 s = ranges[2 * r], e = ranges[2 * r + 1];
    if (e <= s) continue;
    const u64 w0 = s >> 5, w1 = (e - 1) >> 5;
    for (u64 w = w0 + lane; w <= w1; w += 32) {
      u32 bits = ~0u;
      if (w == w0) bits &= ~0u << (s & 31u);
      if (w == w1) bits &= ~0u >> (31u - ((e - 1) & 31u));
      atomicOr(&spmask[w], bits);
    }
  }
}

// bytes path: one thread builds one mask word from 32 symbols
__global__ void k_mask_from_bytes(u32 *spmask, const u8 *__restrict__ bytes, u64 n)
{
  const u64 nw = (n + 31) >> 5;
  for (u64 w = blockIdx.x * (u64) blockDim.x + threadIdx.x; w < nw;
       w += (u64) gridDim.x * blockDim.x) {
    const uint4 *p = reinterpret_cast<const uint4 *>(bytes + w * 32);
    const uint4 a = p[0], c = p[1];
    const u32 v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    u32 bits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const u32 by = (v[i] >> (8 * j)) & 0xffu;
        if (by >= 254u) bits |= 1u << (4 * i + j);
      }
    }
    const u64 lastpos = w * 32 + 31;
    if (lastpos >= n) bits &= (n & 31u) ? ~(~0u << (n & 31u)) : 0u;   // positions >= n handled by k_mask_tail
    if (w * 32 >= n) bits = 0;
    if (bits) atomicOr(&spmask[w], bits);
  }
}

// ---- generic exclusive scan of a u32 array (three launches) ---------------------------
constexpr int SC_NT = 256, SC_IPT = 8, SC_TILE = SC_NT * SC_IPT;

__global__ void __launch_bounds__(SC_NT)
k_scan_tilesums(const u32 *__restrict__ in, u64 count, u32 *__restrict__ tilesum, int popc_mode)
{
  __shared__ u32 scratch[SC_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * SC_TILE + (u64) threadIdx.x * SC_IPT;
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++)
    if (base + k < count) s += popc_mode ? (u32) __popc(in[base + k]) : in[base + k];
  u32 total;
  block_exclusive_sum<SC_NT, u32>(s, scratch, &total);
  if (threadIdx.x == 0) tilesum[blockIdx.x] = total;
}

// single block: in-place exclusive scan of `cnt` values; total -> *total_out (u64)
__global__ void __launch_bounds__(1024)
k_scan_single(u32 *vals, u64 cnt, u64 *total_out)
{
  __shared__ u32 scratch[1024 / 32 + 1];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  u64 grand = 0;
  for (u64 base = 0; base < cnt; base += 1024) {
    const u64 i = base + threadIdx.x;
    const u32 v = i < cnt ? vals[i] : 0u;
    u32 total;
    const u32 ex = block_exclusive_sum<1024, u32>(v, scratch, &total);
    const u32 carry = carry_s;
    if (i < cnt) vals[i] = ex + carry;
    grand += total;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = grand;
}

__global__ void __launch_bounds__(SC_NT)
k_scan_apply(const u32 *in, u32 *out, u64 count,
             const u32 *__restrict__ tileoff)
{
  __shared__ u32 scratch[SC_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * SC_TILE + (u64) threadIdx.x * SC_IPT;
  u32 v[SC_IPT], s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { v[k] = base + k < count ? in[base + k] : 0u; s += v[k]; }
  u32 total;
  u32 ex = block_exclusive_sum<SC_NT, u32>(s, scratch, &total) + tileoff[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { if (base + k < count) out[base + k] = ex; ex += v[k]; }
}

// out[w] = number of set bits in in[0..w) (exclusive popcount prefix per word)
__global__ void __launch_bounds__(SC_NT)
k_scan_apply_popc(const u32 *in, u32 *out, u64 count, const u32 *__restrict__ tileoff)
{
  __shared__ u32 scratch[SC_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * SC_TILE + (u64) threadIdx.x * SC_IPT;
  u32 v[SC_IPT], s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { v[k] = base + k < count ? (u32) __popc(in[base + k]) : 0u; s += v[k]; }
  u32 total;
  u32 ex = block_exclusive_sum<SC_NT, u32>(s, scratch, &total) + tileoff[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) { if (base + k < count) out[base + k] = ex; ex += v[k]; }
}

// ---- special tail: positions of specials ascending (sfx-suffixgetset.c:586-690) -------
// tileoff: exclusive scan of the per-tile popcounts of the mask (over positions < n)
__global__ void __launch_bounds__(SC_NT)
k_emit_special_tail(const u32 *__restrict__ spmask, u64 nmaskwords_n, u64 n,
                    const u32 *__restrict__ tileoff, u32 *__restrict__ sa_tail,
                    u32 *__restrict__ isa, u64 rank_base, unsigned long long *longest)
{
  __shared__ u32 scratch[SC_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * SC_TILE + (u64) threadIdx.x * SC_IPT;
  u32 w[SC_IPT], s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    u32 x = base + k < nmaskwords_n ? spmask[base + k] : 0u;
    const u64 p0 = (base + k) * 32;
    if (p0 + 32 > n) x = p0 >= n ? 0u : (x & ~(~0u << (n - p0)));   // only positions < n
    w[k] = x; s += __popc(x);
  }
  u32 total;
  u32 ex = block_exclusive_sum<SC_NT, u32>(s, scratch, &total) + tileoff[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; k++) {
    u32 x = w[k];
    while (x) {
      const unsigned bit = __ffs(x) - 1u;
      x &= x - 1u;
      const u64 pos = (base + k) * 32 + bit;
      sa_tail[ex] = (u32) pos;
      if (isa) isa[pos] = (u32) (rank_base + ex);
      if (pos == 0) *longest = rank_base + ex;
      ex++;
    }
  }
}

// ---- bucket table (K1): code histogram incl. special k-mers ----------------------------
// leftborder counts every suffix with u >= 1 under its filled pl-code
// (sfx-suffixer.c:1070-1102); countspecialcodes / distpfxidx as bcktab.c:876-901.
// bucket code of a filled key (top pl symbols; the filler counts as the largest symbol)
template <bool DNA>
__device__ __forceinline__ u64 key_code(u64 key, unsigned pl, unsigned K, const KeyFmt &f)
{
  if (DNA) return key >> (64 - 2 * pl);
  u64 code = 0;
  for (unsigned k = 0; k < pl; k++) {
    u64 s = (key >> (64 - f.b * (k + 1))) & 31u;
    if (s >= K) s = K - 1;
    code = code * K + s;
  }
  return code;
}

// code of the first `len` <= pl symbols of a filled key (= key_code / K^(pl-len)); codes fit 32 bits
template <bool DNA>
__device__ __forceinline__ u32 key_code_prefix(u64 key, unsigned len, unsigned K, const KeyFmt &f)
{
  if (DNA) return len ? (u32) (key >> (64u - 2u * len)) : 0u;
  u32 code = 0;
  for (unsigned k = 0; k < len; k++) {
    u32 s = (u32) (key >> (64u - (unsigned) f.b * (k + 1u))) & 31u;
    if (s >= K) s = K - 1u;
    code = code * K + s;
  }
  return code;
}

// COUNT_ALL = false: only the two special-code tables (the bucket starts are then taken
// from the sorted keys by k_analyze_keys -- no atomic per suffix)
template <bool DNA, bool COUNT_ALL>
__global__ void __launch_bounds__(256)
k_count_codes(TextSrc<DNA> src, u64 first, u64 n, unsigned pl, unsigned K,
              u32 *__restrict__ cnt, u32 *__restrict__ csc, u32 *__restrict__ dist,
              const u64 *__restrict__ distoff /* [pl] offsets of level u (u>=1) */)
{
  // positions [first, n)
  const KeyFmt f = src.f;
  for (u64 pos = first + blockIdx.x * (u64) blockDim.x + threadIdx.x; ; pos += (u64) gridDim.x * blockDim.x) {
    const bool inb = pos < n;
    u64 key = 0;
    bool ok = false;
    if (inb) {
      if (COUNT_ALL) ok = src.make_key(pos, key);
      else {
        // only positions with a special among the next pl symbols matter
        const u32 win = mask_window(src.spmask, pos);
        if (!(win & 1u) && (win & ((1u << pl) - 1u))) ok = src.make_key(pos, key);
      }
    }
    u64 code = 0;
    unsigned u = 0;
    if (ok) {
      u = (unsigned) f.m - f.tail(key);
      code = key_code<DNA>(key, pl, K, f);
    }
    if (COUNT_ALL) {
      // warp-aggregated increment (poly-A buckets would otherwise serialise in L2)
      const unsigned active = __ballot_sync(FULL_MASK, ok);
      if (ok) {
        const unsigned peers = __match_any_sync(active, code);
        if ((int) lane_id() == __ffs(peers) - 1) atomicAdd(&cnt[code], (u32) __popc(peers));
      }
    }
    if (ok && u < pl) {
      atomicAdd(&csc[code / K], 1u);
      if (u + 1 < pl) {
        u64 lead = 0;
        if (DNA) lead = code >> (2 * (pl - u));
        else { u64 d = 1; for (unsigned k = u; k < pl; k++) d *= K; lead = code / d; }
        atomicAdd(&dist[distoff[u] + lead], 1u);
      }
    }
    if (!__any_sync(FULL_MASK, inb)) break;
  }
}

// ---- analysis of a sorted (key, pos) array ------------------------------------------------
struct DevStats {
  unsigned long long lcpsum;        // over suffixes with u >= pl
  unsigned long long numlarge;
  unsigned long long longest;
  unsigned int maxlcp;
  unsigned int pad;
};

constexpr int AN_NT = 256, AN_IPT = 8, AN_TILE = AN_NT * AN_IPT;

template <bool DNA>
__device__ __forceinline__ u32 key_lcp(u64 ka, u64 kb, const KeyFmt &f)
{
  constexpr u32 B = DNA ? 2u : 5u;      // bits per symbol (a compile-time divisor)
  const u64 x = (ka ^ kb) & f.symmask();
  u32 l = x ? (u32) __clzll((long long) x) / B : (u32) f.m;
  const u32 ua = (u32) f.m - f.tail(ka), ub = (u32) f.m - f.tail(kb);
  l = l < ua ? l : ua;
  return l < ub ? l : ub;
}

// is j the head of a group (== not tied with its predecessor by a full key)?
__device__ __forceinline__ bool key_head(u64 kprev, u64 kcur, u64 tmask)
{
  return kprev != kcur || (kcur & tmask) != 0;
}

struct AnalyzeArgs {
  u64 N; KeyFmt f; u64 tmask; unsigned pl, K; int seam_prev_valid; u64 seam_prev_key;
  u32 *leftborder; u64 ncodes; u32 *csc; u32 *dist; const u64 *distoff;
  // a code range fills only its own codes [lbfirst, lblast] with global indices (sa_offset + j)
  u64 lbfirst, lblast, sa_offset;
};

// the AN_IPT elements of one thread; INNER: the tile lies strictly inside the array
template <bool DNA, bool FILL_LB, bool INNER>
__device__ __forceinline__ void
analyze_thread(const u64 (&k)[AN_IPT + 2], u64 base, const AnalyzeArgs &a, u32 &unres, u32 &lasthead,
               u64 &lcpword, u32 &mx, unsigned long long &sum, u32 &headbits, u32 &unresbits)
{
  const KeyFmt f = a.f;
  const u64 N = a.N;
  u32 run = 0;                         // length of the current run of equal special keys
#pragma unroll
  for (int i = 0; i < AN_IPT; i++) {
    const u64 j = base + i;
    if (!INNER && j >= N) break;
    // the common case: both neighbours are full keys (no special within their m symbols)
    const bool fullpair = ((k[i] | k[i + 1]) & a.tmask) == 0 && (INNER || j > 0);
    const bool head = fullpair ? k[i] != k[i + 1] : ((!INNER && j == 0) || key_head(k[i], k[i + 1], a.tmask));
    const bool nexthead = (!INNER && j + 1 >= N) || key_head(k[i + 1], k[i + 2], a.tmask);
    if (head) {
      lasthead = (u32) j + 1u;
      u32 l = 0, u = (u32) f.m;
      if (fullpair) {
        constexpr u32 B = DNA ? 2u : 5u;
        l = (u32) __clzll((long long) (k[i] ^ k[i + 1])) / B;   // (they differ in a symbol)
      } else {
        if (INNER || j > 0) l = key_lcp<DNA>(k[i], k[i + 1], f);
        else if (a.seam_prev_valid) l = key_lcp<DNA>(a.seam_prev_key, k[i + 1], f);
        u = (u32) f.m - f.tail(k[i + 1]);
      }
      lcpword |= (u64) l << (8 * i);
      if (u >= a.pl) sum += l;
      mx = l > mx ? l : mx;
      if (FILL_LB && ((!INNER && j == 0) || l < a.pl)) {
        // a new bucket starts here: every code in (code of j-1, code of j] starts at j
        const u64 c1 = key_code<DNA>(k[i + 1], a.pl, a.K, f);
        u64 c0 = (!INNER && j == 0) ? a.lbfirst : key_code<DNA>(k[i], a.pl, a.K, f) + 1;
        for (; c0 <= c1; c0++) a.leftborder[c0] = (u32) (a.sa_offset + j);
      }
      if (FILL_LB && u < a.pl) {
        run++;
        if (i == AN_IPT - 1 || (!INNER && j + 1 >= N) || k[i + 2] != k[i + 1]) {   // the run ends here
          const u64 code = key_code<DNA>(k[i + 1], a.pl, a.K, f);
          atomicAdd(&a.csc[code / a.K], run);
          if (u + 1 < a.pl) {
            u64 lead = 0;
            if (DNA) lead = code >> (2 * (a.pl - u));
            else { u64 d = 1; for (unsigned q = u; q < a.pl; q++) d *= a.K; lead = code / d; }
            atomicAdd(&a.dist[a.distoff[u] + lead], run);
          }
          run = 0;
        }
      }
    }                                   // else: pending, lcp byte 0 = refinement level 0
    if (FILL_LB && !INNER && j + 1 == N) {
      for (u64 c0 = key_code<DNA>(k[i + 1], a.pl, a.K, f) + 1; c0 <= a.lblast; c0++)
        a.leftborder[c0] = (u32) (a.sa_offset + N);
    }
    if (head) headbits |= 1u << i;
    if (!head || !nexthead) { unres++; unresbits |= 1u << i; }
  }
}

// The AN_IPT elements of a thread whose ten keys lie inside the array, some with tail fields
// (reads, proteins: every warp meets some).  The per-element part is branch-free -- about 30
// instructions; what is rare per element (a bucket start, the end of a run of equal keys that
// met a special within pl symbols) is collected in bit masks and handled after the loop, so a
// warp runs those bodies once per thread run instead of once per element.
template <bool DNA, bool FILL_LB>
__device__ __forceinline__ void
analyze_thread_inner(const u64 (&k)[AN_IPT + 2], u64 base, const AnalyzeArgs &a, u32 &unres, u32 &lasthead,
                     u64 &lcpword, u32 &mx, unsigned long long &sum, u32 &headbits, u32 &unresbits)
{
  constexpr u32 B = DNA ? 2u : 5u;
  const KeyFmt f = a.f;
  const u32 m = (u32) f.m, pl = a.pl;
  const u64 symmask = f.symmask();
  const unsigned codesh = 64u - B * pl;
  u32 hb = 0, xnz = 0, spbits = 0, lbbits = 0, lsum = 0;
  u32 tprev = f.tail(k[0]);
#pragma unroll
  for (int i = 0; i < AN_IPT; i++) {
    const u64 x = k[i] ^ k[i + 1];
    const u32 tcur = f.tail(k[i + 1]);
    const u64 xs = x & symmask;
    u32 l = xs ? (u32) __clzll((long long) xs) / B : m;
    const u32 ua = m - tprev, ub = m - tcur;
    l = l < ua ? l : ua;
    l = l < ub ? l : ub;
    const bool head = x != 0 || tcur != 0;
    if (x != 0) xnz |= 1u << i;
    if (head) {
      hb |= 1u << i;
      lcpword |= (u64) l << (8 * i);
      if (ub >= pl) lsum += l;
      mx = l > mx ? l : mx;
      if (FILL_LB && pl > 0 && (x >> codesh) != 0) lbbits |= 1u << i;   // the top pl symbols differ
      if (FILL_LB && ub < pl) spbits |= 1u << i;
    }
    tprev = tcur;
  }
  sum += lsum;
  const u64 xlast = k[AN_IPT] ^ k[AN_IPT + 1];
  if (xlast != 0) xnz |= 1u << AN_IPT;
  if (xlast != 0 || (k[AN_IPT + 1] & a.tmask) != 0) hb |= 1u << AN_IPT;
  headbits = hb & 0xffu;
  unresbits = (~hb | ~(hb >> 1)) & 0xffu;
  unres = (u32) __popc(unresbits);
  lasthead = headbits ? (u32) base + (31u - (u32) __clz(headbits)) + 1u : 0u;
  if (FILL_LB && lbbits) {
    // a new bucket may start at these elements: every code in (code of j-1, code of j] starts at j
    // (the filler counts as the largest symbol, so different top symbols can still be one code)
#pragma unroll
    for (int i = 0; i < AN_IPT; i++)
      if ((lbbits >> i) & 1u) {
        const u32 c1 = key_code_prefix<DNA>(k[i + 1], pl, a.K, f);
        for (u32 c0 = key_code_prefix<DNA>(k[i], pl, a.K, f) + 1u; c0 <= c1 && c0 != 0u; c0++)
          a.leftborder[c0] = (u32) (a.sa_offset + base + i);
      }
  }
  if (FILL_LB && spbits) {
    // keys that met a special within their first pl symbols: equal ones are adjacent; one atomic per
    // run of equal keys inside this thread.  A run ends where the next key differs or the thread ends
    const u32 runend = spbits & ((xnz >> 1) | (1u << (AN_IPT - 1)));
    const u32 runstart = spbits & (xnz | 1u);
#pragma unroll
    for (int i = 0; i < AN_IPT; i++)
      if ((runend >> i) & 1u) {
        const u32 first = 31u - (u32) __clz(runstart & ((2u << i) - 1u));
        const u32 run = (u32) i - first + 1u;
        const u32 u = m - f.tail(k[i + 1]);
        atomicAdd(&a.csc[key_code_prefix<DNA>(k[i + 1], pl - 1u, a.K, f)], run);
        if (u + 1u < pl) atomicAdd(&a.dist[a.distoff[u] + key_code_prefix<DNA>(k[i + 1], u, a.K, f)], run);
      }
  }
}

// pass 1: lcp of resolved neighbours, per-tile count of unresolved + last head; with
// FILL_LB also the whole bucket table from the sorted keys: leftborder[c] = index of the
// first key with code >= c; countspecialcodes / distpfxidx from the keys that met a special
// within their first pl symbols (equal ones are adjacent: one atomic per run and thread).
//
// A warp owns 32 * AN_IPT consecutive keys, a thread AN_IPT of them: four 128-bit loads, the
// two neighbours by shuffle, no shared memory and no block barrier.  A warp whose keys are
// all full (no special within their m symbols -- all but ~1e-4 of a genome) takes a lean
// path of about 16 instructions per key; anything else (tail fields, the two ends of the
// array) goes through analyze_thread.  tile_unres / tile_lasthead must be zero on entry (the
// warps of a tile add / max into them).
template <bool DNA, bool FILL_LB>
__global__ void __launch_bounds__(AN_NT)
k_analyze_keys(const u64 *__restrict__ keys, u64 N, KeyFmt f, unsigned pl, unsigned K,
               u8 *__restrict__ lcp8, u32 *__restrict__ tile_unres,
               u32 *__restrict__ tile_lasthead, DevStats *stats, int seam_prev_valid,
               u64 seam_prev_key, u32 *__restrict__ leftborder, u64 ncodes,
               u32 *__restrict__ csc, u32 *__restrict__ dist, const u64 *__restrict__ distoff,
               u8 *__restrict__ hbits, u8 *__restrict__ ubits /* one byte per thread: bit i = element i */,
               u64 lbfirst, u64 lblast, u64 sa_offset)
{
  static_assert(AN_IPT == 8, "one 8-byte lcp store and one flag byte per thread");
  constexpr u32 WCHUNK = 32 * AN_IPT;                      // keys per warp and iteration
  constexpr u32 B = DNA ? 2u : 5u;
  __shared__ u32 s_max[AN_NT / 32];
  __shared__ unsigned long long s_sum[AN_NT / 32];
  const u64 tmask = f.tailmask();
  const AnalyzeArgs a{N, f, tmask, pl, K, seam_prev_valid, seam_prev_key, leftborder, ncodes, csc, dist, distoff,
                      lbfirst, lblast, sa_offset};
  const unsigned lane = lane_id();
  const u64 nchunks = (N + WCHUNK - 1) / WCHUNK;
  const u64 nwarps = (u64) gridDim.x * (AN_NT / 32);
  u32 mx = 0;
  unsigned long long sum = 0;
  for (u64 c = (u64) blockIdx.x * (AN_NT / 32) + (threadIdx.x >> 5); c < nchunks; c += nwarps) {
    const u64 base = c * WCHUNK + (u64) lane * AN_IPT;
    u64 k[AN_IPT + 2];                   // k[i] = keys[base + i - 1]
    if (base + AN_IPT <= N) {
      const uint4 *p = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
      for (int q = 0; q < AN_IPT / 2; q++) {
        const uint4 v = ld_stream_v4(p + q);
        k[1 + 2 * q] = (u64) v.x | ((u64) v.y << 32);
        k[2 + 2 * q] = (u64) v.z | ((u64) v.w << 32);
      }
    } else {
#pragma unroll
      for (int i = 0; i < AN_IPT; i++) k[1 + i] = base + i < N ? keys[base + i] : 0;
    }
    k[0] = __shfl_up_sync(FULL_MASK, k[AN_IPT], 1);
    k[AN_IPT + 1] = __shfl_down_sync(FULL_MASK, k[1], 1);
    if (lane == 0) k[0] = (base >= 1 && base - 1 < N) ? keys[base - 1] : 0;
    if (lane == 31) k[AN_IPT + 1] = base + AN_IPT < N ? keys[base + AN_IPT] : 0;

    u64 orall = k[0];
#pragma unroll
    for (int i = 1; i < AN_IPT + 2; i++) orall |= k[i];
    u32 unres = 0, lasthead = 0, headbits = 0, unresbits = 0;
    u64 lcpword = 0;
    // (warp-uniform choices: a warp that runs two of the paths pays for both)
    if (c == 0 || (c + 1) * WCHUNK + 1 > N) {                   // the two ends of the array
      analyze_thread<DNA, FILL_LB, false>(k, base, a, unres, lasthead, lcpword, mx, sum, headbits, unresbits);
    } else if (__any_sync(FULL_MASK, (orall & tmask) != 0)) {   // some key met a special
      analyze_thread_inner<DNA, FILL_LB>(k, base, a, unres, lasthead, lcpword, mx, sum, headbits, unresbits);
    } else {
      // all ten keys are full and inside the array: head <=> the keys differ (then in a symbol)
      u32 hb = 0, gapbits = 0, lsum = 0;
      const unsigned codesh = 64u - B * pl;             // DNA: the bucket code is the top 2*pl bits
#pragma unroll
      for (int i = 0; i < AN_IPT; i++) {
        const u64 x = k[i] ^ k[i + 1];
        const u32 l = x ? (u32) __clzll((long long) x) / B : 0u;
        lcpword |= (u64) l << (8 * i);
        lsum += l;
        mx = l > mx ? l : mx;
        if (x) hb |= 1u << i;
        if (FILL_LB && pl > 0 && (x >> codesh) != 0) {
          // a new bucket starts at element i: every code in (code of i-1, code of i] starts there.
          // DNA: one predicated store; empty buckets in between (rare) are filled below
          if (DNA) {
            const u32 c1 = (u32) (k[i + 1] >> codesh), c0 = (u32) (k[i] >> codesh);
            leftborder[c1] = (u32) (sa_offset + base + i);
            if (c1 - c0 > 1u) gapbits |= 1u << i;
          } else {
            const u64 c1 = key_code<DNA>(k[i + 1], pl, K, f);
            for (u64 c0 = key_code<DNA>(k[i], pl, K, f) + 1; c0 <= c1; c0++)
              leftborder[c0] = (u32) (sa_offset + base + i);
          }
        }
      }
      sum += lsum;
      if (k[AN_IPT] != k[AN_IPT + 1]) hb |= 1u << AN_IPT;
      headbits = hb & 0xffu;
      unresbits = (~hb | ~(hb >> 1)) & 0xffu;
      unres = (u32) __popc(unresbits);
      lasthead = headbits ? (u32) base + (31u - (u32) __clz(headbits)) + 1u : 0u;
      if (FILL_LB && DNA && gapbits) {
#pragma unroll
        for (int i = 0; i < AN_IPT; i++)
          if ((gapbits >> i) & 1u) {
            const u32 c1 = (u32) (k[i + 1] >> codesh);
            for (u32 c0 = (u32) (k[i] >> codesh) + 1u; c0 < c1; c0++)
              leftborder[c0] = (u32) (sa_offset + base + i);
          }
      }
    }
    if (base < N) { hbits[base >> 3] = (u8) headbits; ubits[base >> 3] = (u8) unresbits; }
    if (base + AN_IPT <= N) {
      *reinterpret_cast<u64 *>(lcp8 + base) = lcpword;
    } else {
      for (int i = 0; i < AN_IPT && base + i < N; i++) lcp8[base + i] = (u8) (lcpword >> (8 * i));
    }
    const u32 wsum = __reduce_add_sync(FULL_MASK, unres);
    const u32 wmax = __reduce_max_sync(FULL_MASK, lasthead);
    if (lane == 0) {
      const u64 tile = c / (AN_TILE / WCHUNK);
      if (wsum) atomicAdd(&tile_unres[tile], wsum);
      if (wmax) atomicMax(&tile_lasthead[tile], wmax);
    }
  }
  // stats: one atomic per CTA
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(FULL_MASK, sum, d);
    const u32 o = __shfl_xor_sync(FULL_MASK, mx, d);
    mx = o > mx ? o : mx;
  }
  if (lane_id() == 0) { s_sum[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 m2 = 0;
    unsigned long long s2 = 0;
    for (int w = 0; w < AN_NT / 32; w++) { m2 = s_max[w] > m2 ? s_max[w] : m2; s2 += s_sum[w]; }
    if (s2) atomicAdd(&stats->lcpsum, s2);
    if (m2) atomicMax(&stats->maxlcp, m2);
  }
}

// single block: exclusive sum of tile_unres (in place) and exclusive running max of
// tile_lasthead (in place: becomes "last head before this tile", +1 encoded); a thread
// takes 8 consecutive tiles per iteration
__global__ void __launch_bounds__(1024)
k_scan_tiles_sum_max(u32 *tile_unres, u32 *tile_lasthead, u64 ntiles, u64 *total_out)
{
  constexpr int PT = 8;
  __shared__ u32 scratch[1024 / 32 + 1];
  __shared__ u32 s_wlast[32];
  __shared__ u32 carry_sum, carry_max;
  if (threadIdx.x == 0) { carry_sum = 0; carry_max = 0; }
  __syncthreads();
  u64 grand = 0;
  for (u64 base = 0; base < ntiles; base += 1024 * PT) {
    const u64 i0 = base + (u64) threadIdx.x * PT;
    u32 v[PT], h[PT], tsum = 0, tmax = 0;
#pragma unroll
    for (int q = 0; q < PT; q++) {
      v[q] = i0 + q < ntiles ? tile_unres[i0 + q] : 0u;
      h[q] = i0 + q < ntiles ? tile_lasthead[i0 + q] : 0u;
      tsum += v[q];
      tmax = h[q] > tmax ? h[q] : tmax;
    }
    u32 total;
    u32 ex = block_exclusive_sum<1024, u32>(tsum, scratch, &total);
    const u32 incl = block_inclusive_max<1024, u32>(tmax, scratch);
    // exclusive max = inclusive max of the previous thread
    u32 prev = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane_id() == 31) s_wlast[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (lane_id() == 0) prev = threadIdx.x ? s_wlast[(threadIdx.x >> 5) - 1] : 0u;
    const u32 cs = carry_sum, cm = carry_max;
    ex += cs;
    prev = prev > cm ? prev : cm;
#pragma unroll
    for (int q = 0; q < PT; q++) {
      if (i0 + q < ntiles) { tile_unres[i0 + q] = ex; tile_lasthead[i0 + q] = prev; }
      ex += v[q];
      prev = h[q] > prev ? h[q] : prev;
    }
    grand += total;
    __syncthreads();
    if (threadIdx.x == 1023) { carry_sum = cs + total; carry_max = incl > cm ? incl : cm; }
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = grand;
}

// the same over many tiles (one tile per 2048 suffixes: 1.5 M tiles on c4 take a single block 2 ms):
// blocks of 8192 tiles.  mode 0: the sum and the maximum of every block; then the single-block kernel
// above scans those (exclusive sum / exclusive running maximum per block); mode 1: every block scans
// its own tiles in place, carried in by its block's values.
constexpr int STM_BLOCK = 1024 * 8;
__global__ void __launch_bounds__(1024)
k_scan_tiles_blocks(u32 *tile_unres, u32 *tile_lasthead, u64 ntiles, u32 *blk_sum, u32 *blk_max, int mode)
{
  constexpr int PT = 8;
  __shared__ u32 scratch[1024 / 32 + 1];
  __shared__ u32 s_wlast[32];
  const u64 i0 = (u64) blockIdx.x * STM_BLOCK + (u64) threadIdx.x * PT;
  u32 v[PT], h[PT], tsum = 0, tmax = 0;
#pragma unroll
  for (int q = 0; q < PT; q++) {
    v[q] = i0 + q < ntiles ? tile_unres[i0 + q] : 0u;
    h[q] = i0 + q < ntiles ? tile_lasthead[i0 + q] : 0u;
    tsum += v[q];
    tmax = h[q] > tmax ? h[q] : tmax;
  }
  u32 total;
  u32 ex = block_exclusive_sum<1024, u32>(tsum, scratch, &total);
  const u32 incl = block_inclusive_max<1024, u32>(tmax, scratch);
  if (mode == 0) {
    if (threadIdx.x == 1023) { blk_sum[blockIdx.x] = total; blk_max[blockIdx.x] = incl; }
    return;
  }
  u32 prev = __shfl_up_sync(FULL_MASK, incl, 1);
  if (lane_id() == 31) s_wlast[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (lane_id() == 0) prev = threadIdx.x ? s_wlast[(threadIdx.x >> 5) - 1] : 0u;
  const u32 cs = blk_sum[blockIdx.x], cm = blk_max[blockIdx.x];
  ex += cs;
  prev = prev > cm ? prev : cm;
#pragma unroll
  for (int q = 0; q < PT; q++) {
    if (i0 + q < ntiles) { tile_unres[i0 + q] = ex; tile_lasthead[i0 + q] = prev; }
    ex += v[q];
    prev = h[q] > prev ? h[q] : prev;
  }
}

// pass 2: compact the unresolved elements (SA index, position, group head) from the flag
// bytes of pass 1 -- the keys are not read again
__global__ void __launch_bounds__(AN_NT)
k_compact_keys(const u8 *__restrict__ hbits, const u8 *__restrict__ ubits, const u32 *__restrict__ pos, u64 N,
               const u32 *__restrict__ tile_off, const u32 *__restrict__ tile_headbefore, u64 ntiles, u64 M0,
               u32 *__restrict__ uidx, u32 *__restrict__ upos, u32 *__restrict__ ugrp)
{
  __shared__ u32 scratch[AN_NT / 32 + 1];
  __shared__ u32 s_wlast[AN_NT / 32];
  // a block walks over tiles; most hold nothing tied and cost two loads (one block per tile was bound
  // by block scheduling: 1.5 M blocks on c4)
  for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const u32 toff = tile_off[tile];
    const u64 tend = tile + 1 < ntiles ? (u64) tile_off[tile + 1] : M0;
    if (tend == (u64) toff) continue;               // (uniform) nothing tied in this tile
    const u64 base = tile * AN_TILE + (u64) threadIdx.x * AN_IPT;
    const u32 hb8 = base < N ? hbits[base >> 3] : 0u, ub8 = base < N ? ubits[base >> 3] : 0u;
    const u32 unres = (u32) __popc(ub8);
    const u32 lasthead = hb8 ? (u32) base + (31u - (u32) __clz(hb8)) + 1u : 0u;
    u32 total;
    u32 off = block_exclusive_sum<AN_NT, u32>(unres, scratch, &total) + toff;
    // head index carried into this thread = max over previous threads / tiles
    u32 incl = block_inclusive_max<AN_NT, u32>(lasthead, scratch);
    u32 carry = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane_id() == 31) s_wlast[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (lane_id() == 0) carry = threadIdx.x ? s_wlast[(threadIdx.x >> 5) - 1] : 0u;
    const u32 hb = tile_headbefore[tile];
    carry = carry > hb ? carry : hb;          // "+1" encoded index of the current group head
    if (ub8) {
#pragma unroll
      for (int i = 0; i < AN_IPT; i++) {
        const u64 j = base + i;
        if ((hb8 >> i) & 1u) carry = (u32) j + 1u;
        if ((ub8 >> i) & 1u) { uidx[off] = (u32) j; upos[off] = pos[j]; ugrp[off] = carry - 1u; off++; }
      }
    }
    __syncthreads();                           // (s_wlast is reused by the next tile)
  }
}

// ---- prefix doubling ---------------------------------------------------------------------
// key = (group head index, rank of the suffix h positions further)
// Two kernels: the first answers what takes one or two loads (the partner is special or tied itself:
// nearly all partners inside a repeat) and queues the rest; the second runs the binary searches of
// the queued elements with every lane of a warp searching.  (One kernel doing both left 31 lanes of
// nearly every warp waiting for the few that search: 22 % of the round-0 partners of c4 are untied.)
__device__ __forceinline__ void queue_append(bool need, u32 c, u32 *__restrict__ queue, unsigned int *qcount)
{
  const unsigned m = __ballot_sync(FULL_MASK, need);
  if (m == 0) return;
  unsigned base = 0;
  const unsigned lane = lane_id();
  if (lane == (unsigned) __ffs(m) - 1u) base = atomicAdd(qcount, (unsigned) __popc(m));
  base = __shfl_sync(FULL_MASK, base, __ffs(m) - 1);
  if (need) queue[base + (unsigned) __popc(m & lanemask_lt())] = c;
}

// Tie groups of exactly TWO suffixes (chance ties of random sequence: 8.7 M of the 45.7 M ties of c4, and the
// ones whose partner h further is an untied suffix that would have to be SEARCHED) are decided by comparing
// the text of the two suffixes from depth h on, rule R: symbols first; the suffix that meets a special first
// is the larger one; two specials at the same depth by position.  The round then sees two different keys
// (0 for the smaller, 1 for the larger), i.e. a resolved pair.  Undecided after PAIR_WINDOWS windows of 32
// bases (a real repeat): the rank lookup as for every other group.
constexpr int PAIR_WINDOWS = 4;
__device__ __forceinline__ int dna_pair_less(const u64 *__restrict__ words, const u32 *__restrict__ spmask,
                                             u64 pa, u64 pb, u64 off)
{
  for (int it = 0; it < PAIR_WINDOWS; it++, off += 32) {
    const u32 wa = mask_window(spmask, pa + off), wb = mask_window(spmask, pb + off);
    const unsigned la = wa ? (unsigned) __ffs(wa) - 1u : 32u, lb = wb ? (unsigned) __ffs(wb) - 1u : 32u;
    const unsigned lim = la < lb ? la : lb;
    const u64 xa = dna_window(words, pa + off), xb = dna_window(words, pb + off);
    const u64 x = xa ^ xb;
    const unsigned common = x ? (unsigned) (__clzll((long long) x) >> 1) : 32u;
    if (common < lim) return xa < xb ? 1 : 0;              // (the first differing base decides: MSB-first words)
    if (lim < 32u) {
      if (la != lb) return la > lb ? 1 : 0;                // b meets a special first: a is the smaller one
      return pa < pb ? 1 : 0;                              // both meet a special here: by position
    }
  }
  return -1;
}
// is element c one of a group of exactly two?  *partner = the other one
__device__ __forceinline__ bool tie_pair_partner(const u32 *__restrict__ ugrp, u64 M, u64 c, u64 *partner)
{
  const u32 g = ugrp[c];
  const bool prev_same = c > 0 && ugrp[c - 1] == g, next_same = c + 1 < M && ugrp[c + 1] == g;
  if (prev_same == next_same) return false;                // alone (cannot be) or inside a longer run
  if (next_same) { if (c + 2 < M && ugrp[c + 2] == g) return false; *partner = c + 1; return true; }
  if (c >= 2 && ugrp[c - 2] == g) return false;
  *partner = c - 1;
  return true;
}

template <bool DNA>
__global__ void k_build_dkeys(RankMap<DNA> rm, const u32 *__restrict__ upos, const u32 *__restrict__ ugrp,
                              u64 M, u64 h, u64 *__restrict__ dkeys, u32 *__restrict__ queue, unsigned int *qcount,
                              int pairs_by_text)
{
  const u64 stride = (u64) gridDim.x * blockDim.x;
  for (u64 c0 = blockIdx.x * (u64) blockDim.x; c0 < M; c0 += stride) {      // (whole warps stay in the loop)
    const u64 c = c0 + threadIdx.x;
    bool need = false;
    if (c < M) {
      u32 r;
      u64 partner;
      int less = -1;
      if (DNA && pairs_by_text && tie_pair_partner(ugrp, M, c, &partner))
        less = dna_pair_less(rm.src.words, rm.src.spmask, (u64) upos[c], (u64) upos[partner], h);
      if (less >= 0) dkeys[c] = ((u64) ugrp[c] << 32) | (u64) (less ? 0u : 1u);
      else if (rm.get_fast((u64) upos[c] + h, r)) dkeys[c] = ((u64) ugrp[c] << 32) | (u64) r;
      else need = true;
    }
    queue_append(need, (u32) c, queue, qcount);
  }
}
template <bool DNA>
__global__ void k_build_dkeys_search(RankMap<DNA> rm, const u32 *__restrict__ upos, const u32 *__restrict__ ugrp,
                                     u64 h, u64 *__restrict__ dkeys, const u32 *__restrict__ queue,
                                     const unsigned int *__restrict__ qcount)
{
  const u64 nq = *qcount;
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < nq; i += (u64) gridDim.x * blockDim.x) {
    const u32 c = queue[i];
    dkeys[c] = ((u64) ugrp[c] << 32) | (u64) rm.search((u64) upos[c] + h);
  }
}

// bitmap of the tied positions
__global__ void k_tied_bits(const u32 *__restrict__ uidx0, u64 M0, const u32 *__restrict__ sa,
                            u32 *__restrict__ tbits)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 p = sa[uidx0[c]];
    atomicOr(&tbits[p >> 5], 1u << (p & 31u));
  }
}
// the 16-byte words of the rank map (RankMap::rw); spre / tbits / tpre may be null (no specials,
// no ties)
__global__ void k_pack_rankwords(const u32 *__restrict__ spmask, const u32 *__restrict__ spre,
                                 const u32 *__restrict__ tbits, const u32 *__restrict__ tpre, u64 nw,
                                 uint4 *__restrict__ rw)
{
  for (u64 w = blockIdx.x * (u64) blockDim.x + threadIdx.x; w < nw; w += (u64) gridDim.x * blockDim.x)
    rw[w] = make_uint4(spmask[w], spre ? spre[w] : 0u, tbits ? tbits[w] : 0u, tpre ? tpre[w] : 0u);
}
// ranks of the initially tied suffixes: first each as if resolved (its own index), then
// the still tied ones with their group head
template <bool DNA>
__global__ void k_trank_resolved(RankMap<DNA> rm, const u32 *__restrict__ uidx0, u64 M0)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 j = uidx0[c];
    rm.set(rm.sa[j], (u32) (rm.sa_offset + j));
  }
}
template <bool DNA>
__global__ void k_trank_tied(RankMap<DNA> rm, const u32 *__restrict__ upos, const u32 *__restrict__ ugrp, u64 M)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M; c += (u64) gridDim.x * blockDim.x)
    rm.set(upos[c], (u32) (rm.sa_offset + ugrp[c]));
}

// text-driven round: key = (group head index, filled key of the next symbols in the
// short format g, top 32 bits).  A continuation that starts at a special sorts last
// (all filler, u = 0).  Keys whose tail field is non-zero met a special: resolved.
template <bool DNA>
__global__ void k_build_tkeys(TextSrc<DNA> src, KeyFmt g, const u32 *__restrict__ upos,
                              const u32 *__restrict__ ugrp, u64 M, u64 h, u64 *__restrict__ dkeys)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M; c += (u64) gridDim.x * blockDim.x) {
    u64 key;
    if (!src.make_key_fmt((u64) upos[c] + h, key, g)) key = g.symmask() | ((u64) g.m << g.sh);
    dkeys[c] = ((u64) ugrp[c] << 32) | (key >> 32);
  }
}

// `tm`: bits of a sort key that, when set, mark it as resolved whatever its neighbours
// are (the tail field of a text-driven key; 0 for rank keys)
__global__ void __launch_bounds__(AN_NT)
k_analyze_dkeys(const u64 *__restrict__ dkeys, u64 M, u64 tm, u32 *__restrict__ tile_unres,
                u32 *__restrict__ tile_lasthead)
{
  __shared__ u32 scratch[AN_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * AN_TILE + (u64) threadIdx.x * AN_IPT;
  u64 k[AN_IPT + 2];
#pragma unroll
  for (int i = 0; i < AN_IPT + 2; i++) {
    const u64 c = base + i;
    k[i] = (c >= 1 && c - 1 < M) ? dkeys[c - 1] : 0;
  }
  u32 unres = 0, lasthead = 0;
#pragma unroll
  for (int i = 0; i < AN_IPT; i++) {
    const u64 c = base + i;
    if (c >= M) break;
    const bool head = c == 0 || key_head(k[i], k[i + 1], tm);
    const bool nexthead = c + 1 >= M || key_head(k[i + 1], k[i + 2], tm);
    if (head) lasthead = (u32) c + 1u;
    if (!head || !nexthead) unres++;
  }
  u32 total;
  block_exclusive_sum<AN_NT, u32>(unres, scratch, &total);
  const u32 lh = block_inclusive_max<AN_NT, u32>(lasthead, scratch);
  if (threadIdx.x == AN_NT - 1) { tile_unres[blockIdx.x] = total; tile_lasthead[blockIdx.x] = lh; }
}

// write the refined order back, update ranks, record the doubling level of new group
// heads (a lower bound of their lcp), compact what is still tied
template <bool DNA>
__global__ void __launch_bounds__(AN_NT)
k_apply_dkeys(const u64 *__restrict__ dkeys, const u32 *__restrict__ spos /* sorted upos */,
              const u32 *__restrict__ uidx, u64 M, u64 tm,
              const u32 *__restrict__ tile_off, const u32 *__restrict__ tile_headbefore,
              u32 *__restrict__ sa, RankMap<DNA> rm, int have_ranks, u8 *__restrict__ lcp8, u8 level,
              u64 sa_offset, u32 *__restrict__ nidx, u32 *__restrict__ npos,
              u32 *__restrict__ ngrp, DevStats *stats)
{
  __shared__ u32 scratch[AN_NT / 32 + 1];
  const u64 base = (u64) blockIdx.x * AN_TILE + (u64) threadIdx.x * AN_IPT;
  u64 k[AN_IPT + 2];
#pragma unroll
  for (int i = 0; i < AN_IPT + 2; i++) {
    const u64 c = base + i;
    k[i] = (c >= 1 && c - 1 < M) ? dkeys[c - 1] : 0;
  }
  bool head[AN_IPT + 1];
  u32 unres = 0, lasthead = 0;
#pragma unroll
  for (int i = 0; i < AN_IPT + 1; i++) {
    const u64 c = base + i;
    head[i] = c == 0 || c >= M || key_head(k[i], k[i + 1], tm);
  }
#pragma unroll
  for (int i = 0; i < AN_IPT; i++) {
    const u64 c = base + i;
    if (c >= M) break;
    if (head[i]) lasthead = (u32) c + 1u;
    if (!head[i] || !head[i + 1]) unres++;
  }
  u32 total;
  u32 off = block_exclusive_sum<AN_NT, u32>(unres, scratch, &total) + tile_off[blockIdx.x];
  u32 incl = block_inclusive_max<AN_NT, u32>(lasthead, scratch);
  u32 carry = __shfl_up_sync(FULL_MASK, incl, 1);
  __shared__ u32 s_wlast[AN_NT / 32];
  if (lane_id() == 31) s_wlast[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (lane_id() == 0) carry = threadIdx.x ? s_wlast[(threadIdx.x >> 5) - 1] : 0u;
  const u32 hb = tile_headbefore[blockIdx.x];
  carry = carry > hb ? carry : hb;
#pragma unroll
  for (int i = 0; i < AN_IPT; i++) {
    const u64 c = base + i;
    if (c >= M) break;
    if (head[i]) carry = (u32) c + 1u;
    const u32 hc = carry - 1u;                 // compacted index of the group head
    const u32 g = uidx[hc];                    // its SA slot = new rank of the group
    const u32 j = uidx[c];
    const u32 p = spos[c];
    sa[j] = p;
    if (have_ranks) rm.set(p, (u32) (sa_offset + g));   // (no ranks are kept during text-driven rounds)
    if (p == 0) stats->longest = sa_offset + j;
    const u32 oldgrp = (u32) (k[i + 1] >> 32);
    if (head[i] && oldgrp != j) lcp8[j] = level;   // split off in this round
    if (!head[i] || !head[i + 1]) { nidx[off] = j; npos[off] = p; ngrp[off] = g; off++; }
  }
}

// ---- prefix doubling across code ranges (multi-GPU / -parts): rank exchange -------------------
// Each range owns the ranks of its own suffixes (its slice of the inverse suffix array).  In a
// doubling round the rank of position q = p + h is needed; the owner of q is the range whose
// code interval contains the filled key of q.  Ranks of special positions and of q in the own
// range are read directly, the others are routed to their owner.
constexpr int MAX_RANGES = 64;
struct RangeBounds { u64 first_key[MAX_RANGES]; int n; int mine; };
constexpr u32 OWNER_LOCAL = 0xffu;

template <bool DNA>
__global__ void k_round_classify(TextSrc<DNA> src, const u32 *__restrict__ upos, u64 M, u64 h,
                                 RangeBounds rb, RankMap<DNA> rm, u32 *__restrict__ ranks,
                                 u8 *__restrict__ owner, unsigned int *__restrict__ counts)
{
  __shared__ unsigned int s_cnt[MAX_RANGES];
  for (int i = threadIdx.x; i < MAX_RANGES; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M; c += (u64) gridDim.x * blockDim.x) {
    const u64 q = (u64) upos[c] + h;
    u64 key;
    u32 o = OWNER_LOCAL;
    if (src.make_key(q, key)) {               // regular position: find its range
      int lo = 0, hi = rb.n - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rb.first_key[mid] <= key) lo = mid; else hi = mid - 1; }
      if (lo != rb.mine) o = (u32) lo;
    }
    owner[c] = (u8) o;
    if (o == OWNER_LOCAL) ranks[c] = rm.get(q);
    else atomicAdd(&s_cnt[o], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rb.n; i += blockDim.x) if (s_cnt[i]) atomicAdd(&counts[i], s_cnt[i]);
}

// fill the send buffer: segment of range o starts at offs[o]
__global__ void k_round_fill(const u32 *__restrict__ upos, const u8 *__restrict__ owner, u64 M, u64 h,
                             const unsigned int *__restrict__ offs, unsigned int *__restrict__ fill,
                             u32 *__restrict__ sendq, u32 *__restrict__ sendidx)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M; c += (u64) gridDim.x * blockDim.x) {
    const u32 o = owner[c];
    if (o == OWNER_LOCAL) continue;
    const u32 slot = offs[o] + atomicAdd(&fill[o], 1u);
    sendq[slot] = (u32) ((u64) upos[c] + h);
    sendidx[slot] = (u32) c;
  }
}

template <bool DNA>
__global__ void k_rank_lookup(RankMap<DNA> rm, const u32 *__restrict__ q, u64 cnt, u32 *__restrict__ out)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < cnt; i += (u64) gridDim.x * blockDim.x)
    out[i] = rm.get(q[i]);
}

__global__ void k_round_scatter(const u32 *__restrict__ answers, const u32 *__restrict__ sendidx, u64 cnt,
                                u32 *__restrict__ ranks)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < cnt; i += (u64) gridDim.x * blockDim.x)
    ranks[sendidx[i]] = answers[i];
}

__global__ void k_build_dkeys_ranks(const u32 *__restrict__ ugrp, const u32 *__restrict__ ranks, u64 M,
                                    u64 *__restrict__ dkeys)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M; c += (u64) gridDim.x * blockDim.x)
    dkeys[c] = ((u64) ugrp[c] << 32) | (u64) ranks[c];
}

// the filled key of the smallest suffix of bucket `code` (all further symbols smallest, full)

// ---- exact lcp of the deep pairs (Kasai-style comparison from a proven lower bound) ----
// depth[r] = common prefix proven for a pair that was separated in refinement round r
struct DepthTab { u64 d[64]; };

// A thread walks at most DEEP_SOLO windows of 32 bases on its own; pairs that are still equal
// then (long exact repeats: thousands of windows) are queued as (c : 32 | offset : 32) and
// finished by k_deep_lcp_long, one warp per pair and 32 windows per step.
constexpr int DEEP_SOLO = 8;
constexpr u32 DEEP_PENDING = 0xfffffffeu;

template <bool DNA>
__global__ void k_deep_lcp(const u32 *__restrict__ uidx0, const u32 *__restrict__ ugrp0, u64 M0,
                           const u32 *__restrict__ sa, const u64 *__restrict__ words,
                           const u8 *__restrict__ bytes, const u32 *__restrict__ spmask,
                           u64 n, DepthTab depth, u8 *__restrict__ lcp8, u32 *__restrict__ ulcp,
                           DevStats *stats, u64 *__restrict__ queue, unsigned long long *qcount)
{
  u32 mx = 0;
  unsigned long long sum = 0, large = 0;
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 j = uidx0[c];
    u32 l = 0xffffffffu;                       // marker: not a deep entry
    if (ugrp0[c] != j) {
      const u64 a = sa[j - 1], b = sa[j];
      u64 off = depth.d[lcp8[j] & 63u];        // proven common prefix
      bool done = false;
      if (DNA) {
        for (int it = 0; it < DEEP_SOLO; it++) {
          const u32 wa = mask_window(spmask, a + off), wb = mask_window(spmask, b + off);
          const u32 wab = wa | wb;
          const unsigned lim = wab ? (unsigned) __ffs(wab) - 1u : 32u;
          const u64 x = dna_window(words, a + off) ^ dna_window(words, b + off);
          const unsigned common = x ? (unsigned) (__clzll((long long) x) >> 1) : 32u;
          if (common < lim) { off += common; done = true; break; }
          if (lim < 32u) { off += lim; done = true; break; }
          off += 32;
        }
      } else {
        for (;;) {
          const u8 ca = bytes[a + off], cb = bytes[b + off];
          if (ca >= 254u || cb >= 254u || ca != cb) break;
          off++;
        }
        done = true;
      }
      if (done) {
        l = (u32) off;
        lcp8[j] = (u8) (l < 255u ? l : 255u);
        mx = l > mx ? l : mx;
        sum += l;
        if (l >= 255u) large++;
      } else {
        l = DEEP_PENDING;
        queue[atomicAdd(qcount, 1ull)] = (c << 32) | off;
      }
    }
    ulcp[c] = l;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(FULL_MASK, sum, d);
    large += __shfl_xor_sync(FULL_MASK, large, d);
    const u32 o = __shfl_xor_sync(FULL_MASK, mx, d);
    mx = o > mx ? o : mx;
  }
  if (lane_id() == 0) {
    if (sum) atomicAdd(&stats->lcpsum, sum);
    if (large) atomicAdd(&stats->numlarge, large);
    if (mx) atomicMax(&stats->maxlcp, mx);
  }
}

// one warp per queued pair: lane i compares the window 32*i bases further
__global__ void k_deep_lcp_long(const u64 *__restrict__ queue, const unsigned long long *__restrict__ qcount,
                                const u32 *__restrict__ uidx0, const u32 *__restrict__ sa,
                                const u64 *__restrict__ words, const u32 *__restrict__ spmask, u64 n,
                                u8 *__restrict__ lcp8, u32 *__restrict__ ulcp, DevStats *stats)
{
  const unsigned lane = lane_id();
  const u64 nq = *qcount;
  const u64 warps = ((u64) gridDim.x * blockDim.x) >> 5;
  u32 mx = 0;
  unsigned long long sum = 0, large = 0;
  for (u64 q = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += warps) {
    const u64 e = queue[q];
    const u64 c = e >> 32;
    u64 off = e & 0xffffffffull;
    const u32 j = uidx0[c];
    const u64 a = sa[j - 1], b = sa[j];
    for (;;) {
      const u64 pa = a + off + 32u * lane, pb = b + off + 32u * lane;
      unsigned stop = 0;                         // equal regular bases in this lane's window
      if (pa < n && pb < n) {                    // (the end of the text acts as a special)
        const u32 wab = mask_window(spmask, pa) | mask_window(spmask, pb);
        const unsigned lim = wab ? (unsigned) __ffs(wab) - 1u : 32u;
        const u64 x = dna_window(words, pa) ^ dna_window(words, pb);
        const unsigned common = x ? (unsigned) (__clzll((long long) x) >> 1) : 32u;
        stop = common < lim ? common : lim;
      }
      const unsigned ended = __ballot_sync(FULL_MASK, stop < 32u);
      if (ended) {
        const unsigned first = (unsigned) __ffs(ended) - 1u;
        off += 32u * first + __shfl_sync(FULL_MASK, stop, first);
        break;
      }
      off += 1024;
    }
    if (lane == 0) {
      const u32 l = (u32) off;
      lcp8[j] = (u8) (l < 255u ? l : 255u);
      ulcp[c] = l;
      mx = l > mx ? l : mx;
      sum += l;
      if (l >= 255u) large++;
    }
  }
  if (lane == 0) {
    if (sum) atomicAdd(&stats->lcpsum, sum);
    if (large) atomicAdd(&stats->numlarge, large);
    if (mx) atomicMax(&stats->maxlcp, mx);
  }
}

// .llv exception list (lcpoverflow.h:25-29): flags, then scan (generic), then emit
__global__ void k_llv_flags(const u32 *__restrict__ ulcp, u64 M0, u32 *__restrict__ flags)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 l = ulcp[c];
    flags[c] = (l != 0xffffffffu && l >= 255u) ? 1u : 0u;
  }
}
__global__ void k_llv_emit(const u32 *__restrict__ ulcp, const u32 *__restrict__ uidx0, u64 M0,
                           const u32 *__restrict__ offs, u64 sa_offset, u64 *__restrict__ llv)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 l = ulcp[c];
    if (l != 0xffffffffu && l >= 255u) {
      llv[2 * (u64) offs[c]] = sa_offset + uidx0[c];
      llv[2 * (u64) offs[c] + 1] = l;
    }
  }
}

// position of suffix 0 in a sorted key array (first element with key >= key0; suffix 0
// has the smallest position so it leads any stable tie)
template <bool DNA>
__global__ void k_find_longest(TextSrc<DNA> src, const u64 *__restrict__ keys, u64 N,
                               u64 sa_offset, DevStats *stats)
{
  u64 key0;
  if (!src.make_key(0, key0)) return;
  u64 lo = 0, hi = N;
  while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (keys[mid] < key0) lo = mid + 1; else hi = mid; }
  if (lo < N && keys[lo] == key0) stats->longest = sa_offset + lo;
}

// Coarse bucket counts for cutting code ranges over several GPUs: the first `plc` symbols of
// the filled key (at most 4096 codes), counted in shared memory -- no global atomic per
// suffix.  Positions [first, end).
constexpr int CC_MAXCODES = 4096;
// A thread takes a chunk of 16 consecutive positions (keys from TextSrc::gen_block_fast); equal
// codes of neighbouring positions (low-complexity sequence) are added as one run.
template <bool DNA>
__global__ void __launch_bounds__(256)
k_count_coarse(TextSrc<DNA> src, u64 first, u64 end, unsigned plc, unsigned K, u32 ncoarse,
               u32 *__restrict__ cnt)
{
  __shared__ u32 s_c[CC_MAXCODES];
  for (u32 i = threadIdx.x; i < ncoarse; i += blockDim.x) s_c[i] = 0;
  __syncthreads();
  const u64 c0 = first >> 4, c1 = (end + 15) >> 4;          // chunks [16c, 16c + 16) that meet [first, end)
  for (u64 c = c0 + blockIdx.x * (u64) blockDim.x + threadIdx.x; c < c1; c += (u64) gridDim.x * blockDim.x) {
    const u64 p = c << 4;
    u32 prev = 0, run = 0;
    if (p >= first && p + 16 <= end) {
      u64 keys[16];
      src.gen_block_fast(p, keys);                           // (pos0 = 0: item = position; p is 16-aligned)
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (keys[i] == ~0ull) continue;
        const u32 code = key_code_prefix<DNA>(keys[i], plc, K, src.f);
        if (run && code == prev) run++;
        else { if (run) atomicAdd(&s_c[prev], run); prev = code; run = 1; }
      }
    } else {
      for (u64 q = p < first ? first : p; q < end && q < p + 16; q++) {
        u64 key;
        if (!src.make_key(q, key)) continue;
        const u32 code = key_code_prefix<DNA>(key, plc, K, src.f);
        if (run && code == prev) run++;
        else { if (run) atomicAdd(&s_c[prev], run); prev = code; run = 1; }
      }
    }
    if (run) atomicAdd(&s_c[prev], run);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < ncoarse; i += blockDim.x) if (s_c[i]) atomicAdd(&cnt[i], s_c[i]);
}

// gt_suftabparts_new (sfx-partssuf.c:172-347) on the bucket table in HBM: part p ends at
// the first code whose right border reaches the running target (gt_bcktab_findfirstlarger,
// bcktab.c:1322-1381).  out[4*p..] = mincode, maxcode, sa_offset, width; *nout = parts.
__global__ void k_split_ranges(const u32 *__restrict__ lb, u64 ncodes, unsigned numofparts,
                               unsigned long long *__restrict__ out, unsigned *__restrict__ nout)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const u64 total = lb[ncodes];
  unsigned np = 0;
  if (numofparts <= 1 || total <= numofparts || ncodes == 1) {
    out[0] = 0; out[1] = ncodes - 1; out[2] = 0; out[3] = total; *nout = 1; return;
  }
  const u64 width = total / numofparts, rem = total % numofparts;
  u64 mincode = 0, target = 0;
  for (unsigned part = 0; part < numofparts && mincode < ncodes; part++) {
    target += width + (part < rem ? 1 : 0);
    u64 maxcode;
    if (part == numofparts - 1) maxcode = ncodes - 1;
    else {
      u64 lo = 0, hi = ncodes;                 // first c with lb[c+1] >= target
      while (lo < hi) { const u64 mid = (lo + hi) >> 1; if ((u64) lb[mid + 1] < target) lo = mid + 1; else hi = mid; }
      maxcode = lo < mincode ? mincode : lo;
      if (maxcode > ncodes - 1) maxcode = ncodes - 1;
    }
    const u64 w = (u64) lb[maxcode + 1] - (u64) lb[mincode];
    if (w > 0 || part == numofparts - 1) {
      out[4 * np] = mincode; out[4 * np + 1] = maxcode; out[4 * np + 2] = lb[mincode]; out[4 * np + 3] = w; np++;
    }
    mincode = maxcode + 1;
  }
  if (np > 0 && out[4 * (np - 1) + 1] != ncodes - 1) {   // the last part reaches the last code
    out[4 * (np - 1) + 1] = ncodes - 1;
    out[4 * (np - 1) + 3] = (u64) lb[ncodes] - (u64) lb[out[4 * (np - 1)]];
  }
  unsigned keep = 0;                                      // drop empty parts
  for (unsigned p = 0; p < np; p++)
    if (out[4 * p + 3] > 0) { for (int q = 0; q < 4; q++) out[4 * keep + q] = out[4 * p + q]; keep++; }
  if (keep == 0) { out[0] = 0; out[1] = ncodes - 1; out[2] = 0; out[3] = total; keep = 1; }
  *nout = keep;
}

// keys of a list of (valid, non-special) text positions; tailbits (optional): bit i of word w = the key
// of element 32w + i carries a tail field
template <bool DNA>
__global__ void k_keys_from_positions(TextSrc<DNA> src, const u32 *__restrict__ pos, u64 count,
                                      u64 *__restrict__ keys, u32 *__restrict__ tailbits = nullptr)
{
  const u64 tm = src.f.tailmask();
  for (u64 i0 = blockIdx.x * (u64) blockDim.x; i0 < count; i0 += (u64) gridDim.x * blockDim.x) {   // (whole warps)
    const u64 i = i0 + threadIdx.x;
    u64 k = 0;
    if (i < count) {
      src.make_key_fmt(pos[i], k, src.f);
      keys[i] = k;
    }
    if (tailbits) {
      const unsigned b = __ballot_sync(FULL_MASK, i < count && (k & tm) != 0);
      if (lane_id() == 0 && i < count) tailbits[i >> 5] = b;
    }
  }
}

// pairs already in memory without the keys that carry a tail (those are sorted apart, TailSrc)
struct PairSrcNoTail {
  static constexpr bool ALWAYS_VALID = false;
  static constexpr bool BLOCKED_GEN = false;
  const u64 *keys;
  const u32 *vals;
  u64 tailmask;
  __device__ __forceinline__ bool load_key(u64 idx, u64 &k) const
  { k = keys[idx]; return (k & tailmask) == 0; }
  __device__ __forceinline__ u32 load_val(u64 idx) const { return vals[idx]; }
};

__global__ void k_gather_pairs(const u32 *__restrict__ idx, u64 count, const u64 *__restrict__ keys,
                               const u32 *__restrict__ vals, u64 *__restrict__ okeys, u32 *__restrict__ ovals)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x) {
    const u32 j = idx[i];
    okeys[i] = keys[j]; ovals[i] = vals[j];
  }
}

// ---- bwttab (bwttab2file, /root/reference/src/match/sfx-run.c:173-210) -------------------
// bwt[j] = encoded symbol before suffix suf[j] (0..K-1, 254 wildcard, 255 separator);
// UNDEFBWTCHAR (= 254, chardef.h:65) for the suffix that starts at 0
__global__ void k_set_bits(const u64 *__restrict__ positions, u64 count, u32 *__restrict__ bits)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
    atomicOr(&bits[positions[i] >> 5], 1u << (positions[i] & 31u));
}

template <bool DNA>
__global__ void k_bwt(const u32 *__restrict__ sa, u64 count, const u64 *__restrict__ words,
                      const u8 *__restrict__ bytes, const u32 *__restrict__ spmask,
                      const u32 *__restrict__ sepbits, u8 *__restrict__ out)
{
  for (u64 j = blockIdx.x * (u64) blockDim.x + threadIdx.x; j < count; j += (u64) gridDim.x * blockDim.x) {
    const u64 p = sa[j];
    u8 c = 254;
    if (p > 0) {
      const u64 q = p - 1;
      if (!DNA) c = bytes[q];
      else if ((spmask[q >> 5] >> (q & 31u)) & 1u) c = (sepbits && ((sepbits[q >> 5] >> (q & 31u)) & 1u)) ? 255 : 254;
      else c = (u8) ((words[q >> 5] >> (62u - 2u * (unsigned) (q & 31u))) & 3u);
    }
    out[j] = c;
  }
}

// ---- stand-alone sorts of 64-bit records (gt_radixsort_inplace_*, src/core/radix_sort.h) ----
// keys[i] = component `comp` of record perm[i] (perm null: record i), vals[i] = its record index
__global__ void k_records_to_pairs(const u64 *__restrict__ rec, unsigned width, unsigned comp,
                                   const u32 *__restrict__ perm, u64 count, u64 *__restrict__ keys,
                                   u32 *__restrict__ vals)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x) {
    const u32 r = perm ? perm[i] : (u32) i;
    keys[i] = rec[(u64) r * width + comp];
    vals[i] = r;
  }
}
// out[i] = record perm[i]
__global__ void k_gather_records(const u64 *__restrict__ rec, unsigned width, const u32 *__restrict__ perm,
                                 u64 count, u64 *__restrict__ out)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x) {
    const u64 r = perm[i];
    for (unsigned c = 0; c < width; c++) out[i * width + c] = rec[r * width + c];
  }
}

// the tied entries with a suffix-array index below `limit` as (index : 32 | position : 32): the part of the
// table that left for the host before the refinement had finished (gtb_esa_run_to_host).  uidx0 ascends,
// so they are a prefix of the list; *count = its length
__global__ void k_patch_gather(const u32 *__restrict__ uidx0, u64 M0, const u32 *__restrict__ sa, u64 limit,
                               u64 *__restrict__ out, unsigned int *count)
{
  for (u64 c = blockIdx.x * (u64) blockDim.x + threadIdx.x; c < M0; c += (u64) gridDim.x * blockDim.x) {
    const u32 j = uidx0[c];
    if ((u64) j < limit) {
      out[c] = ((u64) j << 32) | (u64) sa[j];
      if (c + 1 == M0 || (u64) uidx0[c + 1] >= limit) *count = (unsigned int) (c + 1);
    }
  }
}

__global__ void k_widen_u32_u64(const u32 *__restrict__ in, u64 *__restrict__ out, u64 count)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
    out[i] = in[i];
}

// ---- read modes (-dir rev | cpl | rcl) ----------------------------------------------------
// The reference reads the sequence through a GtReadmode (gt_encseq_get_encoded_char,
// /root/reference/src/core/encseq.c:6094-6140; extraction of k-mers in reverse / complement,
// src/match/sfx-mapped4.gen:33-86): position i of the text the suffixes are taken from is
// text[n-1-i] (rev, rcl), complemented (cpl, rcl: A<->T, C<->G, i.e. 3 - c; specials stay).
// Here the transformed sequence is materialised once in HBM and everything downstream is
// unchanged.
__device__ __forceinline__ u64 reverse_bases(u64 w)      // order of the 32 2-bit symbols reversed
{
  w = __brevll(w);
  return ((w & 0x5555555555555555ull) << 1) | ((w >> 1) & 0x5555555555555555ull);
}
// `in`: n bases, zero padded by >= 2 words; `out`: nwords_out words, zero beyond base n-1
__global__ void k_readmode_words(const u64 *__restrict__ in, u64 *__restrict__ out, u64 n, u64 nwords_out,
                                 int rev, int cpl)
{
  for (u64 w = blockIdx.x * (u64) blockDim.x + threadIdx.x; w < nwords_out; w += (u64) gridDim.x * blockDim.x) {
    u64 v = 0;
    const u64 first = w * 32;                    // output bases first .. first + 31
    if (first < n) {
      if (!rev) v = in[w];
      else {
        const u64 hi = n - 1 - first;            // input base of output base `first`; the word
        const u64 win = hi >= 31 ? dna_window(in, hi - 31)            // holds input bases hi-31 .. hi
                                 : in[0] >> (2u * (unsigned) (31 - hi));
        v = reverse_bases(win);
      }
      if (cpl) v = ~v;
      const u64 valid = n - first;
      if (valid < 32) v &= ~0ull << (2u * (unsigned) (32 - valid));
    }
    out[w] = v;
  }
}
// byte path: one symbol per byte; cpl only for the 4-letter alphabet (regular symbols 0..3)
__global__ void k_readmode_bytes(const u8 *__restrict__ in, u8 *__restrict__ out, u64 n, int rev, int cpl)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
    const u8 c = in[rev ? n - 1 - i : i];
    out[i] = (cpl && c < 4) ? (u8) (3 - c) : c;
  }
}

__global__ void k_set_u32(u32 *p, u32 v) { *p = v; }

// ---- order-dependent checksum of a result table (genometools_b200/mixhash.py) -------------
//   H = sum_i fin((i + 1) * C1 xor fin(v_i + C2))  mod 2^64,  i = global index of entry v_i,
// fin = splitmix64 finaliser.  The sum composes across shards (each GPU hashes its own part
// with its global offset), the index makes it order-dependent: the same function over the
// reference's files is what a full-size run is compared with (tests/golden/config_md5.json).
__host__ __device__ __forceinline__ u64 mh_fin(u64 z)
{
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ u64 mh_term(u64 index, u64 value)
{
  return mh_fin(((index + 1ull) * 0x9E3779B97F4A7C15ull) ^ mh_fin(value + 0xC2B2AE3D27D4EB4Full));
}
template <typename T>
__global__ void __launch_bounds__(256)
k_mixhash(const T *__restrict__ v, u64 count, u64 index_base, unsigned long long *__restrict__ out)
{
  u64 sum = 0;
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x)
    sum += mh_term(index_base + i, (u64) v[i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
  if (lane_id() == 0 && sum) atomicAdd(out, (unsigned long long) sum);
}

// leftborder: after the scan entry c holds the start of bucket c; entry C = number of
// non-special suffixes (bcktab.c:1274-1304 keeps the same convention on file)
} // namespace gtb
