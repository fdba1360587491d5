// gtb_widen.cpp -- uint32 -> uint64 widening of suffix-table chunks on host cores (gtb_hostio.cuh),
// compiled by the host compiler alone (no CUDA front end: the AVX-512 intrinsics headers are not its
// business).  .suf holds uint64 entries (gt_suffixsortspace_to_file,
// /root/reference/src/match/sfx-suffixgetset.c:462-477), the table in HBM uint32: 4 bytes per entry
// cross PCIe, the other 4 are zeros written here with non-temporal stores.  Widest vector unit the
// CPU has (checked once at run time): one 64-byte store fills a whole cache line of the destination.
#include <immintrin.h>
#include <stdint.h>

namespace {

void widen_sse2(const uint32_t *src, uint64_t *dst, uint64_t n)
{
  uint64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) { dst[i] = src[i]; i++; }
  const __m128i z = _mm_setzero_si128();
  for (; i + 8 <= n; i += 8) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 4));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_unpacklo_epi32(a, z));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 2), _mm_unpackhi_epi32(a, z));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_unpacklo_epi32(b, z));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 6), _mm_unpackhi_epi32(b, z));
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

__attribute__((target("avx2")))
void widen_avx2(const uint32_t *src, uint64_t *dst, uint64_t n)
{
  uint64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u)) { dst[i] = src[i]; i++; }
  for (; i + 8 <= n; i += 8) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 4));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepu32_epi64(a));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_cvtepu32_epi64(b));
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

__attribute__((target("avx512f")))
void widen_avx512(const uint32_t *src, uint64_t *dst, uint64_t n)
{
  uint64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u)) { dst[i] = src[i]; i++; }
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8));
    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_cvtepu32_epi64(a));
    _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i + 8), _mm512_cvtepu32_epi64(b));
  }
  for (; i < n; i++) dst[i] = src[i];
  _mm_sfence();
}

typedef void (*widen_fn)(const uint32_t *, uint64_t *, uint64_t);

widen_fn pick()
{
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return widen_avx512;
  if (__builtin_cpu_supports("avx2")) return widen_avx2;
  return widen_sse2;
}

} // namespace

// dst[i] = src[i] for i < n, streaming (the destination is written once and not read here).
// which: 0 = the widest unit of this CPU, 1 = SSE2, 2 = AVX2, 3 = AVX-512 (tools/widen_bench)
extern "C" void gtb_widen_u32_u64(const uint32_t *src, uint64_t *dst, uint64_t n, int which)
{
  static const widen_fn best = pick();
  if (which == 1) widen_sse2(src, dst, n);
  else if (which == 2) widen_avx2(src, dst, n);
  else if (which == 3) widen_avx512(src, dst, n);
  else best(src, dst, n);
}
