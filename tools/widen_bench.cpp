// tools/widen_bench.cpp -- host-only: how fast do T threads widen uint32 -> uint64 with non-temporal
// stores on this machine (the host half of gtb_esa_copy_suftab_u64)?  Staging buffers of 8 MiB as in
// the library (they stay in the last-level cache), 2 GiB of output per run.
//   g++ -O2 -pthread tools/widen_bench.cpp genometools_b200/csrc/gtb_widen.cpp -o tools/widen_bench
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
extern "C" void gtb_widen_u32_u64(const uint32_t *src, uint64_t *dst, uint64_t n, int which);
int main(int argc, char **argv)
{
  const uint64_t total = (uint64_t) 1 << 28;                 // entries: 2 GiB of uint64
  const uint64_t chunk = (8u << 20) / 4;                     // entries per staging buffer
  uint64_t *dst = (uint64_t *) aligned_alloc(4096, total * 8);
  uint32_t *stage = (uint32_t *) aligned_alloc(4096, 4 * chunk * 4);
  memset(dst, 1, total * 8);
  for (uint64_t i = 0; i < 4 * chunk; i++) stage[i] = (uint32_t) i;
  const unsigned hw = std::thread::hardware_concurrency();
  printf("host threads available: %u\n", hw);
  const char *names[4] = {"best", "sse2", "avx2", "avx512"};
  for (int which = 1; which <= 3; which++) {
    if (which == 3 && !__builtin_cpu_supports("avx512f")) continue;
    if (which == 2 && !__builtin_cpu_supports("avx2")) continue;
    for (unsigned T : {1u, 2u, 4u, 8u, 14u, 16u, 24u, 30u, 32u, 48u, 64u}) {
      if (T > hw) break;
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (unsigned t = 0; t < T; t++) th.emplace_back([=] {
        const uint64_t nchunks = total / chunk;
        for (uint64_t k = 0; k < nchunks; k++) {              // every chunk is split over the threads, as in the library
          const uint64_t lo = chunk * t / T, hi = chunk * (t + 1) / T;
          gtb_widen_u32_u64(stage + (k % 4) * chunk + lo, dst + k * chunk + lo, hi - lo, which);
        }
      });
      for (auto &x : th) x.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("%-6s T=%2u  %.1f GB/s written (%.2f Gentries/s)\n", names[which], T, total * 8 / s / 1e9, total / s / 1e9);
    }
  }
  return 0;
}
