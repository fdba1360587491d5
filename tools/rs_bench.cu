// rs_bench.cu -- development tool (not part of the product): times and verifies the
// onesweep radix pass of genometools_b200/csrc/gtb_radix.cuh for several kernel shapes.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/rs_bench tools/rs_bench.cu
//   tools/rs_bench [N] [reps]
//
// Every shape sorts N (key64,val32) pairs with 8 passes; the result is checked on the host
// (sorted by (key, original index) <=> stable; keys match their origin), then timed.
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include <string>
#include <algorithm>

#ifndef RSB_NOPROFILE
#define GTB_RS_PROFILE 1
#endif
#include "../genometools_b200/csrc/gtb_common.cuh"
#include "../genometools_b200/csrc/gtb_radix.cuh"

namespace gtb {
void ErrBuf::set(const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt); vsnprintf(msg, sizeof msg, fmt, ap); va_end(ap);
}
}
using namespace gtb;

static inline u64 splitmix(u64 &s)
{
  u64 z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

static void gen(std::vector<u64> &k, int mode, u64 seed)
{
  u64 s = seed;
  const u64 n = k.size();
  for (u64 i = 0; i < n; i++) {
    const u64 r = splitmix(s);
    switch (mode) {
      case 0: k[i] = r; break;                                            // uniform
      case 1: k[i] = ((i / 37) << 32) | (r & 0x3ffffffull); break;        // doubling-like (group, rank)
      case 2: k[i] = r & 0x0303030303030303ull; break;                    // 4 values per digit (many ties)
      default: k[i] = 0x1234567800000000ull | (r & 0xff); break;          // all digits but one constant
    }
  }
}

struct Result { std::string name; double ms_pass; double gbs; bool ok; int passes; };

template <class Cfg>
static Result run_cfg(const char *name, const std::vector<u64> &hk, int reps, bool verify, int mode)
{
  Result R{name, 0, 0, true, 0};
  ErrBuf err;
  const u64 N = hk.size();
  cudaStream_t st; cudaStreamCreate(&st);
  RadixWork rw;
  if (radix_work_init(rw, err)) { printf("init: %s\n", err.msg); R.ok = false; return R; }
  u64 *kin, *kb[2]; u32 *vin, *vb[2];
  cudaMalloc(&kin, 8 * N); cudaMalloc(&vin, 4 * N);
  for (int i = 0; i < 2; i++) { cudaMalloc(&kb[i], 8 * N); cudaMalloc(&vb[i], 4 * N); }
  std::vector<u32> hv(N);
  for (u64 i = 0; i < N; i++) hv[i] = (u32) i;
  cudaMemcpy(kin, hk.data(), 8 * N, cudaMemcpyHostToDevice);
  cudaMemcpy(vin, hv.data(), 4 * N, cudaMemcpyHostToDevice);
  PassPlan plan; plan.npass = 0; plan_add_bits(plan, 0, 64);
  double best = 1e30;
  int res = 0; u64 nout = 0;
  for (int r = 0; r < reps + 1; r++) {
    rw.ms_radix = 0; rw.passes = 0;
    PairSrc ps{kin, vin};
    if (radix_sort<PairSrc, PairSrc, Cfg>(rw, st, ps, N, kb, vb, plan, &res, &nout, err)) {
      printf("%s: %s\n", name, err.msg); R.ok = false; break;
    }
    if (r > 0 && rw.passes) best = std::min(best, (double) rw.ms_radix / rw.passes);
    R.passes = rw.passes;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); R.ok = false; }
  if (R.ok && verify) {
    std::vector<u64> ok(N); std::vector<u32> ov(N);
    cudaMemcpy(ok.data(), kb[res], 8 * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(ov.data(), vb[res], 4 * N, cudaMemcpyDeviceToHost);
    u64 bad = 0;
    for (u64 j = 0; j < N && bad < 5; j++) {
      if (ov[j] >= N || hk[ov[j]] != ok[j]) { bad++; printf("  %s mode %d: key/val mismatch at %llu\n", name, mode, (unsigned long long) j); continue; }
      if (j > 0 && !(ok[j - 1] < ok[j] || (ok[j - 1] == ok[j] && ov[j - 1] < ov[j]))) {
        bad++; printf("  %s mode %d: order violated at %llu\n", name, mode, (unsigned long long) j);
      }
    }
    if (nout != N) { bad++; printf("  %s: nout %llu != N\n", name, (unsigned long long) nout); }
    R.ok = bad == 0;
  }
  R.ms_pass = best; R.gbs = 24.0 * N / (best * 1e-3) / 1e9;
#ifdef GTB_RS_PROFILE
  {
    unsigned long long ph[8];
    cudaMemcpyFromSymbol(ph, g_rs_phase, sizeof ph);
    const double tiles = (double) ((N + Cfg::TILE - 1) / Cfg::TILE) * (reps + 1) * std::max(1, R.passes);
    printf("    cycles/tile: load+rank %.0f  scan+publish %.0f  stage %.0f  lookback %.0f  scatter %.0f | look-back hops %.2f/tile, spin reloads %.2f/tile\n",
           ph[0] / tiles, ph[1] / tiles, ph[2] / tiles, ph[3] / tiles, ph[4] / tiles, ph[5] / tiles, ph[6] / tiles);
    memset(ph, 0, sizeof ph);
    cudaMemcpyToSymbol(g_rs_phase, ph, sizeof ph);
  }
#endif
  cudaFree(kin); cudaFree(vin);
  for (int i = 0; i < 2; i++) { cudaFree(kb[i]); cudaFree(vb[i]); }
  radix_work_free(rw);
  cudaStreamDestroy(st);
  return R;
}

static const char *g_only = nullptr;
#define RUN(...) do { \
    if (g_only && !strstr(#__VA_ARGS__, g_only)) break; \
    Result r = run_cfg<RsCfg<__VA_ARGS__>>(#__VA_ARGS__, hk, reps, verify, mode); \
    printf("mode %d  RsCfg<%-34s>  smem %6zu  passes %d  %8.4f ms/pass  %8.1f GB/s  %s\n", mode, r.name.c_str(), \
           RsCfg<__VA_ARGS__>::SMEM, r.passes, r.ms_pass, r.gbs, r.ok ? "ok" : "FAILED"); fflush(stdout); \
    if (!r.ok) failures++; } while (0)

int main(int argc, char **argv)
{
  const u64 N = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  const int maxmode = argc > 3 ? atoi(argv[3]) : 3;
  int failures = 0;
  g_only = getenv("RSB_ONLY");
  std::vector<u64> hk(N);
  for (int mode = 0; mode <= maxmode; mode++) {
    gen(hk, mode, 42 + mode);
    const bool verify = true;
    if (mode == 0) {
      //  NT IPT MINB VAL_EARLY LB
      RUN(256, 16, 3, false, 0);     // no look-back: wrong result, upper bound of the rest
      RUN(256, 16, 3, false, 4);
      RUN(256, 16, 4, false, 4);
      RUN(384, 16, 2, false, 4);
      RUN(384, 16, 2, false, 4, 296);    // + L2 prefetch of the tile one wave of CTAs ahead
      RUN(384, 16, 2, false, 4, 148);
      RUN(384, 16, 2, false, 4, 592);
      RUN(384, 16, 2, false, 6, 296);
      RUN(512, 12, 2, false, 4);
      RUN(512, 16, 2, false, 4);
      RUN(512, 16, 2, false, 2);
      RUN(512, 16, 2, false, 8);
      RUN(1024, 8, 1, false, 4);
      RUN(1024, 12, 1, false, 4);
    } else {
      RUN(256, 16, 3, false, 4);
      RUN(512, 16, 2, false, 4);
    }
  }
  printf("failures: %d\n", failures);
  return failures ? 1 : 0;
}
