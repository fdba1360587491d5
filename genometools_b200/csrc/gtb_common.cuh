// gtb_common.cuh -- shared helpers of libgtb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace gtb {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t  u8;

constexpr unsigned FULL_MASK = 0xffffffffu;

// ----- error plumbing: every host function returns 0 / -1 and leaves a message -----
struct ErrBuf {
  char msg[512];
  ErrBuf() { msg[0] = 0; }
  void set(const char *fmt, ...) __attribute__((format(printf, 2, 3)));
};

#define GTB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      err.set("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,              \
              cudaGetErrorString(e_));                                        \
      return -1;                                                              \
    }                                                                         \
  } while (0)

#define GTB_TRY(expr)                                                         \
  do { if ((expr) != 0) return -1; } while (0)

#define GTB_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t e_ = cudaGetLastError();                                      \
    if (e_ != cudaSuccess) {                                                  \
      err.set("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,          \
              cudaGetErrorString(e_));                                        \
      return -1;                                                              \
    }                                                                         \
  } while (0)

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

// ----- device helpers -----
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p)
{
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Weak load that does not allocate in L1 (SASS LDG.E.NA): served by L2, the point of
// coherence.  Enough for words that are self-validating (written by one 64-bit store,
// tag and payload together).  Unlike ld.relaxed.gpu / ld.global.cg (both LDG.STRONG.GPU
// on sm_100a, which complete one at a time per thread) many of these can be in flight.
__device__ __forceinline__ u64 ld_na_u64(const u64 *p)
{
  u64 v;
  asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_weak_u64(u64 *p, u64 v)
{
  asm volatile("st.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// block-wide exclusive sum; returns the exclusive prefix of `v`, *total = block sum.
// scratch: NT/32 + 1 entries of T.
template <int NT, typename T>
__device__ __forceinline__ T block_exclusive_sum(T v, T *scratch, T *total)
{
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(FULL_MASK, incl, d);
    if (lane >= (unsigned) d) incl += t;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < NT / 32 ? scratch[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T t = __shfl_up_sync(FULL_MASK, wi, d);
      if (lane >= (unsigned) d) wi += t;
    }
    if (lane < NT / 32) scratch[lane] = wi - w;      // exclusive warp offsets
    if (lane == NT / 32 - 1) scratch[NT / 32] = wi;  // total
  }
  __syncthreads();
  T res = scratch[warp] + incl - v;
  *total = scratch[NT / 32];
  __syncthreads();
  return res;
}

// block-wide inclusive max-scan (values are "index+1, 0 = none")
template <int NT, typename T>
__device__ __forceinline__ T block_inclusive_max(T v, T *scratch)
{
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(FULL_MASK, incl, d);
    if (lane >= (unsigned) d) incl = t > incl ? t : incl;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < NT / 32 ? scratch[lane] : T(0);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T t = __shfl_up_sync(FULL_MASK, w, d);
      if (lane >= (unsigned) d) w = t > w ? t : w;
    }
    if (lane < NT / 32) scratch[lane] = w;           // inclusive over warps
  }
  __syncthreads();
  T prev = warp > 0 ? scratch[warp - 1] : T(0);
  T res = prev > incl ? prev : incl;
  __syncthreads();
  return res;
}

} // namespace gtb
