// gtb_shard.cuh -- one enhanced-suffix-array job sharded over several code ranges that can
// reach each other's HBM: the GPUs of one box (north_star: "contiguous .bck code ranges balanced
// by bucket counts are assigned to the 8 GPUs of one box, each holding the full replicated
// encseq").  Included by gtb_esa.cu (one translation unit).
//
// Every range runs the SAME function (run_sharded) -- as a thread of one process (gtb_group, the
// C host's `gt -j N`) or as a process of its own (bench.py under torchrun) -- and meets the others
// at a few all-gathers of small host structs (the only collective the caller has to provide).
// Everything heavy goes through peer memory instead of a collective library:
//
//   * text scan sharded by position: rank r turns the positions of its 1/world slice into filled
//     keys and partitions them by owning range with ONE onesweep pass whose per-bin output
//     pointers are the owners' receive buffers -- the partition kernel stores the positions
//     straight into the owner's HBM over NVLink (rs_owner_scatter with binbase), in text order;
//   * prefix doubling: rank(p + h) of a suffix another range owns is READ from that range's
//     rank map (rank words, tied ranks, sorted keys + bucket table) by the kernel that builds
//     the sort keys (k_build_dkeys_peer) -- no request/answer exchange, two sync points a round;
//   * the bucket table of a job is the sum of the ranges' tables (k_add_u32 over peer pointers).
//
// Mirrors gt_suftabparts_new + the parts loop of the reference (src/match/sfx-partssuf.c:172-347,
// src/match/sfx-suffixer.c:1791-1838), whose parts never need each other because its sorters
// compare text; seams as computelocallcpvalue (src/match/sfx-lcpvalues.c:91-111).
#pragma once
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <mutex>
#include <condition_variable>

namespace gtb {

// ---- a device pointer another range can map ------------------------------------------------
struct PeerPtr {
  u64 ptr;                        // address in the exporting process (allocation base)
  int device, valid;
  int slot, pad;                  // which shareable buffer of the exporter (separate processes:
  u64 alloc_id, size;             // the importer maps allocation alloc_id of `size` bytes, gtb_vmm.cuh)
};

struct RankView {                 // the rank map of one range (RankMap, gtb_esa_kernels.cuh)
  PeerPtr keys, sa, rw, trank, lb;
  u64 N, sa_offset, own_last;
  int has_map, pad;
};

// device side: the rank maps of all ranges as this GPU addresses them
struct PeerMapDev {
  const u64 *keys; const u32 *sa; const uint4 *rw; const u32 *trank; const u32 *lb;
  u64 N, sa_offset, own_last;
};
struct PeerTableDev {
  int n, mine;
  u64 first_key[MAX_RANGES];
  PeerMapDev m[MAX_RANGES];
};

// key = (group head, rank of the suffix h further); the rank of a position another range owns is
// read from the owner's rank map in ITS memory.  Special positions are ranked from this range's own
// rank words (the special bits are the same in every range).
// which range owns the suffix with filled key kq
__device__ __forceinline__ int owner_of_key(const u64 *s_first, int nr, u64 kq)
{
  int lo = 0, hi = nr - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_first[mid] <= kq) lo = mid; else hi = mid - 1; }
  return lo;
}

// Two kernels, as k_build_dkeys: the first answers the partners that are special or tied (one load
// from the own rank words, for a foreign partner two loads from the owner's memory), the second
// searches the queued, untied partners in the owner's sorted keys with all lanes of a warp searching.
template <bool DNA>
__global__ void k_build_dkeys_peer(RankMap<DNA> rm, const PeerTableDev *__restrict__ pt,
                                   const u32 *__restrict__ upos, const u32 *__restrict__ ugrp,
                                   u64 M, u64 h, u64 *__restrict__ dkeys, u32 *__restrict__ queue, unsigned int *qcount,
                                   int pairs_by_text)
{
  __shared__ u64 s_first[MAX_RANGES];
  const int nr = pt->n, mine = pt->mine;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) s_first[i] = pt->first_key[i];
  __syncthreads();
  const u64 stride = (u64) gridDim.x * blockDim.x;
  for (u64 c0 = blockIdx.x * (u64) blockDim.x; c0 < M; c0 += stride) {      // (whole warps stay in the loop)
    const u64 c = c0 + threadIdx.x;
    bool need = false;
    if (c < M) {
      const u64 q = (u64) upos[c] + h;
      u32 r = 0;
      bool have = true;
      u64 partner;
      int less = -1;
      if (DNA && pairs_by_text && tie_pair_partner(ugrp, M, c, &partner))      // (a tie group lies inside one range)
        less = dna_pair_less(rm.src.words, rm.src.spmask, (u64) upos[c], (u64) upos[partner], h);
      if (less >= 0) r = less ? 0u : 1u;
      else if (q >= rm.n) r = (u32) rm.n;
      else {
        const uint4 w = __ldg(rm.rw + (q >> 5));
        const u32 bit = 1u << (q & 31u), below = bit - 1u;
        if (w.x & bit) r = (u32) (rm.nonspecials + w.y + (u32) __popc(w.x & below));
        else {
          u64 kq;
          rm.src.make_key_fmt(q, kq, rm.src.f);
          const int o = owner_of_key(s_first, nr, kq);
          if (o == mine) {
            if (w.z & bit) r = rm.trank[w.w + (u32) __popc(w.z & below)]; else have = false;
          } else {
            const PeerMapDev &m = pt->m[o];
            const uint4 wo = m.rw[q >> 5];
            if (wo.z & bit) r = m.trank[wo.w + (u32) __popc(wo.z & below)]; else have = false;
          }
        }
      }
      if (have) dkeys[c] = ((u64) ugrp[c] << 32) | (u64) r;
      need = !have;
    }
    queue_append(need, (u32) c, queue, qcount);
  }
}

template <bool DNA>
__global__ void k_build_dkeys_peer_search(RankMap<DNA> rm, const PeerTableDev *__restrict__ pt,
                                          const u32 *__restrict__ upos, const u32 *__restrict__ ugrp, u64 h,
                                          u64 *__restrict__ dkeys, const u32 *__restrict__ queue,
                                          const unsigned int *__restrict__ qcount)
{
  __shared__ u64 s_first[MAX_RANGES];
  const int nr = pt->n, mine = pt->mine;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) s_first[i] = pt->first_key[i];
  __syncthreads();
  const u64 nq = *qcount;
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < nq; i += (u64) gridDim.x * blockDim.x) {
    const u32 c = queue[i];
    const u64 q = (u64) upos[c] + h;
    u64 kq;
    rm.src.make_key_fmt(q, kq, rm.src.f);
    const int o = owner_of_key(s_first, nr, kq);
    u32 r;
    if (o == mine) r = rm.search(q);
    else {
      const PeerMapDev &m = pt->m[o];
      const u64 code = key_code<DNA>(kq, rm.pl, rm.K, rm.src.f);
      u64 a = (u64) m.lb[code] - m.sa_offset;
      u64 b = code == m.own_last ? m.N : (u64) m.lb[code + 1] - m.sa_offset;
      while (a < b) {
        const u64 mid = (a + b) >> 1;
        const u64 km = m.keys[mid];
        const bool less = km < kq || (km == kq && (u64) m.sa[mid] < q);
        if (less) a = mid + 1; else b = mid;
      }
      r = (u32) (m.sa_offset + a);
    }
    dkeys[c] = ((u64) ugrp[c] << 32) | (u64) r;
  }
}

__global__ void k_add_u32(u32 *__restrict__ dst, const u32 *__restrict__ src, u64 count)
{
  for (u64 i = blockIdx.x * (u64) blockDim.x + threadIdx.x; i < count; i += (u64) gridDim.x * blockDim.x) {
    const u32 v = src[i];
    if (v) dst[i] += v;
  }
}

} // namespace gtb
