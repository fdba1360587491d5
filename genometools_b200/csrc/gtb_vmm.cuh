// gtb_vmm.cuh -- device buffers that another PROCESS can map with full-size pages.
//
// A job sharded over one process per GPU (bench.py under torchrun) reads the other ranges' rank
// maps in place (gtb_shard.cuh).  Memory mapped with cudaIpcOpenMemHandle is unusable for that:
// measured on 2 B200 (c4, round 0 of the doubling: 19.6 M rank lookups, half of them in the other
// GPU's memory) the lookup kernel takes 66.5 ms through a cudaIpc mapping against 3.5 ms through
// cudaDeviceEnablePeerAccess inside one process -- every random access misses the TLB.  So the
// buffers that peers read are allocated with the virtual-memory-management API (cuMemCreate, 2 MB
// granularity), exported as POSIX file descriptors, passed over an abstract unix datagram socket
// (SCM_RIGHTS) and mapped by the importer with cuMemMap at the same granularity.
// The driver entry points are taken from the runtime (cudaGetDriverEntryPoint): the library does
// not link against libcuda.
#pragma once
#include <cuda.h>
#include <atomic>
#include <errno.h>
#include <poll.h>
#include <stddef.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include "gtb_common.cuh"

namespace gtb {

struct VmmApi {
  CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*exportHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*importHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  bool ok = false;
};

static inline const VmmApi &vmm_api()
{
  static VmmApi api = [] {
    VmmApi a;
    bool ok = true;
    auto get = [&](const char *name, void **fp) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &q) != cudaSuccess || *fp == nullptr) { cudaGetLastError(); ok = false; }
    };
    get("cuMemGetAllocationGranularity", (void **) &a.getGranularity);
    get("cuMemCreate", (void **) &a.create);
    get("cuMemRelease", (void **) &a.release);
    get("cuMemAddressReserve", (void **) &a.addressReserve);
    get("cuMemAddressFree", (void **) &a.addressFree);
    get("cuMemMap", (void **) &a.map);
    get("cuMemUnmap", (void **) &a.unmap);
    get("cuMemSetAccess", (void **) &a.setAccess);
    get("cuMemExportToShareableHandle", (void **) &a.exportHandle);
    get("cuMemImportFromShareableHandle", (void **) &a.importHandle);
    a.ok = ok;
    return a;
  }();
  return api;
}

static inline CUmemAllocationProp vmm_prop(int device)
{
  CUmemAllocationProp p;
  memset(&p, 0, sizeof p);
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// physical memory on `device`, mapped read/write for `device`; *size_out = bytes rounded to the granularity
static inline int vmm_alloc(int device, size_t bytes, void **ptr, CUmemGenericAllocationHandle *mh, size_t *size_out, ErrBuf &err)
{
  const VmmApi &a = vmm_api();
  if (!a.ok) { err.set("the CUDA driver lacks the virtual memory management API (cuMemCreate ...)"); return -1; }
  const CUmemAllocationProp prop = vmm_prop(device);
  size_t gran = 0;
  if (a.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = size_t(2) << 20;
  const size_t size = (bytes + gran - 1) / gran * gran;
  CUresult r = a.create(mh, size, &prop, 0);
  if (r != CUDA_SUCCESS) { err.set("cuMemCreate of %zu bytes on device %d failed (CUresult %d)", size, device, (int) r); return -1; }
  CUdeviceptr dp = 0;
  r = a.addressReserve(&dp, size, gran, 0, 0);
  if (r == CUDA_SUCCESS) {
    r = a.map(dp, size, 0, *mh, 0);
    if (r == CUDA_SUCCESS) {
      CUmemAccessDesc ad;
      memset(&ad, 0, sizeof ad);
      ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = device; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      r = a.setAccess(dp, size, &ad, 1);
      if (r != CUDA_SUCCESS) a.unmap(dp, size);
    }
    if (r != CUDA_SUCCESS) a.addressFree(dp, size);
  }
  if (r != CUDA_SUCCESS) { a.release(*mh); err.set("mapping %zu bytes of shareable device memory failed (CUresult %d)", size, (int) r); return -1; }
  *ptr = (void *) dp; *size_out = size;
  return 0;
}

static inline void vmm_free(void *ptr, CUmemGenericAllocationHandle mh, size_t size)
{
  const VmmApi &a = vmm_api();
  if (!a.ok || !ptr) return;
  a.unmap((CUdeviceptr) ptr, size);
  a.addressFree((CUdeviceptr) ptr, size);
  a.release(mh);
}

// map an allocation of another process (received as a file descriptor) for `device`
static inline int vmm_import(int device, int fd, size_t size, void **ptr, CUmemGenericAllocationHandle *mh, ErrBuf &err)
{
  const VmmApi &a = vmm_api();
  if (!a.ok) { err.set("the CUDA driver lacks the virtual memory management API"); return -1; }
  CUresult r = a.importHandle(mh, (void *) (uintptr_t) fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) { err.set("cuMemImportFromShareableHandle failed (CUresult %d)", (int) r); return -1; }
  const CUmemAllocationProp prop = vmm_prop(device);
  size_t gran = 0;
  if (a.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = size_t(2) << 20;
  CUdeviceptr dp = 0;
  r = a.addressReserve(&dp, size, gran, 0, 0);
  if (r == CUDA_SUCCESS) {
    r = a.map(dp, size, 0, *mh, 0);
    if (r == CUDA_SUCCESS) {
      CUmemAccessDesc ad;
      memset(&ad, 0, sizeof ad);
      ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = device; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      r = a.setAccess(dp, size, &ad, 1);
      if (r != CUDA_SUCCESS) a.unmap(dp, size);
    }
    if (r != CUDA_SUCCESS) a.addressFree(dp, size);
  }
  if (r != CUDA_SUCCESS) { a.release(*mh); err.set("mapping %zu bytes of a peer's memory on device %d failed (CUresult %d)", size, device, (int) r); return -1; }
  *ptr = (void *) dp;
  return 0;
}

static inline int vmm_export_fd(CUmemGenericAllocationHandle mh, int *fd, ErrBuf &err)
{
  const VmmApi &a = vmm_api();
  int f = -1;
  CUresult r = a.exportHandle(&f, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS || f < 0) { err.set("cuMemExportToShareableHandle failed (CUresult %d)", (int) r); return -1; }
  *fd = f;
  return 0;
}

// ---- file descriptors between the processes of one job: abstract unix datagram sockets ----
struct FdMsg { int from, slot; unsigned long long alloc_id, size; };

static inline socklen_t fd_sock_addr(struct sockaddr_un *sa, const char *key, int rank)
{
  memset(sa, 0, sizeof *sa);
  sa->sun_family = AF_UNIX;
  const int len = snprintf(sa->sun_path + 1, sizeof sa->sun_path - 1, "gtb200-%s-%d", key, rank);   // leading NUL: abstract
  return (socklen_t) (offsetof(struct sockaddr_un, sun_path) + 1 + (size_t) len);
}

static inline int fd_sock_open(const char *key, int rank, ErrBuf &err)
{
  const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) { err.set("socket() failed: %s", strerror(errno)); return -1; }
  struct sockaddr_un sa;
  const socklen_t len = fd_sock_addr(&sa, key, rank);
  if (bind(s, (struct sockaddr *) &sa, len) != 0) { err.set("bind of the descriptor socket of rank %d failed: %s", rank, strerror(errno)); close(s); return -1; }
  struct timeval tv; tv.tv_sec = 60; tv.tv_usec = 0;
  setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  return s;
}

static inline int fd_send(int sock, const char *key, int to_rank, const FdMsg &m, int fd, ErrBuf &err)
{
  struct sockaddr_un sa;
  const socklen_t len = fd_sock_addr(&sa, key, to_rank);
  struct msghdr mh;
  memset(&mh, 0, sizeof mh);
  struct iovec iov; iov.iov_base = const_cast<FdMsg *>(&m); iov.iov_len = sizeof m;
  alignas(struct cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof ctl);
  mh.msg_name = &sa; mh.msg_namelen = len; mh.msg_iov = &iov; mh.msg_iovlen = 1;
  mh.msg_control = ctl; mh.msg_controllen = sizeof ctl;
  struct cmsghdr *cm = CMSG_FIRSTHDR(&mh);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  // non-blocking: returns 1 when the peer's queue is full (or the peer has not bound its socket yet) --
  // the caller empties its own socket and tries again, so that ranks sending to each other never wait
  // on each other
  if (sendmsg(sock, &mh, MSG_DONTWAIT) == (ssize_t) sizeof m) return 0;
  if (errno == ECONNREFUSED || errno == ENOENT || errno == EAGAIN || errno == EWOULDBLOCK || errno == ENOBUFS) return 1;
  err.set("sending a memory descriptor to rank %d failed: %s", to_rank, strerror(errno));
  return -1;
}

static inline bool fd_sock_readable(int sock)
{
  struct pollfd p; p.fd = sock; p.events = POLLIN; p.revents = 0;
  return poll(&p, 1, 0) == 1 && (p.revents & POLLIN);
}

static inline int fd_recv(int sock, FdMsg *m, int *fd, ErrBuf &err)
{
  struct msghdr mh;
  memset(&mh, 0, sizeof mh);
  struct iovec iov; iov.iov_base = m; iov.iov_len = sizeof *m;
  alignas(struct cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  mh.msg_iov = &iov; mh.msg_iovlen = 1; mh.msg_control = ctl; mh.msg_controllen = sizeof ctl;
  const ssize_t r = recvmsg(sock, &mh, MSG_CMSG_CLOEXEC);
  if (r != (ssize_t) sizeof *m) { err.set("receiving a memory descriptor failed: %s", r < 0 ? strerror(errno) : "short message"); return -1; }
  struct cmsghdr *cm = CMSG_FIRSTHDR(&mh);
  if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) { err.set("descriptor message without a descriptor"); return -1; }
  memcpy(fd, CMSG_DATA(cm), sizeof(int));
  return 0;
}

} // namespace gtb
