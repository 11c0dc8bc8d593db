// context.cuh -- solver context: device, stream, optional NCCL communicator,
// error reporting, device buffers.  NCCL is dlopen()ed so that the single-GPU
// path has no link-time dependency on it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200sfm.h"
#include "common.cuh"
#include "p2p_allreduce.cuh"
#include "pcg.cuh"

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;   // optional: error path only
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    // Prefer an already-loaded libnccl (e.g. the one bundled with torch).
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    CommAbort = (decltype(CommAbort))dlsym(lib, "ncclCommAbort");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
      err = "libnccl is missing a required symbol";
      return false;
    }
    return true;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

struct NcclError {
  std::string msg;
};

struct b200sfm_ctx {
  int device = 0;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  std::string err;
  long long launches = 0;
  double* h_scal = nullptr;   // pinned readback area
  static constexpr int kHScal = 4096;
  // PCG loop state (pinned control slots, events, device scratch): owned by the context, not by a problem, so that a
  // one-shot solve does not pay cudaMallocHost / cudaMalloc / cudaFree (a device-wide synchronisation) on every call
  b200::PcgHost pcgh;

  // small sums go over peer memory when it is set up (p2p_allreduce.cuh), everything else through NCCL
  b200::P2PAllReduce p2p;
  long long p2p_calls = 0;
  void allreduce_sum(double* buf, size_t n) {
    if (world > 1 && p2p.ready && n > 0 && n <= p2p.cap && (reinterpret_cast<uintptr_t>(buf) & 15) == 0) {   // double2 accesses
      p2p.launch(stream, buf, n);
      ++p2p_calls;
      ++launches;
      return;
    }
    allreduce_sum_on(stream, buf, n);
  }
  void allreduce_sum_on(cudaStream_t st, double* buf, size_t n) {
    if (world == 1 || n == 0) return;
    if (!comm) throw NcclError{"communicator was aborted after an earlier failure"};
    ncclResult_t r = nccl_api().AllReduce(buf, buf, n, ncclFloat64, ncclSum, comm, st);
    if (r != ncclSuccess) throw NcclError{std::string("ncclAllReduce(sum): ") + nccl_api().GetErrorString(r)};
  }
  // Second stream for the first half of a split all-reduce (the BA mat-vec reduces the cameras below C/2 while pass B
  // still runs over the segments of the upper half): created on first use, destroyed with the context.  Every rank
  // issues its collectives in the same order (comm stream first, then the main stream), as NCCL requires.
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_half = nullptr, ev_comm = nullptr;
  void ensure_comm_stream() {
    if (comm_stream) return;
    B200_CUDA_OK(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
    B200_CUDA_OK(cudaEventCreateWithFlags(&ev_half, cudaEventDisableTiming));
    B200_CUDA_OK(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
  }
  void allreduce_max(double* buf, size_t n) {
    if (world == 1 || n == 0) return;
    if (!comm) throw NcclError{"communicator was aborted after an earlier failure"};
    ncclResult_t r = nccl_api().AllReduce(buf, buf, n, ncclFloat64, ncclMax, comm, stream);
    if (r != ncclSuccess) throw NcclError{std::string("ncclAllReduce(max): ") + nccl_api().GetErrorString(r)};
  }
};

namespace b200 {

// Device buffers come from the stream-ordered allocator (cudaMallocAsync on the context's stream, default
// memory pool with an unbounded release threshold): a one-shot solve allocates a few GB in ~20 buffers, and
// with the pool a repeated call re-uses them instead of paying cudaMalloc/cudaFree every time.  Every ABI
// entry point installs its context's stream for the calling thread (AllocScope); without one -- or with
// B200SFM_ASYNC_ALLOC=0 -- plain cudaMalloc/cudaFree are used.
inline cudaStream_t& alloc_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
struct AllocScope {
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t s) : prev(alloc_stream()) { alloc_stream() = s; }
  ~AllocScope() { alloc_stream() = prev; }
};
inline bool async_alloc_enabled() {
  static const bool on = !(getenv("B200SFM_ASYNC_ALLOC") && atoi(getenv("B200SFM_ASYNC_ALLOC")) == 0);
  return on;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (!count) return;
    if (alloc_stream() && async_alloc_enabled()) B200_CUDA_OK(cudaMallocAsync(&p, count * sizeof(T), alloc_stream()));
    else B200_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
  }
  void release() {
    if (p) {
      // stream-ordered free when a context stream is installed (cudaFree is valid for both kinds otherwise)
      if (alloc_stream() && async_alloc_enabled()) cudaFreeAsync(p, alloc_stream());
      else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
  void zero(cudaStream_t s) {
    if (n) B200_CUDA_OK(cudaMemsetAsync(p, 0, bytes(), s));
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count) B200_CUDA_OK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) B200_CUDA_OK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

// Persisting-L2 window over an array that one pass writes and the next gathers (design v2: z4).  The
// set-aside is bounded by the device limits; every call is best-effort (a refusal only costs bandwidth).
inline void l2_persist_window(cudaStream_t s, int device, void* base, size_t bytes) {
  int max_persist = 0, max_window = 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
  if (max_persist <= 0 || max_window <= 0 || bytes == 0) { cudaGetLastError(); return; }
  const size_t setaside = std::min((size_t)max_persist, bytes);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside);
  cudaStreamAttrValue a{};
  a.accessPolicyWindow.base_ptr = base;
  a.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
  a.accessPolicyWindow.hitRatio = std::min(1.0f, (float)setaside / (float)a.accessPolicyWindow.num_bytes);
  a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a);
  cudaGetLastError();
}
inline void l2_persist_clear(cudaStream_t s) {
  cudaStreamAttrValue a{};
  a.accessPolicyWindow.num_bytes = 0;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a);
  cudaCtxResetPersistingL2Cache();
  cudaGetLastError();
}

struct EventTimer {
  // pool of event pairs timing selected kernels; summed after a final sync
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  ~EventTimer() {
    for (auto e : ev) cudaEventDestroy(e);
  }
  cudaEvent_t next() {
    if (used == ev.size()) {
      cudaEvent_t e;
      B200_CUDA_OK(cudaEventCreate(&e));
      ev.push_back(e);
    }
    return ev[used++];
  }
  void reset() { used = 0; }
};

#define B200_LAUNCH(ctx, kernel, grid, block, smem, ...)                       \
  do {                                                                         \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);           \
    ++(ctx)->launches;                                                         \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace b200
