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

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
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

  void allreduce_sum(double* buf, size_t n) {
    if (world == 1 || n == 0) return;
    ncclResult_t r = nccl_api().AllReduce(buf, buf, n, ncclFloat64, ncclSum, comm, stream);
    if (r != ncclSuccess) throw NcclError{std::string("ncclAllReduce(sum): ") + nccl_api().GetErrorString(r)};
  }
  void allreduce_max(double* buf, size_t n) {
    if (world == 1 || n == 0) return;
    ncclResult_t r = nccl_api().AllReduce(buf, buf, n, ncclFloat64, ncclMax, comm, stream);
    if (r != ncclSuccess) throw NcclError{std::string("ncclAllReduce(max): ") + nccl_api().GetErrorString(r)};
  }
};

namespace b200 {

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
    if (count) B200_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
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
