// p2p_allreduce.cuh -- one-shot sum all-reduce over NVLink peer memory for the small, latency-bound vectors of the
// multi-GPU solvers (the 6 C doubles of every PCG mat-vec: 480 KB at config 4).
//
// EXPERIMENT, opt-in (B200SFM_P2P_AR=1): it is correct (2- and 8-rank parity of the BA solve) but SLOWER than NCCL's
// all-reduce of this size on both 2 GPUs (28.2 vs 26.9 ms per step) and 8 GPUs (15.1 vs 13.75 ms), profiles/r2_scaling.md.
// Every rank publishes its vector in a buffer the peers have mapped (CUDA IPC), raises one flag per peer, and then reads
// all N published vectors straight over NVLink / NVSwitch and adds them IN RANK ORDER: one hop, and the result is
// bitwise identical on every rank (the replicated PCG control flow relies on that).  What it pays for that hop -- a
// system-scope fence after the publication, N flag stores and an acquire spin per call -- costs more than NCCL's
// protocol does (the measured build still issued one release store, i.e. one system fence, PER PEER); the next thing to try is publishing straight from the producing kernel (pass B) instead of a copy.
//
//   kernel, per rank:   publish my vector  ->  last CTA: release-store epoch into flag[me] of every peer
//                       every CTA: acquire-spin until my flag[r] >= epoch for all r  ->  sum_r buf_r[i] (r = 0 .. N-1)
//   two slots alternate by epoch: a rank can only be one all-reduce ahead of the slowest one (it needs everybody's flag
//   of the previous epoch to get past it), so slot (e & 1) is never overwritten while somebody still reads it.
// The grid is at most one CTA per SM (co-resident: CTAs spin on flags), the collective is issued on the context's stream
// like an NCCL call and, like NCCL, deadlocks if a rank does not issue it -- the callers are SPMD.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace b200 {

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kP2PThreads = 512;
constexpr int kP2PMaxWorld = 16;
constexpr unsigned long long kP2PTimeoutNs = 20ull * 1000 * 1000 * 1000;   // 20 s: far above any skew between ranks

struct P2PPeers {
  double* buf[kP2PMaxWorld];                 // published vectors of every rank (own included), [2][cap]
  unsigned long long* flags[kP2PMaxWorld];   // flag array of every rank: flags[r][s] = last epoch rank s has published
};

__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_sum(double* __restrict__ data, size_t n, int world, int rank,
                                                                P2PPeers peers, size_t slot_off, unsigned long long epoch,
                                                                unsigned* __restrict__ counter, int* __restrict__ err) {
  // a rank that failed elsewhere never raises its flag: give up after kP2PTimeoutNs instead of hanging the GPU; once the
  // error word is set every later call returns at once and the host reports B200SFM_ERR_NCCL at the end of the solve
  if (*reinterpret_cast<volatile int*>(err)) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double* mine = peers.buf[rank] + slot_off;
  {
    const size_t n2p = n / 2;
    for (size_t i = t0; i < n2p; i += stride) reinterpret_cast<double2*>(mine)[i] = reinterpret_cast<const double2*>(data)[i];
    if ((n & 1) && t0 == 0) mine[n - 1] = data[n - 1];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {   // the last CTA of this rank: the whole vector is published
      *counter = 0;                                  // (every CTA has arrived; the next launch starts from 0)
      __threadfence_system();                        // ONE system fence, then plain (relaxed) flag stores: a release store
      for (int r = 0; r < world; ++r)                // per peer compiles to a MEMBAR.ALL.SYS each (profiles/r2_sass_excerpt.txt)
        st_relaxed_sys_u64(peers.flags[r] + rank, epoch);
    }
  }
  if (threadIdx.x < world) {
    const unsigned long long* f = peers.flags[rank] + threadIdx.x;
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys_u64(f) < epoch) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > kP2PTimeoutNs) { *reinterpret_cast<volatile int*>(err) = 1; break; }
    }
  }
  __syncthreads();
  // (the acquire above orders these reads after the peers' publication; peer memory is not cached in this GPU's L2 and
  //  L1 is bypassed with .cv, so plain vector loads are enough -- every load of a thread is in flight at once)
  const size_t n2 = n / 2;
  for (size_t i = t0; i < n2; i += stride) {
    double2 s = make_double2(0.0, 0.0);
#pragma unroll 8
    for (int r = 0; r < world; ++r) {   // fixed order: bitwise identical on all ranks
      const double2 v = __ldcv(reinterpret_cast<const double2*>(peers.buf[r] + slot_off) + i);
      s.x += v.x;
      s.y += v.y;
    }
    reinterpret_cast<double2*>(data)[i] = s;
  }
  if ((n & 1) && t0 == 0) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(peers.buf[r] + slot_off + n - 1);
    data[n - 1] = s;
  }
}

// Host side: owned by the context of a multi-rank run.  setup() is collective (all ranks, same order).
struct P2PAllReduce {
  bool ready = false;
  int world = 1, rank = 0;
  size_t cap = 0;                      // doubles per slot
  void* own = nullptr;                 // [2][cap] doubles | flags[kP2PMaxWorld]
  void* opened[kP2PMaxWorld] = {};
  unsigned* counter = nullptr;   // [0]: CTA arrival counter, [1]: error word (timeout)
  P2PPeers peers{};
  unsigned long long epoch = 0;

  static size_t bytes_for(size_t cap_doubles) { return 2 * cap_doubles * sizeof(double) + kP2PMaxWorld * sizeof(unsigned long long); }

  // gather: collective exchange of `bytes` bytes per rank (implemented by the caller over NCCL); returns false on error
  template <class Gather>
  bool setup(int device, int rank_, int world_, size_t cap_doubles, Gather&& gather) {
    world = world_; rank = rank_; cap = cap_doubles;
    if (world < 2 || world > kP2PMaxWorld) return false;
    bool ok = true;
    // device of every rank, then peer access from here to each of them
    std::vector<int> devs(world, -1);
    devs[rank] = device;
    if (!gather(devs.data(), sizeof(int))) return false;
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      int can = 0;
      if (devs[r] == device || cudaDeviceCanAccessPeer(&can, device, devs[r]) != cudaSuccess || !can) ok = false;
    }
    cudaIpcMemHandle_t mine{};
    if (ok) {
      ok = cudaMalloc(&own, bytes_for(cap)) == cudaSuccess && cudaMemset(own, 0, bytes_for(cap)) == cudaSuccess &&
           cudaMalloc(&counter, 2 * sizeof(unsigned)) == cudaSuccess && cudaMemset(counter, 0, 2 * sizeof(unsigned)) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine, own) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    }
    std::vector<cudaIpcMemHandle_t> handles(world);
    std::memset(handles.data(), 0, handles.size() * sizeof(cudaIpcMemHandle_t));
    handles[rank] = mine;
    std::vector<int> oks(world, 0);
    oks[rank] = ok ? 1 : 0;
    if (!gather(handles.data(), sizeof(cudaIpcMemHandle_t)) || !gather(oks.data(), sizeof(int))) { release(); return false; }
    for (int r = 0; r < world; ++r) ok = ok && oks[r] == 1;
    if (ok) {
      for (int r = 0; r < world && ok; ++r) {
        void* p = own;
        if (r != rank) {
          ok = cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
          if (ok) opened[r] = p;
        }
        peers.buf[r] = reinterpret_cast<double*>(p);
        peers.flags[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p) + 2 * cap * sizeof(double));
      }
    }
    // every rank must reach the same verdict
    std::fill(oks.begin(), oks.end(), 0);
    oks[rank] = ok ? 1 : 0;
    if (!gather(oks.data(), sizeof(int))) { release(); return false; }
    for (int r = 0; r < world; ++r) ok = ok && oks[r] == 1;
    cudaGetLastError();
    if (!ok) { release(); return false; }
    ready = true;
    return true;
  }
  void launch(cudaStream_t s, double* data, size_t n) {
    ++epoch;
    // one double2 per thread when the vector fits (a single round of NVLink latency), at most 120 co-resident CTAs
    const int grid = (int)std::min<size_t>(std::max<size_t>((n / 2 + kP2PThreads - 1) / kP2PThreads, 1), 120);
    p2p_allreduce_sum<<<grid, kP2PThreads, 0, s>>>(data, n, world, rank, peers, (epoch & 1) * cap, epoch, counter,
                                                    reinterpret_cast<int*>(counter + 1));
    B200_CUDA_OK(cudaGetLastError());
  }
  // true when a wait timed out since setup (synchronises the stream)
  bool timed_out(cudaStream_t s) {
    if (!ready) return false;
    int h = 0;
    if (cudaMemcpyAsync(&h, counter + 1, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) return true;
    if (cudaStreamSynchronize(s) != cudaSuccess) return true;
    return h != 0;
  }
  void release() {
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (opened[r]) { cudaIpcCloseMemHandle(opened[r]); opened[r] = nullptr; }
    if (own) { cudaFree(own); own = nullptr; }
    if (counter) { cudaFree(counter); counter = nullptr; }
    ready = false;
    cudaGetLastError();
  }
};

}  // namespace b200
