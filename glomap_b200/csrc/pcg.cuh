// pcg.cuh -- block-preconditioned conjugate gradients on the reduced camera
// system, entirely stream-ordered on the device (no host sync inside an
// iteration).  The reference factors this system with CHOLMOD
// (bundle_adjustment.cc:94-96, global_positioning.cc:551-559); north_star
// mandates PCG with one all-reduce per mat-vec.
//
// All inner products are DETERMINISTIC: every CTA writes its partial sum to
// part[which][blockIdx.x] and the consuming kernel re-sums the partials in a
// fixed order.  With replicated camera-sized vectors this makes alpha, beta and
// the convergence test bit-identical on every rank, so all ranks take the same
// control-flow decisions without an extra collective.
//
// Scalars: dots[it][0] = p.q, dots[it][1] = r.z after iteration it,
// dots[it][2] = r.r after iteration it (it = 0: initial values).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kPcgThreads = 128;

// fixed-order sum of n partials by the first warp of the CTA; result broadcast
// through shared memory to all threads.
__device__ __forceinline__ double sum_partials(const double* __restrict__ part, int n, double* sh) {
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *sh = s;
  }
  __syncthreads();
  return *sh;
}
// CTA-level deterministic sum -> part[blockIdx.x]
__device__ __forceinline__ void write_partial(double v, double* __restrict__ part, double* scratch) {
  v = block_sum(v, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// q_c = (A_c + diag(D_c)) p_c + yw_c ; part_pq[blk] = partial p.q
template <int B>
__global__ void __launch_bounds__(kPcgThreads) pcg_apply_diag(int nb, const double* __restrict__ A,
                                                              const double* __restrict__ D,
                                                              const double* __restrict__ p,
                                                              const double* __restrict__ yw, double* __restrict__ q,
                                                              double* __restrict__ part_pq) {
  constexpr int NP = B * (B + 1) / 2;
  __shared__ double scratch[32];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double pq = 0.0;
  if (c < nb) {
    double a[NP], pv[B], out[B];
#pragma unroll
    for (int k = 0; k < NP; ++k) a[k] = A[(size_t)c * NP + k];
#pragma unroll
    for (int k = 0; k < B; ++k) pv[k] = p[(size_t)c * B + k];
    sym_packed_mul<B>(a, pv, out);
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const double qv = out[k] + D[(size_t)c * B + k] * pv[k] + (yw ? yw[(size_t)c * B + k] : 0.0);
      q[(size_t)c * B + k] = qv;
      pq += qv * pv[k];
    }
  }
  write_partial(pq, part_pq, scratch);
}

// alpha = rz / pq; x += alpha p; r -= alpha q; z = Minv r; partial r.z, r.r
template <int B>
__global__ void __launch_bounds__(kPcgThreads) pcg_update(int nb, int nblk, const double* __restrict__ Minv,
                                                          const double* __restrict__ p, const double* __restrict__ q,
                                                          double* __restrict__ x, double* __restrict__ r,
                                                          double* __restrict__ z, const double* __restrict__ dots_prev,
                                                          const double* __restrict__ part_pq,
                                                          double* __restrict__ part_rz, double* __restrict__ part_rr,
                                                          double* __restrict__ dots_it) {
  constexpr int NP = B * (B + 1) / 2;
  __shared__ double scratch[32];
  __shared__ double sh;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double pq = sum_partials(part_pq, nblk, &sh);
  const double rz = dots_prev[1];
  const double alpha = (pq > 0.0) ? rz / pq : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) dots_it[0] = pq;
  double rzn = 0.0, rr = 0.0;
  if (c < nb) {
    double m[NP], rv[B], zv[B];
#pragma unroll
    for (int k = 0; k < NP; ++k) m[k] = Minv[(size_t)c * NP + k];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const size_t i = (size_t)c * B + k;
      x[i] += alpha * p[i];
      rv[k] = r[i] - alpha * q[i];
      r[i] = rv[k];
      rr += rv[k] * rv[k];
    }
    sym_packed_mul<B>(m, rv, zv);
#pragma unroll
    for (int k = 0; k < B; ++k) {
      z[(size_t)c * B + k] = zv[k];
      rzn += rv[k] * zv[k];
    }
  }
  write_partial(rzn, part_rz, scratch);
  write_partial(rr, part_rr, scratch);
}

// beta = rz_new / rz; p = z + beta p; clears the mat-vec accumulator yw;
// CTA 0 publishes dots[it][1] = r.z, dots[it][2] = r.r
template <int B>
__global__ void __launch_bounds__(kPcgThreads) pcg_direction(int nb, int nblk, const double* __restrict__ z,
                                                             double* __restrict__ p, double* __restrict__ yw,
                                                             const double* __restrict__ dots_prev,
                                                             const double* __restrict__ part_rz,
                                                             const double* __restrict__ part_rr,
                                                             double* __restrict__ dots_it) {
  __shared__ double sh, sh2;
  const double rzn = sum_partials(part_rz, nblk, &sh);
  if (blockIdx.x == 0) {
    const double rr = sum_partials(part_rr, nblk, &sh2);
    if (threadIdx.x == 0) {
      dots_it[1] = rzn;
      dots_it[2] = rr;
    }
  }
  const double rz = dots_prev[1];
  const double beta = (rz > 0.0) ? rzn / rz : 0.0;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nb) {
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const size_t i = (size_t)c * B + k;
      p[i] = z[i] + beta * p[i];
      if (yw) yw[i] = 0.0;
    }
  }
}

// x = 0; r = b; z = Minv r; p = z; yw = 0; partial r.z, r.r
template <int B>
__global__ void __launch_bounds__(kPcgThreads) pcg_init(int nb, const double* __restrict__ Minv,
                                                        const double* __restrict__ b, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ z,
                                                        double* __restrict__ p, double* __restrict__ yw,
                                                        double* __restrict__ part_rz, double* __restrict__ part_rr) {
  constexpr int NP = B * (B + 1) / 2;
  __shared__ double scratch[32];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, rr = 0.0;
  if (c < nb) {
    double m[NP], rv[B], zv[B];
#pragma unroll
    for (int k = 0; k < NP; ++k) m[k] = Minv[(size_t)c * NP + k];
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const size_t i = (size_t)c * B + k;
      rv[k] = b[i];
      x[i] = 0.0;
      r[i] = rv[k];
      if (yw) yw[i] = 0.0;
      rr += rv[k] * rv[k];
    }
    sym_packed_mul<B>(m, rv, zv);
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const size_t i = (size_t)c * B + k;
      z[i] = zv[k];
      p[i] = zv[k];
      rz += rv[k] * zv[k];
    }
  }
  write_partial(rz, part_rz, scratch);
  write_partial(rr, part_rr, scratch);
}

// dots0[1] = sum part_rz, dots0[2] = sum part_rr   (single CTA)
__global__ void __launch_bounds__(kPcgThreads) pcg_publish_init(int nblk, const double* __restrict__ part_rz,
                                                                const double* __restrict__ part_rr,
                                                                double* __restrict__ dots0) {
  __shared__ double sh, sh2;
  const double rz = sum_partials(part_rz, nblk, &sh);
  const double rr = sum_partials(part_rr, nblk, &sh2);
  if (threadIdx.x == 0) {
    dots0[0] = 0.0;
    dots0[1] = rz;
    dots0[2] = rr;
  }
}

}  // namespace b200
