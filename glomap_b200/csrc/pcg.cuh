// pcg.cuh -- block-preconditioned conjugate gradients on the reduced camera
// system, entirely stream-ordered on the device (no host sync inside an
// iteration).  The reference factors this system with CHOLMOD
// (bundle_adjustment.cc:94-96, global_positioning.cc:551-559); north_star
// mandates PCG with one all-reduce per mat-vec.
//
// Scalars live in a device array dots[it][4] = {p.q, r.z (next), r.r (next), -}
// indexed by iteration so that no kernel ever resets a value another kernel of
// the same iteration still reads.
#pragma once
#include "common.cuh"

namespace b200 {

// q_c = (A_c + diag(D_c)) p_c + yw_c ; dots[it][0] += p.q
//   A packed symmetric B x B per block, yw = the (negative) Schur part already
//   accumulated by the mat-vec kernel (and all-reduced).
template <int B>
__global__ void pcg_apply_diag(int nb, const double* __restrict__ A, const double* __restrict__ D,
                               const double* __restrict__ p, const double* __restrict__ yw, double* __restrict__ q,
                               double* __restrict__ dots_it) {
  constexpr int NP = B * (B + 1) / 2;
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
  pq = warp_sum(pq);
  if ((threadIdx.x & 31) == 0 && pq != 0.0) atomicAdd(&dots_it[0], pq);
}

// alpha = rz / pq; x += alpha p; r -= alpha q; z = Minv r; dots[it][1] += r.z; dots[it][2] += r.r
template <int B>
__global__ void pcg_update(int nb, const double* __restrict__ Minv, const double* __restrict__ p,
                           const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r,
                           double* __restrict__ z, const double* __restrict__ rz_ptr, double* __restrict__ dots_it) {
  constexpr int NP = B * (B + 1) / 2;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const double pq = dots_it[0];
  const double rz = *rz_ptr;
  const double alpha = (pq > 0.0) ? rz / pq : 0.0;
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
  rzn = warp_sum(rzn);
  rr = warp_sum(rr);
  if ((threadIdx.x & 31) == 0) {
    if (rzn != 0.0) atomicAdd(&dots_it[1], rzn);
    if (rr != 0.0) atomicAdd(&dots_it[2], rr);
  }
}

// beta = rz_new / rz; p = z + beta p; also clears the mat-vec accumulator yw
template <int B>
__global__ void pcg_direction(int nb, const double* __restrict__ z, double* __restrict__ p, double* __restrict__ yw,
                              const double* __restrict__ rz_ptr, const double* __restrict__ dots_it) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * B) return;
  const double rz = *rz_ptr;
  const double beta = (rz > 0.0) ? dots_it[1] / rz : 0.0;
  p[i] = z[i] + beta * p[i];
  if (yw) yw[i] = 0.0;
}

// x = 0; r = b; z = Minv r; p = z; init[1] = r.z; init[2] = r.r; yw = 0
template <int B>
__global__ void pcg_init(int nb, const double* __restrict__ Minv, const double* __restrict__ b, double* __restrict__ x,
                         double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                         double* __restrict__ yw, double* __restrict__ dots0) {
  constexpr int NP = B * (B + 1) / 2;
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
  rz = warp_sum(rz);
  rr = warp_sum(rr);
  if ((threadIdx.x & 31) == 0) {
    if (rz != 0.0) atomicAdd(&dots0[1], rz);
    if (rr != 0.0) atomicAdd(&dots0[2], rr);
  }
}

}  // namespace b200
