// ba_kernels_v2.cuh -- "compact row" layout of the BA hot path (design v2).
//
// v1 stores the 6x3 block W_o = J_cam^T J_pt (144 B/observation) and scatters
// W_o z_p into y[cam] with 6 FP64 atomics per observation; ncu shows that
// mat-vec bound by L2 operations (profiles/r1_matvec_ncu.md).  v2 stores what
// W is made of instead -- per observation the robustified projective Jacobian
// J = sqrt(rho') d(pi)/d(X_c) (2x3) and a = R X (3) -- in BOTH traversal orders
//     Jp[N][10]  point order  {J, a, pad}        80 B, tiles moved by TMA
//     Jc[Nv][10] camera order {J, a, pad}        80 B
// and applies W = J_c^T J_p on the fly, J_c = [-2 J [a]x | J] (masked),
// J_p = J R(q_cam).  The implicit-Schur mat-vec becomes two streaming passes
//     pass A (point order):  s_p = sum_o R^T J^T (J v_o),  z_p = Vinv s_p -> z[P][4]
//                            v_o = m_t x_t - 2 m_r (a x x_r)
//     pass B (camera order): y_c -= sum_o J_c^T (J (R z_p))   one warp per <= 256-observation
//                            segment of ONE camera: register accumulation, shuffle
//                            reduction, 6 atomics per segment instead of per observation.
// All arithmetic stays FP64; only the traffic changes:
//     bytes/mat-vec = 84 N + 80 P (pass A) + 84 N + 32 P (pass B)  vs  152 N + 6 FP64 RED per observation.
#pragma once
#include "ba_kernels.cuh"

namespace b200 {

constexpr int kXq = 12;   // per-camera gather record of pass A: x(6) q(4) mask pad -> 96 B

struct BAViewV2 {
  const double* Jp;   // [N][10]   (aliases BAView::W)
  double* Jc;         // [Nv][10]
  double* z4;         // [P][4]
};

// xq[c] = {x_c (6), q_c (4), mask, 0}
__global__ void ba2_pack_xq(int C, const double* __restrict__ x, const double* __restrict__ cam_rec,
                            double* __restrict__ xq) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double* o = xq + (size_t)c * kXq;
  const double* r = cam_rec + (size_t)c * kCamRec;
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = x[(size_t)c * 6 + k];
  o[6] = r[0]; o[7] = r[1]; o[8] = r[2]; o[9] = r[3];
  o[10] = (double)(__double_as_longlong(r[7]) & 0xff);
  o[11] = 0.0;
}

// ---------------------------------------------------------------------------
// camera-order linearisation: U_c, g_c AND the camera-order rows Jc
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ba2_linearize_cams(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec,
                                                         const double* __restrict__ intr_rec,
                                                         const double* __restrict__ points, double huber_a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= v.n_segs) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  const double4 q4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec);
  const double4 t4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec + 4);
  const double* irc = intr_rec + (size_t)cam_rec_intr(t4c) * kIntrRec;
  double U[21], g[6];
#pragma unroll
  for (int k = 0; k < 21; ++k) U[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  for (int i = b + lane; i < e; i += 32) {
    const int pt = v.pt_c[i];
    const double2 xy = v.xy_c[i];
    const double X0 = points[3 * (size_t)pt], X1 = points[3 * (size_t)pt + 1], X2 = points[3 * (size_t)pt + 2];
    ObsLin o;
    linearize_obs(q4c, t4c, irc, X0, X1, X2, xy, huber_a, o);
    double2* row = reinterpret_cast<double2*>(v2.Jc + (size_t)i * kJcDoubles);
    row[0] = make_double2(o.J[0], o.J[1]);
    row[1] = make_double2(o.J[2], o.J[3]);
    row[2] = make_double2(o.J[4], o.J[5]);
    row[3] = make_double2(o.a[0], o.a[1]);
    row[4] = make_double2(o.a[2], 0.0);
    double Jc[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Jc[a][k] = o.Jr[3 * a + k];
        Jc[a][3 + k] = o.Jt[3 * a + k];
      }
    int idx = 0;
#pragma unroll
    for (int i2 = 0; i2 < 6; ++i2) {
#pragma unroll
      for (int j = i2; j < 6; ++j) U[idx++] += Jc[0][i2] * Jc[0][j] + Jc[1][i2] * Jc[1][j];
      g[i2] += Jc[0][i2] * o.r[0] + Jc[1][i2] * o.r[1];
    }
  }
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    const double s = warp_sum(U[k]);
    if (lane == k && s != 0.0) atomicAdd(&v.U[(size_t)cam * 21 + k], s);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double s = warp_sum(g[k]);
    if (lane == 21 + k && s != 0.0) atomicAdd(&v.gc[(size_t)cam * 6 + k], s);
  }
}

// J_c = [-2 J [a]x (masked) | J (masked)] from a compact row
__device__ __forceinline__ void jc_from_row(const double J[6], const double a[3], int mask, double Jc[2][6]) {
  const bool rvar = !(mask & 1), tvar = !(mask & 2);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double j0 = J[3 * r], j1 = J[3 * r + 1], j2 = J[3 * r + 2];
    Jc[r][0] = rvar ? -2.0 * (j1 * a[2] - j2 * a[1]) : 0.0;
    Jc[r][1] = rvar ? -2.0 * (j2 * a[0] - j0 * a[2]) : 0.0;
    Jc[r][2] = rvar ? -2.0 * (j0 * a[1] - j1 * a[0]) : 0.0;
    Jc[r][3] = tvar ? j0 : 0.0;
    Jc[r][4] = tvar ? j1 : 0.0;
    Jc[r][5] = tvar ? j2 : 0.0;
  }
}

// ---------------------------------------------------------------------------
// Schur-Jacobi diagonal from the camera-order rows:
//   Sd_c = sum_o J_c^T (J R Vinv_p R^T J^T) J_c
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ba2_schur_diag(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= v.n_segs) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  const double4 q4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec);
  const double4 t4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec + 4);
  const int mask = (int)(__double_as_longlong(t4c.w) & 0xff);
  const double q[4] = {q4c.x, q4c.y, q4c.z, q4c.w};
  double R[9];
  quat_to_R(q, R);
  double S[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) S[k] = 0.0;
  for (int i = b + lane; i < e; i += 32) {
    const double2* row = reinterpret_cast<const double2*>(v2.Jc + (size_t)i * kJcDoubles);
    const double2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
    const double J[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
    const double a[3] = {r3.x, r3.y, r4.x};
    const int pt = v.pt_c[i];
    const double2* vp = reinterpret_cast<const double2*>(v.Vinv + (size_t)pt * 6);
    const double2 v0 = vp[0], v1 = vp[1], v2_ = vp[2];
    const double vi[6] = {v0.x, v0.y, v1.x, v1.y, v2_.x, v2_.y};
    // Jp = J R (2x3);  T = Jp Vinv (2x3);  M2 = T Jp^T (2x2 sym)
    double Jp[2][3], T[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Jp[r][c] = J[3 * r] * R[c] + J[3 * r + 1] * R[3 + c] + J[3 * r + 2] * R[6 + c];
#pragma unroll
    for (int r = 0; r < 2; ++r) sym3_mul(vi, Jp[r], T[r]);
    const double m00 = T[0][0] * Jp[0][0] + T[0][1] * Jp[0][1] + T[0][2] * Jp[0][2];
    const double m01 = T[0][0] * Jp[1][0] + T[0][1] * Jp[1][1] + T[0][2] * Jp[1][2];
    const double m11 = T[1][0] * Jp[1][0] + T[1][1] * Jp[1][1] + T[1][2] * Jp[1][2];
    double Jc[2][6];
    jc_from_row(J, a, mask, Jc);
    // S += Jc^T M2 Jc
    double A0[6], A1[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      A0[k] = m00 * Jc[0][k] + m01 * Jc[1][k];
      A1[k] = m01 * Jc[0][k] + m11 * Jc[1][k];
    }
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) S[idx++] += Jc[0][r] * A0[c] + Jc[1][r] * A1[c];
  }
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    const double s = warp_sum(S[k]);
    if (lane == k && s != 0.0) atomicAdd(&v.Sd[(size_t)cam * 21 + k], s);
  }
}

// ---------------------------------------------------------------------------
// pass A (point order):  s_p = [g_p] + sum_o J_p^T (J_c x_c);  z_p = Vinv s_p
//   MODE 0: z -> z4[P][4]                      (mat-vec)
//   MODE 2: back-substitution epilogue (points_new, step scalars), as ba_schur_pass<2>
// ---------------------------------------------------------------------------
struct K3v2Smem {
  alignas(128) double Jt[kTile * kJpDoubles];
  double t[3][kTile + 1];
  double z[3][kTilePts + 1];
  unsigned pb[kTilePts + 1];
  double scratch[32];
  alignas(8) uint64_t mbar;
};

template <int MODE>
__global__ void __launch_bounds__(kTile, B200_K3_MIN_CTAS) ba2_pass_a(BAView v, BAViewV2 v2, const double* __restrict__ xq,
                                                                     const double* __restrict__ points,
                                                                     double* __restrict__ points_new, double radius,
                                                                     double* __restrict__ bscal) {
  extern __shared__ unsigned char smem_raw[];
  K3v2Smem& sm = *reinterpret_cast<K3v2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int tile = blockIdx.x;
  const int tid = threadIdx.x;
  const int4 td = v.tile_desc[tile];
  const int p0 = td.x, npts = td.y, n = td.w;
  const unsigned o0 = (unsigned)td.z, o1 = o0 + (unsigned)n;
  const int nchunks = (n + kTile - 1) / kTile;
  constexpr uint32_t kRowBytes = kJpDoubles * 8;
  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    fence_mbar_init();
    if (n > 0) {
      const int nc0 = min(kTile, n);
      mbar_arrive_expect_tx(&sm.mbar, (uint32_t)nc0 * kRowBytes);
      tma_load_1d(sm.Jt, v2.Jp + (size_t)o0 * kJpDoubles, (uint32_t)nc0 * kRowBytes, &sm.mbar);
    }
  }
  // prefetch the camera record of the first chunk
  double2 g0 = make_double2(0, 0), g1 = g0, g2 = g0, g3 = g0, g4 = g0, g5 = g0;
  if (tid < n) {
    const int cam = v.obs_cam[o0 + tid];
    const double2* gp_ = reinterpret_cast<const double2*>(xq + (size_t)cam * kXq);
    g0 = gp_[0]; g1 = gp_[1]; g2 = gp_[2]; g3 = gp_[3]; g4 = gp_[4]; g5 = gp_[5];
  }
  if (tid < npts) {
    sm.pb[tid] = v.pt_begin[p0 + tid];
    if (tid == npts - 1) sm.pb[npts] = o1;
    sm.z[0][tid] = sm.z[1][tid] = sm.z[2][tid] = 0.0;
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * kTile;
    const int nc = min(kTile, n - c0);
    if (tid == 0 && ch > 0) {
      mbar_arrive_expect_tx(&sm.mbar, (uint32_t)nc * kRowBytes);
      tma_load_1d(sm.Jt, v2.Jp + (size_t)(o0 + c0) * kJpDoubles, (uint32_t)nc * kRowBytes, &sm.mbar);
    }
    const bool active = tid < nc;
    if (active && ch > 0) {
      const int cam = v.obs_cam[o0 + c0 + tid];
      const double2* gp_ = reinterpret_cast<const double2*>(xq + (size_t)cam * kXq);
      g0 = gp_[0]; g1 = gp_[1]; g2 = gp_[2]; g3 = gp_[3]; g4 = gp_[4]; g5 = gp_[5];
    }
    mbar_wait(&sm.mbar, phase);
    phase ^= 1;
    double t0 = 0, t1 = 0, t2 = 0;
    if (active) {
      const double2* jr = reinterpret_cast<const double2*>(sm.Jt + tid * kJpDoubles);
      const double2 r0 = jr[0], r1 = jr[1], r2 = jr[2], r3 = jr[3], r4 = jr[4];
      const double J[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
      const double a[3] = {r3.x, r3.y, r4.x};
      const double xr[3] = {g0.x, g0.y, g1.x}, xt[3] = {g1.y, g2.x, g2.y};
      const double q[4] = {g3.x, g3.y, g4.x, g4.y};
      const int mask = (int)g5.x;
      const double mr = (mask & 1) ? 0.0 : 1.0, mt = (mask & 2) ? 0.0 : 1.0;
      // v = m_t x_t - 2 m_r (a x x_r)
      const double vv[3] = {mt * xt[0] - 2.0 * mr * (a[1] * xr[2] - a[2] * xr[1]),
                            mt * xt[1] - 2.0 * mr * (a[2] * xr[0] - a[0] * xr[2]),
                            mt * xt[2] - 2.0 * mr * (a[0] * xr[1] - a[1] * xr[0])};
      const double u0 = J[0] * vv[0] + J[1] * vv[1] + J[2] * vv[2];
      const double u1 = J[3] * vv[0] + J[4] * vv[1] + J[5] * vv[2];
      const double h[3] = {J[0] * u0 + J[3] * u1, J[1] * u0 + J[4] * u1, J[2] * u0 + J[5] * u1};
      double R[9];
      quat_to_R(q, R);
      t0 = R[0] * h[0] + R[3] * h[1] + R[6] * h[2];   // R^T h
      t1 = R[1] * h[0] + R[4] * h[1] + R[7] * h[2];
      t2 = R[2] * h[0] + R[5] * h[1] + R[8] * h[2];
    }
    sm.t[0][tid] = t0;
    sm.t[1][tid] = t1;
    sm.t[2][tid] = t2;
    __syncthreads();
    for (int item = tid; item < npts * 3; item += kTile) {
      const int j = item / 3, k = item - 3 * j;
      const int lo = max((int)sm.pb[j] - (int)(o0 + c0), 0), hi = min((int)sm.pb[j + 1] - (int)(o0 + c0), nc);
      double acc = 0.0;
      for (int i = lo; i < hi; ++i) acc += sm.t[k][i];
      sm.z[k][j] += acc;
    }
    __syncthreads();
  }
  double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if (tid < npts) {
    const size_t p = (size_t)(p0 + tid);
    const bool pvalid = (int)(sm.pb[tid + 1] - sm.pb[tid]) >= v.min_views;
    double z[3] = {0.0, 0.0, 0.0};
    if (pvalid) {
      double s[3] = {sm.z[0][tid], sm.z[1][tid], sm.z[2][tid]};
      double g[3] = {0, 0, 0};
      if (MODE != 0) {
        g[0] = v.gp[3 * p]; g[1] = v.gp[3 * p + 1]; g[2] = v.gp[3 * p + 2];
        s[0] += g[0]; s[1] += g[1]; s[2] += g[2];
      }
      double vi[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) vi[k] = v.Vinv[6 * p + k];
      sym3_mul(vi, s, z);
      if (MODE == 2) {
        double v6[6], js[3], Dp[3];
#pragma unroll
        for (int k = 0; k < 6; ++k) v6[k] = v.V[6 * p + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) js[k] = v.jscale_p[3 * p + k];
        point_damping(v6, js, radius, Dp);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double dp = -z[k];
          const double xo = points[3 * p + k];
          points_new[3 * p + k] = xo + dp;
          b0 += g[k] * dp;
          b1 += Dp[k] * dp * dp;
          b2 += dp * dp;
          b3 += xo * xo;
        }
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 3; ++k) points_new[3 * p + k] = points[3 * p + k];
    }
    if (MODE == 0) *reinterpret_cast<double4*>(v2.z4 + 4 * p) = make_double4(z[0], z[1], z[2], 0.0);
  }
  if (MODE == 2) {
    b0 = block_sum(b0, sm.scratch);
    b1 = block_sum(b1, sm.scratch);
    b2 = block_sum(b2, sm.scratch);
    b3 = block_sum(b3, sm.scratch);
    if (tid == 0) {
      atomicAdd(&bscal[0], b0);
      atomicAdd(&bscal[1], b1);
      atomicAdd(&bscal[2], b2);
      atomicAdd(&bscal[3], b3);
    }
  }
}

// z4[p] = Vinv_p g_p   (right-hand side: no observation pass needed)
__global__ void ba2_point_rhs_z(BAView v, BAViewV2 v2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= v.P) return;
  double vi[6], g[3], z[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) vi[k] = v.Vinv[6 * (size_t)p + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) g[k] = v.gp[3 * (size_t)p + k];
  sym3_mul(vi, g, z);
  const bool pvalid = (int)(v.pt_begin[p + 1] - v.pt_begin[p]) >= v.min_views;
  *reinterpret_cast<double4*>(v2.z4 + 4 * (size_t)p) = pvalid ? make_double4(z[0], z[1], z[2], 0.0) : make_double4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------
// pass B (camera order): y_c -= sum_{o in segment} J_c^T ( J ( R z_p ) )
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ba2_pass_b(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec,
                                                 double* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= v.n_segs) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  const double4 q4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec);
  const double4 t4c = *reinterpret_cast<const double4*>(cam_rec + (size_t)cam * kCamRec + 4);
  const int mask = (int)(__double_as_longlong(t4c.w) & 0xff);
  const double mr = (mask & 1) ? 0.0 : 1.0, mt = (mask & 2) ? 0.0 : 1.0;
  const double q[4] = {q4c.x, q4c.y, q4c.z, q4c.w};
  double R[9];
  quat_to_R(q, R);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = b + lane; i < e; i += 32) {
    const int pt = v.pt_c[i];
    const double4 z = *reinterpret_cast<const double4*>(v2.z4 + 4 * (size_t)pt);
    const double2* row = reinterpret_cast<const double2*>(v2.Jc + (size_t)i * kJcDoubles);
    const double2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];
    const double J[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
    const double a[3] = {r3.x, r3.y, r4.x};
    const double bz[3] = {R[0] * z.x + R[1] * z.y + R[2] * z.z, R[3] * z.x + R[4] * z.y + R[5] * z.z,
                          R[6] * z.x + R[7] * z.y + R[8] * z.z};
    const double u0 = J[0] * bz[0] + J[1] * bz[1] + J[2] * bz[2];
    const double u1 = J[3] * bz[0] + J[4] * bz[1] + J[5] * bz[2];
    const double h[3] = {J[0] * u0 + J[3] * u1, J[1] * u0 + J[4] * u1, J[2] * u0 + J[5] * u1};
    // J_r^T u = 2 a x h ; J_t^T u = h
    acc[0] += 2.0 * (a[1] * h[2] - a[2] * h[1]);
    acc[1] += 2.0 * (a[2] * h[0] - a[0] * h[2]);
    acc[2] += 2.0 * (a[0] * h[1] - a[1] * h[0]);
    acc[3] += h[0];
    acc[4] += h[1];
    acc[5] += h[2];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double s = warp_sum(acc[k]) * (k < 3 ? mr : mt);
    if (lane == k && s != 0.0) atomicAdd(&y[(size_t)cam * 6 + k], -s);
  }
}

}  // namespace b200
