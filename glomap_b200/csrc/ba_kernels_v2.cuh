// ba_kernels_v2.cuh -- "compact row" layout of the BA hot path (design v2).
//
// v1 stores the 6x3 block W_o = J_cam^T J_pt (144 B/observation) and scatters
// W_o z_p into y[cam] with 6 FP64 atomics per observation; ncu shows that
// mat-vec bound by L2 operations (profiles/r1_matvec_ncu.md).  v2 exploits
//     J_cam = J_pt G,   G = [ -2 [X]x R^T | R^T ]   (3 x 6;  X = the world point, R = R(q_cam))
// (left-perturbed rotation, translation), so that with the symmetric 3x3
//     A_o = J_pt^T J_pt        (what the observation adds to V_p)
// W_o = G^T A_o,  W_o^T x = A_o v  with  v = R^T x_t - 2 X x (R^T x_r),
// W_o z = [ 2 R (X x w) ; R w ],  w = A_o z.   Only A_o (48 B) is stored per
// observation, in BOTH traversal orders:
//     Ap[N][6]   point order  (tiles moved by TMA)
//     Ac         camera order {A_o, X_p} (72 B, written by the camera-order linearisation), SoA in
//                groups of 32 rows: element k of padded row r at ((r >> 5) * 9 + k) * 32 + (r & 31); every
//                segment starts on a group boundary (seg_row0), so lane l of the segment's warp owns
//                rows l, l + 32, ... and every load/store of the warp is one contiguous 256-B line pair
// and the implicit-Schur mat-vec is two streaming passes
//     pass A (point order):  s_p = sum_o A_o v_o,  z_p = Vinv s_p -> z4[P]      gathers R^T x (48 B)
//     pass B (camera order): y_c -= R-rotated sum_o [2 X x (A_o z_p) ; A_o z_p]  gathers z_p (32 B)
//                            one warp per <= 256-observation segment of ONE camera: register
//                            accumulation, shuffle reduction, 6 atomics per segment.
// All arithmetic stays FP64.  Algorithmic bytes per mat-vec: 52 N + 80 P (A) + 76 N + 32 P (B).
#pragma once
#include "ba_kernels.cuh"
#include "pcg.cuh"

namespace b200 {

struct BAViewV2 {
  const double* Ap;   // [N][6]   (aliases BAView::W)
  double* Ac;         // camera-order rows, SoA-32 (see above)
  double* z4;         // [P][4]
  // stored-row intrinsics path (NK > 0, see below): B_o = rho' J_pt^T J_k in camera order, the frame x intrinsics
  // cross blocks of every image, the variable-parameter table and the number of frames (block C + k = intrinsics k)
  double* Bc = nullptr;              // [rows32][3 * NK][32]
  double* Ufk = nullptr;             // [C][6][NK]
  const IntrVarRec* ivar = nullptr;  // [K]
  int C = 0;
};

// ---------------------------------------------------------------------------
// Variable intrinsics WITHOUT recomputing the projection chain in the mat-vec ("stored-row" path, NK <= 2 variable
// parameters per camera: SIMPLE_PINHOLE f; SIMPLE_RADIAL f, k; PINHOLE fx, fy -- the reference default
// optimize_intrinsics = true, optimize_principal_point = false, bundle_adjustment.cc:273-293).
// Intrinsics block k is pseudo-camera block C + k of the reduced system (ba_kernels_ext.cuh).  With J_k = d e / d(params)
// (2 x NK) and B_o = rho' J_pt^T J_k (3 x NK, stored next to A_o in BOTH orders, 24 NK bytes each):
//     W^T x   per point:   s_p = sum_o ( A_o w_o + B_o x_k(o) )                    (pass A, x_k rides in the xq record)
//     W z     per image:   y_f -= G^T sum_o A_o z_p,   y_k -= sum_o B_o^T z_p       (pass B)
//     U x:    block diagonal U_ff, U_kk by pcg_apply_diag; the frame x intrinsics coupling is ONE 6 x NK block per image
//             (an image has one camera):  y_f += U_fk x_k,  y_k += U_fk^T x_f       (ba2k_cross, C threads)
// so the per-iteration cost over the constant-intrinsics path is 48 NK bytes per observation of streamed rows.
// ---------------------------------------------------------------------------
template <int NK>
__device__ __forceinline__ void obs_intr_rows(const ObsCore& o, const double* __restrict__ ir, const IntrVarRec& iv,
                                              const double Jp[6], double Jk[2][NK > 0 ? NK : 1],
                                              double B[NK > 0 ? 3 * NK : 1]) {
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    double jx = 0.0, jy = 0.0;
    if (j < iv.mb) intr_param_jac(ir, iv.pidx[j], o.uv[0], o.uv[1], 1.0, jx, jy);
    if (!o.valid) jx = jy = 0.0;
    Jk[0][j] = jx;
    Jk[1][j] = jy;
#pragma unroll
    for (int c = 0; c < 3; ++c) B[3 * j + c] = o.rho1 * (Jp[c] * jx + Jp[3 + c] * jy);
  }
}

// xp[c] = { R^T x_r , R^T x_t, pad, pad }: 64-B rows, so pass A gathers a camera with one 256-bit and one
// 128-bit load out of a single line  (masked dofs of x are zero already: PCG keeps them at 0)
constexpr int kXqStride = 8;
__global__ void ba2_pack_x(int C, const double* __restrict__ x, const double* __restrict__ cam_rec,
                           double* __restrict__ xp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* r = cam_rec + (size_t)c * kCamRec;
  const double q[4] = {r[0], r[1], r[2], r[3]};
  double R[9];
  quat_to_R(q, R);
  const double* xc = x + (size_t)c * 6;
  double* o = xp + (size_t)c * kXqStride;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = R[k] * xc[0] + R[3 + k] * xc[1] + R[6 + k] * xc[2];
    o[3 + k] = R[k] * xc[3] + R[3 + k] * xc[4] + R[6 + k] * xc[5];
  }
}

// Head of a PCG iteration fused with the packing of its search direction: p = z + beta p (pcg_direction) and
// xp[c] = {R^T p_r, R^T p_t} for pass A, one thread per camera -- one launch instead of two per iteration.
__global__ void __launch_bounds__(kPcgThreads) ba2_pcg_direction_pack(int nb, int nblk, int it, int min_it, double rel_tol,
                                                                      const double* __restrict__ z, double* __restrict__ p,
                                                                      double* __restrict__ yw,
                                                                      const double* __restrict__ dots_pp,
                                                                      const double* __restrict__ part_rz,
                                                                      const double* __restrict__ part_rr,
                                                                      double* __restrict__ dots_pub, PcgCtl* __restrict__ ctl,
                                                                      const double* __restrict__ cam_rec,
                                                                      double* __restrict__ xp, int n_pack) {
  __shared__ double sh3[3];
  double beta;
  if (!pcg_direction_head(nblk, it, min_it, rel_tol, dots_pp, part_rz, part_rr, nullptr, dots_pub, ctl, sh3, beta)) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb) return;
  double pv[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const size_t i = (size_t)c * 6 + k;
    pv[k] = (it == 1) ? z[i] : z[i] + beta * p[i];
    p[i] = pv[k];
    yw[i] = 0.0;
  }
  if (c >= n_pack) return;   // pseudo-camera blocks (intrinsics) have no record: their x rides in the frames' rows
  const double* r = cam_rec + (size_t)c * kCamRec;
  const double q[4] = {r[0], r[1], r[2], r[3]};
  double R[9];
  quat_to_R(q, R);
  double* o = xp + (size_t)c * kXqStride;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = R[k] * pv[0] + R[3 + k] * pv[1] + R[6 + k] * pv[2];
    o[3 + k] = R[k] * pv[3] + R[3 + k] * pv[4] + R[6 + k] * pv[5];
  }
}

// pts4[p] = {X_p, 0}: 32-B rows for the camera-order gathers (one LDG.E.256 per observation)
__global__ void ba2_pad_points(int P, const double* __restrict__ points, double* __restrict__ pts4) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double x = points[3 * (size_t)p], y = points[3 * (size_t)p + 1], z = points[3 * (size_t)p + 2];
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(pts4 + 4 * (size_t)p), "d"(x), "d"(y), "d"(z), "d"(0.0) : "memory");
}

// ---------------------------------------------------------------------------
// camera-order linearisation: U_c, g_c AND the camera-order rows Ac = {A_o, X_p}
// ---------------------------------------------------------------------------
template <int NK>
__global__ void __launch_bounds__(128, NK > 0 ? 3 : B200_LC_MIN_CTAS) ba2_linearize_cams(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec,
                                                         const double* __restrict__ intr_rec,
                                                         const double* __restrict__ /*points: read through v2.z4 (padded copy)*/, double huber_a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= v.n_segs) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  const double4 q4c = ld_rec32(cam_rec + (size_t)cam * kCamRec);
  const double4 t4c = ld_rec32(cam_rec + (size_t)cam * kCamRec + 4);
  const double* irc = intr_rec + (size_t)v.seg_intr[warp] * kIntrRec;
  const double* src = sensor_of_seg(v, warp);
  double U[21], g[6];
#pragma unroll
  for (int k = 0; k < 21; ++k) U[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  const int cmask = (int)(__double_as_longlong(t4c.w) & 0xff);
  const bool tvar = !(cmask & 2), rvar = !(cmask & 1);
  double* row = v2.Ac + (size_t)(v.seg_row0[warp] >> 5) * (kJcDoubles * 32) + lane;
  // stored-row intrinsics path: B_o rows, U_kk / g_k of the segment's intrinsics block and the image's 6 x NK cross block
  constexpr int NKK = NK > 0 ? NK : 1;
  const int blk = v.seg_intr[warp];
  IntrVarRec iv{};
  if (NK > 0) iv = v2.ivar[blk];
  double* rowB = NK > 0 ? v2.Bc + (size_t)(v.seg_row0[warp] >> 5) * (3 * NK * 32) + lane : nullptr;
  double Ukk[NKK * (NKK + 1) / 2], gk[NKK], Ufk[6][NKK];
#pragma unroll
  for (int k = 0; k < NKK * (NKK + 1) / 2; ++k) Ukk[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NKK; ++k) {
    gk[k] = 0.0;
#pragma unroll
    for (int i2 = 0; i2 < 6; ++i2) Ufk[i2][k] = 0.0;
  }
  // index -> point gather -> ~400 instructions: ncu (r2b) shows 1/3 of all stall samples on the first use of the gathered
  // point and on the address computed from the streamed index.  Register pipeline: the index of iteration + 2 and the
  // point / pixel of iteration + 1 are in flight while iteration + 0 is computed.  The points come from the 32-B padded
  // copy (pts4 = the z4 buffer, idle during linearisation): ONE 256-bit gather instead of three 64-bit ones
  // (96 -> 32 L1 wavefronts per warp and observation).
  const double* __restrict__ pts4 = v2.z4;
  int i = b + lane;
  int pt_nxt = 0;
  double2 xy = make_double2(0, 0);
  double4 Xc = make_double4(0, 0, 0, 0);
  if (i < e) {
    const int pt0 = ld_stream(v.pt_c + i);
    if (i + 32 < e) pt_nxt = ld_stream(v.pt_c + i + 32);
    xy = ld_stream(v.xy_c + i);
    Xc = ld_rec32(pts4 + 4 * (size_t)pt0);
  }
  for (; i < e; i += 32, row += kJcDoubles * 32) {
    double4 Xn = Xc;
    double2 xyn = xy;
    int pt_nn = 0;
    if (i + 32 < e) {
      Xn = ld_rec32(pts4 + 4 * (size_t)pt_nxt);
      xyn = ld_stream(v.xy_c + i + 32);
    }
    if (i + 64 < e) pt_nn = ld_stream(v.pt_c + i + 64);
    const double X0 = Xc.x, X1 = Xc.y, X2 = Xc.z;
    ObsCore o;
    obs_core(q4c, t4c, irc, src, X0, X1, X2, xy, huber_a, o);
    double Jp[6], A[6], bo[3];
    obs_point_blocks(o, Jp, A, bo);
#pragma unroll
    for (int k = 0; k < 6; ++k) st_stream(row + 32 * k, A[k]);
    st_stream(row + 192, X0);
    st_stream(row + 224, X1);
    st_stream(row + 256, X2);
    // camera blocks: J_t = J, J_r = J (-2 [R X]x)  (EigenQuaternionManifold: left perturbation of angle 2|d|), masked;
    // U += rho' Jc^T Jc, g += rho' Jc^T e
    double Jc[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double j0 = o.J[3 * a], j1 = o.J[3 * a + 1], j2 = o.J[3 * a + 2];
      Jc[a][0] = rvar ? -2.0 * (j1 * o.RX[2] - j2 * o.RX[1]) : 0.0;
      Jc[a][1] = rvar ? -2.0 * (j2 * o.RX[0] - j0 * o.RX[2]) : 0.0;
      Jc[a][2] = rvar ? -2.0 * (j0 * o.RX[1] - j1 * o.RX[0]) : 0.0;
      Jc[a][3] = tvar ? j0 : 0.0;
      Jc[a][4] = tvar ? j1 : 0.0;
      Jc[a][5] = tvar ? j2 : 0.0;
    }
    const double e0 = o.rho1 * o.e[0], e1 = o.rho1 * o.e[1];
    int idx = 0;
#pragma unroll
    for (int i2 = 0; i2 < 6; ++i2) {
      const double s0 = o.rho1 * Jc[0][i2], s1 = o.rho1 * Jc[1][i2];
#pragma unroll
      for (int j = i2; j < 6; ++j) U[idx++] += s0 * Jc[0][j] + s1 * Jc[1][j];
      g[i2] += Jc[0][i2] * e0 + Jc[1][i2] * e1;
    }
    if (NK > 0) {
      double Jk[2][NKK], Bo[3 * NKK];
      obs_intr_rows<NK>(o, irc, iv, Jp, Jk, Bo);
#pragma unroll
      for (int k = 0; k < 3 * NK; ++k) st_stream(rowB + 32 * k, Bo[k]);
      rowB += 3 * NK * 32;
      int ik = 0;
#pragma unroll
      for (int a = 0; a < NK; ++a) {
        const double s0 = o.rho1 * Jk[0][a], s1 = o.rho1 * Jk[1][a];
#pragma unroll
        for (int c = a; c < NK; ++c) Ukk[ik++] += s0 * Jk[0][c] + s1 * Jk[1][c];
        gk[a] += Jk[0][a] * e0 + Jk[1][a] * e1;
#pragma unroll
        for (int i2 = 0; i2 < 6; ++i2) Ufk[i2][a] += Jc[0][i2] * s0 + Jc[1][i2] * s1;
      }
    }
    Xc = Xn; xy = xyn; pt_nxt = pt_nn;
  }
  if (NK > 0) {
    const size_t kb = (size_t)(v2.C + blk);
    int ik = 0;
#pragma unroll
    for (int a = 0; a < NK; ++a) {
#pragma unroll
      for (int c = a; c < NK; ++c) {
        const double sacc = warp_sum(Ukk[ik++]);
        if (lane == 0 && sacc != 0.0) atomicAdd(&v.U[kb * 21 + sym_idx(6, a, c)], sacc);
      }
      const double sg = warp_sum(gk[a]);
      if (lane == 0 && sg != 0.0) atomicAdd(&v.gc[kb * 6 + a], sg);
#pragma unroll
      for (int i2 = 0; i2 < 6; ++i2) {
        const double sf = warp_sum(Ufk[i2][a]);
        if (lane == 0 && sf != 0.0) atomicAdd(&v2.Ufk[((size_t)cam * 6 + i2) * NK + a], sf);
      }
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

// ---------------------------------------------------------------------------
// Schur-Jacobi diagonal from the camera-order rows, accumulated in the world
// frame:  Sd_c = Rb ( sum_o Gh^T N Gh ) Rb^T,  N = A_o Vinv_p A_o,
//         Gh = [ -2 [X]x | I ],  Rb = blockdiag(R, R)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ba2_schur_diag(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= v.n_segs) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  // world-frame accumulators: RR (sym 6), RT (full 9), TT (sym 6)
  double RR[6] = {0, 0, 0, 0, 0, 0}, RT[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, TT[6] = {0, 0, 0, 0, 0, 0};
  const double* row = v2.Ac + (size_t)(v.seg_row0[warp] >> 5) * (kJcDoubles * 32) + lane;
  for (int i = b + lane; i < e; i += 32, row += kJcDoubles * 32) {
    const double A[6] = {ld_stream(row), ld_stream(row + 32), ld_stream(row + 64), ld_stream(row + 96), ld_stream(row + 128),
                         ld_stream(row + 160)};
    const double X[3] = {ld_stream(row + 192), ld_stream(row + 224), ld_stream(row + 256)};
    const int pt = ld_stream(v.pt_c + i);
    const double2* vp = reinterpret_cast<const double2*>(v.Vinv + (size_t)pt * 6);
    const double2 v0 = vp[0], v1 = vp[1], v2_ = vp[2];
    const double vi[6] = {v0.x, v0.y, v1.x, v1.y, v2_.x, v2_.y};
    // T = Vinv A (columns), N = A T (symmetric 3x3)
    const double Ac0[3] = {A[0], A[1], A[2]}, Ac1[3] = {A[1], A[3], A[4]}, Ac2[3] = {A[2], A[4], A[5]};
    double T0[3], T1[3], T2[3];
    sym3_mul(vi, Ac0, T0);
    sym3_mul(vi, Ac1, T1);
    sym3_mul(vi, Ac2, T2);
    double N[3][3];
    {
      double c0[3], c1[3], c2[3];
      sym3_mul(A, T0, c0);
      sym3_mul(A, T1, c1);
      sym3_mul(A, T2, c2);
#pragma unroll
      for (int r = 0; r < 3; ++r) { N[r][0] = c0[r]; N[r][1] = c1[r]; N[r][2] = c2[r]; }
    }
    // P = 2 [X]x N  (rows: 2 X x N_col);  rr = -2 P [X]x;  rt = P;  tt = N
    double P[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double n0 = N[0][c], n1 = N[1][c], n2 = N[2][c];
      P[0][c] = 2.0 * (X[1] * n2 - X[2] * n1);
      P[1][c] = 2.0 * (X[2] * n0 - X[0] * n2);
      P[2][c] = 2.0 * (X[0] * n1 - X[1] * n0);
    }
    // (P [X]x)[r][c] = sum_k P[r][k] K[k][c],  K = [X]x = [[0,-X2,X1],[X2,0,-X0],[-X1,X0,0]]
    double PK[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      PK[r][0] = P[r][1] * X[2] - P[r][2] * X[1];
      PK[r][1] = -P[r][0] * X[2] + P[r][2] * X[0];
      PK[r][2] = P[r][0] * X[1] - P[r][1] * X[0];
    }
    RR[0] += -2.0 * PK[0][0]; RR[1] += -2.0 * PK[0][1]; RR[2] += -2.0 * PK[0][2];
    RR[3] += -2.0 * PK[1][1]; RR[4] += -2.0 * PK[1][2]; RR[5] += -2.0 * PK[2][2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) RT[3 * r + c] += P[r][c];
    TT[0] += N[0][0]; TT[1] += N[0][1]; TT[2] += N[0][2]; TT[3] += N[1][1]; TT[4] += N[1][2]; TT[5] += N[2][2];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) RR[k] = warp_sum(RR[k]);
#pragma unroll
  for (int k = 0; k < 9; ++k) RT[k] = warp_sum(RT[k]);
#pragma unroll
  for (int k = 0; k < 6; ++k) TT[k] = warp_sum(TT[k]);
  if (lane == 0) {
    const double4 q4c = ld_rec32(cam_rec + (size_t)cam * kCamRec);
    const double4 t4c = ld_rec32(cam_rec + (size_t)cam * kCamRec + 4);
    const int mask = (int)(__double_as_longlong(t4c.w) & 0xff);
    const double q[4] = {q4c.x, q4c.y, q4c.z, q4c.w};
    double R[9];
    quat_to_R(q, R);
    // full 6x6 in the world frame, then S = Rb M Rb^T
    double M[6][6];
    const double rr[3][3] = {{RR[0], RR[1], RR[2]}, {RR[1], RR[3], RR[4]}, {RR[2], RR[4], RR[5]}};
    const double tt[3][3] = {{TT[0], TT[1], TT[2]}, {TT[1], TT[3], TT[4]}, {TT[2], TT[4], TT[5]}};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        M[r][c] = rr[r][c];
        M[3 + r][3 + c] = tt[r][c];
        M[r][3 + c] = RT[3 * r + c];
        M[3 + c][r] = RT[3 * r + c];
      }
    double T[6][6];
    for (int blk = 0; blk < 2; ++blk)       // T = Rb M
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 6; ++c)
          T[3 * blk + r][c] = R[3 * r] * M[3 * blk][c] + R[3 * r + 1] * M[3 * blk + 1][c] + R[3 * r + 2] * M[3 * blk + 2][c];
    int idx = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c, ++idx) {
        const int cb = c / 3, cc = c % 3;     // S[r][c] = sum_k T[r][3cb+k] R[cc][k]
        double s = T[r][3 * cb] * R[3 * cc] + T[r][3 * cb + 1] * R[3 * cc + 1] + T[r][3 * cb + 2] * R[3 * cc + 2];
        const bool rfix = (r < 3) ? (mask & 1) : (mask & 2), cfix = (c < 3) ? (mask & 1) : (mask & 2);
        if (rfix || cfix) s = 0.0;
        if (s != 0.0) atomicAdd(&v.Sd[(size_t)cam * 21 + idx], s);
      }
  }
}

// ---------------------------------------------------------------------------
// pass A (point order):  s_p = [g_p] + sum_o A_o v_o,  v_o = x'_t - 2 X_p x x'_r ;  z_p = Vinv s_p
//   MODE 0: z -> z4[P][4]                      (mat-vec)
//   MODE 2: back-substitution epilogue (points_new, step scalars), as ba_schur_pass<2>
// ---------------------------------------------------------------------------
struct K3v2Smem {
  alignas(128) double At[kTile * kJpDoubles];
  double t[3][kTile + 1];
  double z[3][kTilePts + 1];
  double X[3][kTilePts + 1];
  unsigned pb[kTilePts + 1];
  double scratch[32];
  alignas(8) uint64_t mbar;
};

template <int MODE>
__global__ void __launch_bounds__(kTile, MODE == 0 ? B200_PA_MIN_CTAS : B200_K3_MIN_CTAS) ba2_pass_a(BAView v, BAViewV2 v2, const double* __restrict__ xp,
                                                                     const double* __restrict__ points,
                                                                     double* __restrict__ points_new, double radius,
                                                                     double* __restrict__ bscal,
                                                                     const PcgCtl* __restrict__ ctl) {
  extern __shared__ __align__(128) unsigned char smem_raw[];   // dynamic shared memory starts 128-B aligned (no static __shared__ in these kernels)
  if (ctl && ctl->done) return;   // the PCG stopping rule has fired: the queued iterations are no-ops
  K3v2Smem& sm = *reinterpret_cast<K3v2Smem*>(smem_raw);
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
      tma_load_1d_stream(sm.At, v2.Ap + (size_t)o0 * kJpDoubles, (uint32_t)nc0 * kRowBytes, &sm.mbar);
    }
  }
  // prefetch the first chunk: camera index -> R^T x record, local point index
  double xr0 = 0, xr1 = 0, xr2 = 0, xt0 = 0;
  double2 xt12 = make_double2(0, 0);
  int pl_pf = 0;
  if (tid < n) {
    const int cam = ld_stream(v.obs_cam + o0 + tid);
    pl_pf = ld_stream(v.obs_pt + o0 + tid) - p0;
    ld_nc_256(xp + (size_t)cam * kXqStride, xr0, xr1, xr2, xt0);
    xt12 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)cam * kXqStride + 4));
  }
  if (tid < npts) {
    sm.pb[tid] = v.pt_begin[p0 + tid];
    if (tid == npts - 1) sm.pb[npts] = o1;
#pragma unroll
    for (int k = 0; k < 3; ++k) sm.X[k][tid] = points[3 * (size_t)(p0 + tid) + k];
    sm.z[0][tid] = sm.z[1][tid] = sm.z[2][tid] = 0.0;
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * kTile;
    const int nc = min(kTile, n - c0);
    if (tid == 0 && ch > 0) {
      mbar_arrive_expect_tx(&sm.mbar, (uint32_t)nc * kRowBytes);
      tma_load_1d_stream(sm.At, v2.Ap + (size_t)(o0 + c0) * kJpDoubles, (uint32_t)nc * kRowBytes, &sm.mbar);
    }
    const bool active = tid < nc;
    if (active && ch > 0) {
      const int cam = ld_stream(v.obs_cam + o0 + c0 + tid);
      pl_pf = ld_stream(v.obs_pt + o0 + c0 + tid) - p0;
      ld_nc_256(xp + (size_t)cam * kXqStride, xr0, xr1, xr2, xt0);
      xt12 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)cam * kXqStride + 4));
    }
    mbar_wait(&sm.mbar, phase);
    phase ^= 1;
    double t0 = 0, t1 = 0, t2 = 0;
    if (active) {
      const double2* ar = reinterpret_cast<const double2*>(sm.At + tid * kJpDoubles);
      const double2 a0 = ar[0], a1 = ar[1], a2 = ar[2];
      const double X[3] = {sm.X[0][pl_pf], sm.X[1][pl_pf], sm.X[2][pl_pf]};
      const double xr[3] = {xr0, xr1, xr2}, xt[3] = {xt0, xt12.x, xt12.y};
      const double vv[3] = {xt[0] - 2.0 * (X[1] * xr[2] - X[2] * xr[1]), xt[1] - 2.0 * (X[2] * xr[0] - X[0] * xr[2]),
                            xt[2] - 2.0 * (X[0] * xr[1] - X[1] * xr[0])};
      t0 = a0.x * vv[0] + a0.y * vv[1] + a1.x * vv[2];
      t1 = a0.y * vv[0] + a1.y * vv[1] + a2.x * vv[2];
      t2 = a1.x * vv[0] + a2.x * vv[1] + a2.y * vv[2];
    }
    // per-point sums through shared memory: thread -> (point j, component k).  (A warp-shuffle segmented
    // reduction was measured slower: SHFL shares the LSU data pipe that bounds this kernel, profiles/r1_v2_sweep.md.)
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
      // 48-B rows are 16-B aligned: three 128-bit loads (8-B streaming loads would re-fetch the sector six times)
      const double2* vp = reinterpret_cast<const double2*>(v.Vinv + 6 * p);
      const double2 va = vp[0], vb = vp[1], vc = vp[2];
      const double vi[6] = {va.x, va.y, vb.x, vb.y, vc.x, vc.y};
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
          const double xo = sm.X[k][tid];
          points_new[3 * p + k] = xo + dp;
          b0 += g[k] * dp;
          b1 += Dp[k] * dp * dp;
          b2 += dp * dp;
          b3 += xo * xo;
        }
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 3; ++k) points_new[3 * p + k] = sm.X[k][tid];
    }
    if (MODE == 0) st_keep4(v2.z4 + 4 * p, make_double4(z[0], z[1], z[2], 0.0), l2_policy_evict_last());
  }
  if (MODE == 2) {
    // step scalars: one partial row per tile (bscal = part[n_tiles][4]), column-summed by ba_colsum afterwards --
    // deterministic, and no 4 x n_tiles same-address atomics
    b0 = warp_sum(b0);
    b1 = warp_sum(b1);
    b2 = warp_sum(b2);
    b3 = warp_sum(b3);
    if ((tid & 31) == 0) {
      sm.scratch[(tid >> 5)] = b0;
      sm.scratch[4 + (tid >> 5)] = b1;
      sm.scratch[8 + (tid >> 5)] = b2;
      sm.scratch[12 + (tid >> 5)] = b3;
    }
    __syncthreads();
    if (tid < 4) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < kTile / 32; ++w) a += sm.scratch[4 * tid + w];
      bscal[(size_t)tile * 4 + tid] = a;
    }
  }
}

// Column sums of the per-tile step scalars part[rows][4], stage 1: 4 partial sums per CTA (grid-stride over the
// rows, one 32-B load per row); stage 2 is ba_colsum over the gridDim.x partial rows.  Deterministic.
__global__ void __launch_bounds__(256) ba2_sum4_stage1(int rows, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double scratch[32];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const double2* row = reinterpret_cast<const double2*>(part + 4 * r);
    const double2 u = row[0], w = row[1];
    a0 += u.x; a1 += u.y; a2 += w.x; a3 += w.y;
  }
  a0 = block_sum(a0, scratch);
  a1 = block_sum(a1, scratch);
  a2 = block_sum(a2, scratch);
  a3 = block_sum(a3, scratch);
  if (threadIdx.x == 0) {
    double* o = out + 4 * (size_t)blockIdx.x;
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
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
// pass B (camera order): y_c -= [ 2 R sum (X x w) ; R sum w ],  w = A_o z_p
// ---------------------------------------------------------------------------
template <int NK>
__global__ void __launch_bounds__(128, B200_PB_MIN_CTAS) ba2_pass_b(BAView v, BAViewV2 v2, const double* __restrict__ cam_rec,
                                                 double* __restrict__ y, const PcgCtl* __restrict__ ctl, int seg_lo,
                                                 int seg_hi) {
  if (ctl && ctl->done) return;
  const int warp = seg_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);   // segments [seg_lo, seg_hi)
  const int lane = threadIdx.x & 31;
  if (warp >= seg_hi) return;
  const int cam = v.seg_cam[warp];
  const int b = v.seg_begin[warp], e = v.seg_end[warp];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  constexpr int NKK = NK > 0 ? NK : 1;
  double accK[NKK];
#pragma unroll
  for (int k = 0; k < NKK; ++k) accK[k] = 0.0;
  const double* rowB0 = NK > 0 ? v2.Bc + (size_t)(v.seg_row0[warp] >> 5) * (3 * NK * 32) + lane : nullptr;
  // two observations per lane and iteration: both index loads, then both gathers, are in flight together
  const double* row0 = v2.Ac + (size_t)(v.seg_row0[warp] >> 5) * (kJcDoubles * 32) + lane;
  const uint64_t keep = l2_policy_evict_last();
  // all point indices of the segment first (<= kSeg / 32 per lane): the z gathers then depend on nothing but
  // these registers, so every iteration costs one memory latency instead of two (index, then gather)
  int ptr[kSeg / 32];
#if B200_PB_PREFETCH
#pragma unroll
  for (int j = 0; j < kSeg / 32; ++j) {
    const int i = b + lane + 32 * j;
    ptr[j] = i < e ? ld_stream(v.pt_c + i) : -1;
  }
#endif
#pragma unroll
  for (int j = 0; j < kSeg / 32; j += 2) {
    if (b + 32 * j >= e) break;
#if !B200_PB_PREFETCH
    {
      const int i = b + lane + 32 * j;
      ptr[j] = i < e ? ld_stream(v.pt_c + i) : -1;
      ptr[j + 1] = i + 32 < e ? ld_stream(v.pt_c + i + 32) : -1;
    }
#endif
    const int pt0 = ptr[j], pt1 = ptr[j + 1];
    const bool ok0 = pt0 >= 0, ok1 = pt1 >= 0;
    const double* r0p = row0 + (size_t)j * (kJcDoubles * 32);
    const double* r1p = ok1 ? r0p + kJcDoubles * 32 : r0p;
    double a[kJcDoubles], c[kJcDoubles];
    double4 z0 = make_double4(0, 0, 0, 0), z1 = z0;
    if (ok0) {
#pragma unroll
      for (int k = 0; k < kJcDoubles; ++k) a[k] = ld_stream(r0p + 32 * k);
#pragma unroll
      for (int k = 0; k < kJcDoubles; ++k) c[k] = ld_stream(r1p + 32 * k);
      z0 = ld_keep4(v2.z4 + 4 * (size_t)pt0, keep);
      z1 = ld_keep4(v2.z4 + 4 * (size_t)(ok1 ? pt1 : pt0), keep);
    }
    if (NK > 0 && ok0) {   // y_k -= sum_o B_o^T z_p
      const double* b0p = rowB0 + (size_t)j * (3 * NK * 32);
      const double* b1p = ok1 ? b0p + 3 * NK * 32 : b0p;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        accK[k] += ld_stream(b0p + 32 * (3 * k)) * z0.x + ld_stream(b0p + 32 * (3 * k + 1)) * z0.y +
                   ld_stream(b0p + 32 * (3 * k + 2)) * z0.z;
        if (ok1)
          accK[k] += ld_stream(b1p + 32 * (3 * k)) * z1.x + ld_stream(b1p + 32 * (3 * k + 1)) * z1.y +
                     ld_stream(b1p + 32 * (3 * k + 2)) * z1.z;
      }
    }
    if (ok0) {
      const double w0 = a[0] * z0.x + a[1] * z0.y + a[2] * z0.z;
      const double w1 = a[1] * z0.x + a[3] * z0.y + a[4] * z0.z;
      const double w2 = a[2] * z0.x + a[4] * z0.y + a[5] * z0.z;
      const double X0 = a[6], X1 = a[7], X2 = a[8];
      acc[0] += 2.0 * (X1 * w2 - X2 * w1);
      acc[1] += 2.0 * (X2 * w0 - X0 * w2);
      acc[2] += 2.0 * (X0 * w1 - X1 * w0);
      acc[3] += w0;
      acc[4] += w1;
      acc[5] += w2;
    }
    if (ok1) {
      const double w0 = c[0] * z1.x + c[1] * z1.y + c[2] * z1.z;
      const double w1 = c[1] * z1.x + c[3] * z1.y + c[4] * z1.z;
      const double w2 = c[2] * z1.x + c[4] * z1.y + c[5] * z1.z;
      const double X0 = c[6], X1 = c[7], X2 = c[8];
      acc[0] += 2.0 * (X1 * w2 - X2 * w1);
      acc[1] += 2.0 * (X2 * w0 - X0 * w2);
      acc[2] += 2.0 * (X0 * w1 - X1 * w0);
      acc[3] += w0;
      acc[4] += w1;
      acc[5] += w2;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
  if (lane < 6) {
    const double4 q4c = ld_rec32(cam_rec + (size_t)cam * kCamRec);
    const double4 t4c = ld_rec32(cam_rec + (size_t)cam * kCamRec + 4);
    const int mask = (int)(__double_as_longlong(t4c.w) & 0xff);
    const double q[4] = {q4c.x, q4c.y, q4c.z, q4c.w};
    double R[9];
    quat_to_R(q, R);
    const int blk = lane / 3, r = lane % 3;
    const bool fixed = blk == 0 ? (mask & 1) : (mask & 2);
    const double s = R[3 * r] * acc[3 * blk] + R[3 * r + 1] * acc[3 * blk + 1] + R[3 * r + 2] * acc[3 * blk + 2];
    if (!fixed && s != 0.0) atomicAdd(&y[(size_t)cam * 6 + lane], -s);
  }
  if (NK > 0) {
    const size_t kb = (size_t)(v2.C + v.seg_intr[warp]);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const double sk = warp_sum(accK[k]);
      if (lane == 0 && sk != 0.0) atomicAdd(&y[kb * 6 + k], -sk);
    }
  }
}

// x_k of every image's intrinsics block into the two spare doubles of its xq row (pass A gathers ONE 64-B record)
__global__ void ba2k_pack_xk(int C, const double* __restrict__ cam_rec, const double* __restrict__ x, double* __restrict__ xp,
                             const PcgCtl* __restrict__ ctl) {
  if (ctl && ctl->done) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int blk = (int)(__double_as_longlong(cam_rec[(size_t)c * kCamRec + 7]) >> 8);
  xp[(size_t)c * kXqStride + 6] = x[(size_t)(C + blk) * 6];
  xp[(size_t)c * kXqStride + 7] = x[(size_t)(C + blk) * 6 + 1];
}

// frame x intrinsics coupling of U x:  y_f += U_fk x_k,  y_k += U_fk^T x_f   (one thread per image; the y_k sums of a
// CTA are combined in shared memory when the blocks fit, so a single shared camera costs one atomic per CTA and dof)
template <int NK>
__global__ void __launch_bounds__(256) ba2k_cross(int C, int K, const double* __restrict__ cam_rec, const double* __restrict__ Ufk,
                                                  const double* __restrict__ x, double* __restrict__ y,
                                                  const PcgCtl* __restrict__ ctl) {
  constexpr int kBins = 256;
  __shared__ double bins[kBins * NK];
  if (ctl && ctl->done) return;
  const bool use_bins = K <= kBins;
  if (use_bins) {
    for (int i = threadIdx.x; i < K * NK; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
  }
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const int blk = (int)(__double_as_longlong(cam_rec[(size_t)c * kCamRec + 7]) >> 8);
    double xf[6], xk[NK], tk[NK];
#pragma unroll
    for (int i = 0; i < 6; ++i) xf[i] = x[(size_t)c * 6 + i];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      xk[k] = x[(size_t)(C + blk) * 6 + k];
      tk[k] = 0.0;
    }
    const double* u = Ufk + (size_t)c * 6 * NK;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double yf = 0.0;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const double uik = u[i * NK + k];
        yf += uik * xk[k];
        tk[k] += uik * xf[i];
      }
      if (yf != 0.0) y[(size_t)c * 6 + i] += yf;   // this thread owns y_f of its image (pass B has finished: stream order)
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      if (tk[k] == 0.0) continue;
      if (use_bins) atomicAdd(&bins[blk * NK + k], tk[k]);
      else atomicAdd(&y[(size_t)(C + blk) * 6 + k], tk[k]);
    }
  }
  if (use_bins) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * NK; i += blockDim.x)
      if (bins[i] != 0.0) atomicAdd(&y[(size_t)(C + i / NK) * 6 + (i % NK)], bins[i]);
  }
}

}  // namespace b200
