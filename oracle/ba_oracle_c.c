/* ORACLE (test infrastructure, NOT product code) -- plain C / OpenMP restatement
 * of the per-iteration arithmetic of the reference's bundle adjustment
 * (glomap/estimators/bundle_adjustment.cc:115-190 residual blocks,
 * :192-242 Schur ordering points -> cameras, :99 ceres::Solve), used
 *   - by tests/ to cross-check oracle/ba_oracle.py at sizes numpy+splu cannot reach,
 *   - by bench.py as the `cpu_baseline` / `--impl reference` leg ("port": the
 *     reference itself needs Ceres + COLMAP, which cannot be built here).
 * PARITY UNPINNED: see oracle/ceres_lm.py.  The LM control flow lives in
 * oracle/ba_oracle_fast.py; this file holds the hot loops:
 *   ba_c_linearize  residuals, Huber corrector, analytic Jacobians, U, V, W, g
 *   ba_c_schur      S = (U + Dc) - W (V + Dp)^-1 W^T  (dense), b = -(gc - W Vinv gp)
 *   ba_c_backsub    dp = -Vinv (gp + W^T dc)
 *   ba_c_cost       1/2 sum rho
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Z_EPS 1e-12

static void quat_to_R(const double* q, double* R) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

static void intr_unpack(int model, const double* p, double* f) { /* fx fy cx cy k1 k2 */
  f[4] = f[5] = 0;
  if (model == 0) { f[0] = f[1] = p[0]; f[2] = p[1]; f[3] = p[2]; }
  else if (model == 1) { f[0] = p[0]; f[1] = p[1]; f[2] = p[2]; f[3] = p[3]; }
  else if (model == 2) { f[0] = f[1] = p[0]; f[2] = p[1]; f[3] = p[2]; f[4] = p[3]; }
  else { f[0] = f[1] = p[0]; f[2] = p[1]; f[3] = p[2]; f[4] = p[3]; f[5] = p[4]; }
}

static void huber(double s, double a, double* rho0, double* rho1) {
  double b = a * a;
  if (s > b) { double r = sqrt(s); *rho0 = 2 * a * r - b; *rho1 = a / r; }
  else { *rho0 = s; *rho1 = 1; }
}

/* residual (+ Jacobians) of one observation; returns 0 if behind the camera */
static int obs_eval(const double* R, const double* t, const double* f, const double* X, const double* xy, double a,
                    int mask, double* r, double* rho0, double* Jr, double* Jt, double* Jp) {
  double rx = R[0] * X[0] + R[1] * X[1] + R[2] * X[2];
  double ry = R[3] * X[0] + R[4] * X[1] + R[5] * X[2];
  double rz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2];
  double xc = rx + t[0], yc = ry + t[1], zc = rz + t[2];
  if (!(zc > Z_EPS)) return 0;
  double iz = 1 / zc, u = xc * iz, v = yc * iz, r2 = u * u + v * v;
  double d = 1 + r2 * (f[4] + f[5] * r2), dd = f[4] + 2 * f[5] * r2;
  double r0 = f[0] * u * d + f[2] - xy[0], r1 = f[1] * v * d + f[3] - xy[1];
  double rho1;
  huber(r0 * r0 + r1 * r1, a, rho0, &rho1);
  double w = sqrt(rho1);
  r[0] = w * r0; r[1] = w * r1;
  if (!Jr) return 1;
  double a00 = d + 2 * u * u * dd, a01 = 2 * u * v * dd, a11 = d + 2 * v * v * dd;
  double J[6] = {w * f[0] * a00 * iz, w * f[0] * a01 * iz, -w * f[0] * iz * (a00 * u + a01 * v),
                 w * f[1] * a01 * iz, w * f[1] * a11 * iz, -w * f[1] * iz * (a01 * u + a11 * v)};
  for (int k = 0; k < 6; ++k) Jt[k] = (mask & 2) ? 0 : J[k];
  for (int q = 0; q < 2; ++q) {
    double j0 = J[3 * q], j1 = J[3 * q + 1], j2 = J[3 * q + 2];
    Jr[3 * q + 0] = (mask & 1) ? 0 : -2 * (j1 * rz - j2 * ry);
    Jr[3 * q + 1] = (mask & 1) ? 0 : -2 * (j2 * rx - j0 * rz);
    Jr[3 * q + 2] = (mask & 1) ? 0 : -2 * (j0 * ry - j1 * rx);
    for (int b = 0; b < 3; ++b) Jp[3 * q + b] = j0 * R[b] + j1 * R[3 + b] + j2 * R[6 + b];
  }
  return 1;
}

/* U [C][36] dense 6x6, gc [C][6], V [P][9] dense 3x3, gp [P][3], W [N][18]; returns cost */
double ba_c_linearize(int C, int P, const int64_t* ptb, const int32_t* obs_cam, const double* obs_xy,
                      const int32_t* cam_intr, const int32_t* intr_model, const double* intr, const double* quat,
                      const double* trans, const double* points, const uint8_t* cam_mask, int min_views, double huber_a,
                      double* U, double* gc, double* V, double* gp, double* W) {
  double* Rm = (double*)malloc(sizeof(double) * 9 * C);
  double* F = (double*)malloc(sizeof(double) * 6 * C);
  for (int c = 0; c < C; ++c) {
    quat_to_R(quat + 4 * c, Rm + 9 * c);
    intr_unpack(intr_model[cam_intr[c]], intr + 12 * cam_intr[c], F + 6 * c);
  }
  memset(U, 0, sizeof(double) * 36 * C);
  memset(gc, 0, sizeof(double) * 6 * C);
  double cost = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cost)
  for (int p = 0; p < P; ++p) {
    double Vp[9] = {0}, g[3] = {0};
    int64_t b = ptb[p], e = ptb[p + 1];
    for (int64_t o = b; o < e; ++o) memset(W + 18 * o, 0, sizeof(double) * 18);
    if (e - b >= min_views) {
      for (int64_t o = b; o < e; ++o) {
        int c = obs_cam[o];
        double r[2], rho0, Jr[6], Jt[6], Jp[6];
        if (!obs_eval(Rm + 9 * c, trans + 3 * c, F + 6 * c, points + 3 * p, obs_xy + 2 * o, huber_a,
                      cam_mask ? cam_mask[c] : 0, r, &rho0, Jr, Jt, Jp))
          continue;
        cost += 0.5 * rho0;
        double Jc[2][6];
        for (int q = 0; q < 2; ++q)
          for (int k = 0; k < 3; ++k) { Jc[q][k] = Jr[3 * q + k]; Jc[q][3 + k] = Jt[3 * q + k]; }
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) Vp[3 * i + j] += Jp[i] * Jp[j] + Jp[3 + i] * Jp[3 + j];
          g[i] += Jp[i] * r[0] + Jp[3 + i] * r[1];
        }
        double* w = W + 18 * o;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) w[3 * i + j] = Jc[0][i] * Jp[j] + Jc[1][i] * Jp[3 + j];
        for (int i = 0; i < 6; ++i) {
          for (int j = 0; j < 6; ++j) {
            double val = Jc[0][i] * Jc[0][j] + Jc[1][i] * Jc[1][j];
#pragma omp atomic
            U[36 * (size_t)c + 6 * i + j] += val;
          }
          double gv = Jc[0][i] * r[0] + Jc[1][i] * r[1];
#pragma omp atomic
          gc[6 * (size_t)c + i] += gv;
        }
      }
    }
    memcpy(V + 9 * (size_t)p, Vp, sizeof(Vp));
    memcpy(gp + 3 * (size_t)p, g, sizeof(g));
  }
  free(Rm);
  free(F);
  return cost;
}

static int inv3(const double* a, double* inv) {
  double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[2] * a[7] - a[1] * a[8], c02 = a[1] * a[5] - a[2] * a[4];
  double det = a[0] * c00 + a[3] * c01 + a[6] * c02;
  if (!(fabs(det) > 0)) { memset(inv, 0, 72); return 0; }
  double id = 1 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) * id; inv[4] = (a[0] * a[8] - a[2] * a[6]) * id; inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) * id; inv[7] = (a[1] * a[6] - a[0] * a[7]) * id; inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return 1;
}

/* S [6C x 6C] row-major dense (full), b [6C], Vinv [P][9].  Dc [6C], Dp [3P] = LM damping. */
void ba_c_schur(int C, int P, const int64_t* ptb, const int32_t* obs_cam, int min_views, const double* U,
                const double* gc, const double* V, const double* gp, const double* W, const double* Dc,
                const double* Dp, double* S, double* b, double* Vinv) {
  const size_t n = (size_t)6 * C;
  memset(S, 0, sizeof(double) * n * n);
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) S[(6 * (size_t)c + i) * n + 6 * c + j] = U[36 * (size_t)c + 6 * i + j];
      S[(6 * (size_t)c + i) * n + 6 * c + i] += Dc[6 * c + i];
      b[6 * c + i] = -gc[6 * c + i];
    }
#pragma omp parallel for schedule(dynamic, 64)
  for (int p = 0; p < P; ++p) {
    int64_t bb = ptb[p], e = ptb[p + 1];
    double* vi = Vinv + 9 * (size_t)p;
    if (e - bb < min_views) { memset(vi, 0, 72); continue; }
    double Vd[9];
    memcpy(Vd, V + 9 * (size_t)p, 72);
    Vd[0] += Dp[3 * p]; Vd[4] += Dp[3 * p + 1]; Vd[8] += Dp[3 * p + 2];
    inv3(Vd, vi);
    double vg[3];
    for (int i = 0; i < 3; ++i) vg[i] = vi[3 * i] * gp[3 * p] + vi[3 * i + 1] * gp[3 * p + 1] + vi[3 * i + 2] * gp[3 * p + 2];
    for (int64_t o1 = bb; o1 < e; ++o1) {
      const double* w1 = W + 18 * o1;
      const int c1 = obs_cam[o1];
      double T[18]; /* W1 Vinv (6x3) */
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = w1[3 * i] * vi[j] + w1[3 * i + 1] * vi[3 + j] + w1[3 * i + 2] * vi[6 + j];
      for (int i = 0; i < 6; ++i) {
        double add = w1[3 * i] * vg[0] + w1[3 * i + 1] * vg[1] + w1[3 * i + 2] * vg[2];
#pragma omp atomic
        b[6 * c1 + i] += add;
      }
      for (int64_t o2 = bb; o2 < e; ++o2) {
        const double* w2 = W + 18 * o2;
        const int c2 = obs_cam[o2];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            double val = T[3 * i] * w2[3 * j] + T[3 * i + 1] * w2[3 * j + 1] + T[3 * i + 2] * w2[3 * j + 2];
#pragma omp atomic
            S[(6 * (size_t)c1 + i) * n + 6 * c2 + j] -= val;
          }
      }
    }
  }
}

/* dp = -Vinv (gp + W^T dc) */
void ba_c_backsub(int P, const int64_t* ptb, const int32_t* obs_cam, int min_views, const double* Vinv,
                  const double* gp, const double* W, const double* dc, double* dp) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int p = 0; p < P; ++p) {
    int64_t b = ptb[p], e = ptb[p + 1];
    double s[3] = {gp[3 * p], gp[3 * p + 1], gp[3 * p + 2]};
    if (e - b < min_views) { dp[3 * p] = dp[3 * p + 1] = dp[3 * p + 2] = 0; continue; }
    for (int64_t o = b; o < e; ++o) {
      const double* w = W + 18 * o;
      const double* x = dc + 6 * obs_cam[o];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) s[j] += w[3 * i + j] * x[i];
    }
    const double* vi = Vinv + 9 * (size_t)p;
    for (int i = 0; i < 3; ++i) dp[3 * p + i] = -(vi[3 * i] * s[0] + vi[3 * i + 1] * s[1] + vi[3 * i + 2] * s[2]);
  }
}

double ba_c_cost(int C, int P, const int64_t* ptb, const int32_t* obs_cam, const double* obs_xy,
                 const int32_t* cam_intr, const int32_t* intr_model, const double* intr, const double* quat,
                 const double* trans, const double* points, int min_views, double huber_a) {
  double* Rm = (double*)malloc(sizeof(double) * 9 * C);
  double* F = (double*)malloc(sizeof(double) * 6 * C);
  for (int c = 0; c < C; ++c) {
    quat_to_R(quat + 4 * c, Rm + 9 * c);
    intr_unpack(intr_model[cam_intr[c]], intr + 12 * cam_intr[c], F + 6 * c);
  }
  double cost = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cost)
  for (int p = 0; p < P; ++p) {
    int64_t b = ptb[p], e = ptb[p + 1];
    if (e - b < min_views) continue;
    for (int64_t o = b; o < e; ++o) {
      int c = obs_cam[o];
      double r[2], rho0;
      if (obs_eval(Rm + 9 * c, trans + 3 * c, F + 6 * c, points + 3 * p, obs_xy + 2 * o, huber_a, 0, r, &rho0, 0, 0, 0))
        cost += 0.5 * rho0;
    }
  }
  free(Rm);
  free(F);
  return cost;
}

int ba_c_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
