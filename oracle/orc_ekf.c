/*
 * orc_ekf.c -- CPU oracle: EKF predict / measurement update.  TEST INFRASTRUCTURE ONLY (see orc_ekf.h
 * for what is restated and how it is pinned).
 */
#include "orc_ekf.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* constants of kalmanClass.h:17-36 and common.h:114 */
#define MEAS_COV 0.5
#define PT_MEAS_COV 5.0
#define VX_COV 0.01
#define VY_COV 0.01
#define VZ_COV 0.5
#define WX_COV 0.01
#define WY_COV 0.01
#define WZ_COV 0.01
#define PHI_MEAS_COV 0.1
#define THETA_MEAS_COV 0.1
#define Z_MEAS_COV 0.2
#define DT 0.2
#define ORC_PI 3.1415926535

static double *mat(int r, int c) { return (double *)calloc((size_t)r * (size_t)c > 0 ? (size_t)r * c : 1, sizeof(double)); }

/* C (r x c) = A (r x k) * B (k x c), dot products accumulated in ascending k (C must not alias) */
static void gemm(const double *A, const double *B, double *C, int r, int k, int c)
{
  int i, j, l;
  for (i = 0; i < r; ++i) {
    double *Ci = C + (size_t)i * c;
    for (j = 0; j < c; ++j) Ci[j] = 0.0;
    for (l = 0; l < k; ++l) {
      const double a = A[(size_t)i * k + l];
      const double *Bl = B + (size_t)l * c;
      if (a == 0.0) continue; /* adding a*b = +-0 never changes a finite partial sum */
      for (j = 0; j < c; ++j) Ci[j] += a * Bl[j];
    }
  }
}

static void transpose(const double *A, double *At, int r, int c)
{
  int i, j;
  for (i = 0; i < r; ++i)
    for (j = 0; j < c; ++j) At[(size_t)j * r + i] = A[(size_t)i * c + j];
}

/* generate_Rbe, common.h:174-186 (row-major 3x3), including the reference's sign in entry [6] */
void orc_ekf_rbe(double phi, double theta, double psi, double *R)
{
  R[0] = cos(psi) * cos(theta);
  R[3] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  R[6] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
  R[1] = sin(psi) * cos(theta);
  R[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
  R[7] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  R[2] = -sin(theta);
  R[5] = cos(theta) * sin(phi);
  R[8] = cos(theta) * cos(phi);
}

/* get_Rbe_derivs, kalmanClass.cpp:1275-1298; matrices start zeroed (newMatrix :1202-1207).  The second
 * assignment to R_be_theta(1,1) and the missing (2,1) are the reference's. */
#define S_(M, i, j, v) (M)[3 * (i) + (j)] = (v)
void orc_ekf_rbe_derivs(double phi, double theta, double psi, double *Rp, double *Rt, double *Rs)
{
  memset(Rp, 0, 9 * sizeof(double));
  memset(Rt, 0, 9 * sizeof(double));
  memset(Rs, 0, 9 * sizeof(double));
  S_(Rp, 0, 0, 0.0);
  S_(Rp, 1, 0, cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi));
  S_(Rp, 2, 0, -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi));
  S_(Rp, 0, 1, 0.0);
  S_(Rp, 1, 1, sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi));
  S_(Rp, 2, 1, -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi));
  S_(Rp, 0, 2, 0.0);
  S_(Rp, 1, 2, cos(theta) * cos(phi));
  S_(Rp, 2, 2, -cos(theta) * sin(phi));

  S_(Rt, 0, 0, -cos(psi) * sin(theta));
  S_(Rt, 1, 0, cos(psi) * cos(theta) * sin(phi));
  S_(Rt, 2, 0, cos(psi) * cos(theta) * cos(phi));
  S_(Rt, 0, 1, -sin(psi) * sin(theta));
  S_(Rt, 1, 1, sin(psi) * cos(theta) * sin(phi));
  S_(Rt, 1, 1, sin(psi) * cos(theta) * cos(phi));
  S_(Rt, 0, 2, -cos(theta));
  S_(Rt, 1, 2, -sin(theta) * sin(phi));
  S_(Rt, 2, 2, -sin(theta) * cos(phi));

  S_(Rs, 0, 0, -sin(psi) * cos(theta));
  S_(Rs, 1, 0, -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi));
  S_(Rs, 2, 0, -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi));
  S_(Rs, 0, 1, cos(psi) * cos(theta));
  S_(Rs, 1, 1, cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi));
  S_(Rs, 2, 1, cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi));
  S_(Rs, 0, 2, 0.0);
  S_(Rs, 1, 2, 0.0);
  S_(Rs, 2, 2, 0.0);
}

/* GetSplitMatrices, kalmanClass.cpp:1144-1159 (de Casteljau split of a cubic at t) */
void orc_ekf_split_matrices(double t, double *A1, double *A2)
{
  static const double binom[4][4] = {{1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 1, 0}, {1, 3, 3, 1}};
  int i, j;
  memset(A1, 0, 16 * sizeof(double));
  memset(A2, 0, 16 * sizeof(double));
  for (i = 0; i < 4; ++i) {
    for (j = 0; j <= i; ++j) A1[i * 4 + j] = binom[i][j] * pow(t, j) * pow((1 - t), (i - j));
    for (j = i; j < 4; ++j) A2[4 * i + j] = binom[3 - i][j - i] * pow(t, (j - i)) * pow((1 - t), (3 - j));
  }
}

/* GetPredictedMeasurement, kalmanClass.cpp:1162-1199 */
void orc_ekf_predicted_curve(const double *x, const double *A, int col, double *zhat)
{
  const double Tx = x[0], Ty = x[1], psi = x[5];
  double t8[8], Rot[64];
  int i, j;
  for (i = 0; i < 4; ++i) {
    double ax = 0.0, ay = 0.0;
    for (j = 0; j < 4; ++j) {
      ax += A[4 * i + j] * x[col + j];
      ay += A[4 * i + j] * x[col + 4 + j];
    }
    t8[i] = ax - Tx;
    t8[i + 4] = ay - Ty;
  }
  memset(Rot, 0, sizeof Rot);
  for (i = 0; i < 4; ++i) {
    Rot[8 * i + i] = cos(psi);
    Rot[8 * (i + 4) + i + 4] = cos(psi);
    Rot[8 * (i + 4) + i] = -sin(psi);
    Rot[8 * i + i + 4] = sin(psi);
  }
  gemm(Rot, t8, zhat, 8, 8, 1);
}

/* ---- cvInvert(S, Sinv, CV_SVD): pseudo-inverse through a one-sided Jacobi SVD ----------------------
 * Works on the columns of W = A (n x n): rotations orthogonalise column pairs until convergence; then
 * A = U diag(w) V^T with U = normalised columns of W; A^+ = V diag(1/w) U^T, singular values below
 * DBL_EPSILON * 2 * sum(w) treated as zero (the rule of OpenCV's SVD back-substitution). */
int orc_pinv_svd(const double *A, double *Ainv, int n)
{
  double *W = mat(n, n), *V = mat(n, n), *w = mat(n, 1);
  int i, j, k, sweep, changed = 1;
  double sumw = 0.0, thr;
  if (!W || !V || !w) { free(W); free(V); free(w); return -1; }
  memcpy(W, A, sizeof(double) * (size_t)n * n);
  for (i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (sweep = 0; sweep < 60 && changed; ++sweep) {
    changed = 0;
    for (i = 0; i < n - 1; ++i)
      for (j = i + 1; j < n; ++j) {
        double a = 0.0, b = 0.0, p = 0.0, c, s, beta, gamma, delta;
        for (k = 0; k < n; ++k) {
          const double wi = W[(size_t)k * n + i], wj = W[(size_t)k * n + j];
          a += wi * wi; b += wj * wj; p += wi * wj;
        }
        if (fabs(p) <= DBL_EPSILON * sqrt(a * b)) continue;
        changed = 1;
        p *= 2;
        beta = a - b;
        gamma = hypot(p, beta);
        if (beta < 0) {
          delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        for (k = 0; k < n; ++k) {
          const double wi = W[(size_t)k * n + i], wj = W[(size_t)k * n + j];
          const double vi = V[(size_t)k * n + i], vj = V[(size_t)k * n + j];
          W[(size_t)k * n + i] = c * wi + s * wj;
          W[(size_t)k * n + j] = -s * wi + c * wj;
          V[(size_t)k * n + i] = c * vi + s * vj;
          V[(size_t)k * n + j] = -s * vi + c * vj;
        }
      }
  }
  for (i = 0; i < n; ++i) {
    double a = 0.0;
    for (k = 0; k < n; ++k) a += W[(size_t)k * n + i] * W[(size_t)k * n + i];
    w[i] = sqrt(a);
    sumw += w[i];
  }
  thr = DBL_EPSILON * 2 * sumw;
  /* Ainv[r][c] = sum_i V[r][i] * (1/w_i) * U[c][i],  U[c][i] = W[c][i]/w_i */
  for (i = 0; i < n * n; ++i) Ainv[i] = 0.0;
  for (i = 0; i < n; ++i) {
    double inv2;
    if (!(w[i] > thr)) continue;
    inv2 = 1.0 / (w[i] * w[i]);
    for (j = 0; j < n; ++j) {
      const double v = V[(size_t)j * n + i] * inv2;
      for (k = 0; k < n; ++k) Ainv[(size_t)j * n + k] += v * W[(size_t)k * n + i];
    }
  }
  free(W); free(V); free(w);
  return 0;
}

static void wrap_angle(double *a)
{
  while (*a > ORC_PI) *a -= 2 * ORC_PI;
  while (*a < -ORC_PI) *a += 2 * ORC_PI;
}

/* the shared tail of both updates, kalmanClass.cpp:863-926 / :1062-1106: K, state and covariance */
static int kalman_tail(double *x, double *P, int N, int M, const double *H, const double *R, const double *z,
                       const double *zhat, double *S_out, double *K_out)
{
  double *Ht = mat(N, M), *temp = mat(N, M), *S = mat(M, M), *Sinv = mat(M, M), *K0 = mat(N, M), *K = mat(N, M);
  double *Kt = mat(M, N), *Kt2 = mat(M, N), *innov = mat(M, 1), *delx = mat(N, 1);
  double *rowbuf;
  int i, j, rc = 0;
  if (!Ht || !temp || !S || !Sinv || !K0 || !K || !Kt || !Kt2 || !innov || !delx) { rc = -1; goto done; }
  transpose(H, Ht, M, N);
  gemm(P, Ht, temp, N, N, M);        /* temp = P Ht          */
  gemm(H, temp, S, M, N, M);         /* S = H temp           */
  for (i = 0; i < M * M; ++i) S[i] += R[i]; /* cvAdd(S,R,S) */
  if (orc_pinv_svd(S, Sinv, M) != 0) { rc = -1; goto done; }
  gemm(Ht, Sinv, K0, N, M, M);       /* K = Ht Sinv          */
  gemm(P, K0, K, N, N, M);           /* K = P K              */
  for (i = 0; i < M; ++i) innov[i] = z[i] - zhat[i];
  gemm(K, innov, delx, N, M, 1);
  for (i = 0; i < N; ++i) x[i] += delx[i];
  wrap_angle(&x[3]);
  wrap_angle(&x[4]);
  wrap_angle(&x[5]);
  transpose(K, Kt, N, M);
  gemm(S, Kt, Kt2, M, M, N);         /* Kt2 = S Kt           */
  /* delP = K Kt2 ; P -= delP, one row at a time to bound memory */
  rowbuf = mat(1, N);
  if (!rowbuf) { rc = -1; goto done; }
  for (i = 0; i < N; ++i) {
    gemm(K + (size_t)i * M, Kt2, rowbuf, 1, M, N);
    for (j = 0; j < N; ++j) P[(size_t)i * N + j] -= rowbuf[j];
  }
  free(rowbuf);
  /* P = 0.5 P + 0.5 P^T (cvAddWeighted) */
  for (i = 0; i < N; ++i)
    for (j = i; j < N; ++j) {
      const double a = P[(size_t)i * N + j], b = P[(size_t)j * N + i];
      P[(size_t)i * N + j] = a * 0.5 + b * 0.5;
      P[(size_t)j * N + i] = b * 0.5 + a * 0.5;
    }
  if (S_out) memcpy(S_out, S, sizeof(double) * (size_t)M * M);
  if (K_out) memcpy(K_out, K, sizeof(double) * (size_t)N * M);
done:
  free(Ht); free(temp); free(S); free(Sinv); free(K0); free(K); free(Kt); free(Kt2); free(innov); free(delx);
  return rc;
}

/* rows of H and predicted measurements of the point landmarks, kalmanClass.cpp:815-851 / :1016-1050 */
static void point_rows(const double *x, int N, double *H, double *zhat, int row0, const int *point_col, int n_pts)
{
  double R_be[9], Rd[3][9], xe[3], xb[3], t3[3];
  int k, i, j;
  orc_ekf_rbe(x[3], x[4], x[5], R_be);
  orc_ekf_rbe_derivs(x[3], x[4], x[5], Rd[0], Rd[1], Rd[2]);
  for (k = 0; k < n_pts; ++k) {
    const int c = point_col[k];
    for (i = 0; i < 3; ++i) xe[i] = x[c + i] - x[i];
    gemm(R_be, xe, xb, 3, 3, 1);
    for (i = 0; i < 3; ++i) zhat[row0 + 3 * k + i] = xb[i];
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) {
        H[(size_t)(row0 + 3 * k + i) * N + j] = -R_be[3 * i + j];
        H[(size_t)(row0 + 3 * k + i) * N + c + j] = R_be[3 * i + j];
      }
    for (j = 0; j < 3; ++j) {
      gemm(Rd[j], xe, t3, 3, 3, 1);
      for (i = 0; i < 3; ++i) H[(size_t)(row0 + 3 * k + i) * N + 3 + j] = t3[i];
    }
  }
}

int orc_ekf_update_curves_points(double *x, double *P, int N, const double *zc, int n, const double *A,
                                 const int *curve_col, const double *point_meas, const int *point_col,
                                 int n_pts, double *H_out, double *S_out, double *K_out)
{
  const int cms = n * 8 + 3, pms = n_pts * 3, M = cms + pms;
  double *z = mat(M, 1), *H = mat(M, N), *R = mat(M, M), *zhat = mat(M, 1);
  double Rot[64], Rotderiv[64], temp8[64], temp8b[64], t81[8], t81b[8];
  const double Tx = x[0], Ty = x[1], psi = x[5];
  int i, j, k, rc;
  if (!z || !H || !R || !zhat) { free(z); free(H); free(R); free(zhat); return -1; }
  for (i = 0; i < cms; ++i) z[i] = zc[i];
  for (i = 0; i < pms; ++i) z[cms + i] = point_meas[i];
  for (i = 0; i < cms; ++i) R[(size_t)i * M + i] = pow(MEAS_COV, 2.0);
  R[(size_t)(n * 8) * M + n * 8] = pow(Z_MEAS_COV, 2.0);
  R[(size_t)(n * 8 + 1) * M + n * 8 + 1] = pow(PHI_MEAS_COV, 2.0);
  R[(size_t)(n * 8 + 2) * M + n * 8 + 2] = pow(THETA_MEAS_COV, 2.0);
  for (i = cms; i < M; ++i) R[(size_t)i * M + i] = pow(PT_MEAS_COV, 2.0);

  memset(Rot, 0, sizeof Rot);
  memset(Rotderiv, 0, sizeof Rotderiv);
  for (i = 0; i < 4; ++i) {
    Rot[8 * i + i] = cos(psi);
    Rot[8 * (i + 4) + i + 4] = cos(psi);
    Rot[8 * (i + 4) + i] = -sin(psi);
    Rot[8 * i + i + 4] = sin(psi);
    Rotderiv[8 * i + i] = -sin(psi);
    Rotderiv[8 * (i + 4) + i + 4] = -sin(psi);
    Rotderiv[8 * (i + 4) + i] = -cos(psi);
    Rotderiv[8 * i + i + 4] = cos(psi);
  }
  for (k = 0; k < n; ++k) {
    const int c = curve_col[k];
    const double *Ak = A + 16 * k;
    memset(temp8, 0, sizeof temp8);
    for (i = 0; i < 4; ++i) {
      for (j = 0; j < 4; ++j) {
        temp8[8 * i + j] = Ak[4 * i + j];
        temp8[8 * (i + 4) + j + 4] = Ak[4 * i + j];
      }
      t81[i] = x[c + i];
      t81[i + 4] = x[c + 4 + i];
    }
    gemm(temp8, t81, t81b, 8, 8, 1);
    for (i = 0; i < 4; ++i) t81b[i] -= Tx;
    for (i = 4; i < 8; ++i) t81b[i] -= Ty;
    gemm(Rotderiv, t81b, t81, 8, 8, 1);
    gemm(Rot, temp8, temp8b, 8, 8, 8);
    for (i = 0; i < 8; ++i)
      for (j = 0; j < 8; ++j) H[(size_t)(i + k * 8) * N + c + j] = temp8b[8 * i + j];
    for (i = 0; i < 4; ++i) {
      H[(size_t)(i + k * 8) * N + 0] = -cos(psi);
      H[(size_t)(i + k * 8) * N + 1] = -sin(psi);
      H[(size_t)(i + k * 8) * N + 5] = t81[i];
    }
    for (i = 4; i < 8; ++i) {
      H[(size_t)(i + k * 8) * N + 0] = sin(psi);
      H[(size_t)(i + k * 8) * N + 1] = -cos(psi);
      H[(size_t)(i + k * 8) * N + 5] = t81[i];
    }
  }
  H[(size_t)(n * 8) * N + 2] = 1.0;
  H[(size_t)(n * 8 + 1) * N + 3] = 1.0;
  H[(size_t)(n * 8 + 2) * N + 4] = 1.0;
  point_rows(x, N, H, zhat, cms, point_col, n_pts);
  for (k = 0; k < n; ++k) orc_ekf_predicted_curve(x, A + 16 * k, curve_col[k], zhat + 8 * k);
  zhat[n * 8] = x[2];
  zhat[n * 8 + 1] = x[3];
  zhat[n * 8 + 2] = x[4];
  if (H_out) memcpy(H_out, H, sizeof(double) * (size_t)M * N);
  rc = kalman_tail(x, P, N, M, H, R, z, zhat, S_out, K_out);
  free(z); free(H); free(R); free(zhat);
  return rc;
}

int orc_ekf_update_points(double *x, double *P, int N, const double *point_meas, const int *point_col, int n_pts)
{
  const int M = 3 * n_pts;
  double *H = mat(M, N), *R = mat(M, M), *zhat = mat(M, 1);
  int i, rc;
  if (!H || !R || !zhat) { free(H); free(R); free(zhat); return -1; }
  for (i = 0; i < M; ++i) R[(size_t)(M + 1) * i] = pow(PT_MEAS_COV, 2.0);
  point_rows(x, N, H, zhat, 0, point_col, n_pts);
  rc = kalman_tail(x, P, N, M, H, R, point_meas, zhat, NULL, NULL);
  free(H); free(R); free(zhat);
  return rc;
}

/* PredictKF, kalmanClass.cpp:161-272 */
void orc_ekf_predict(double *x, double *P, int N)
{
  const int RS = ORC_ROBOT_STATE, NI = N - RS;
  double Reb[9], Rebt[9], t33[9], t33b[9], Q[144], F[144], Ft[144], Prr[144], Prr2[144];
  double *Pri = mat(RS, NI > 0 ? NI : 1), *Pri2 = mat(RS, NI > 0 ? NI : 1);
  const double phi = x[3], theta = x[4], psi = x[5];
  int i, j;
  memset(F, 0, sizeof F);
  for (i = 0; i < RS; ++i) F[(RS + 1) * i] = 1.0;
  for (i = 0; i < 6; ++i) F[RS * i + (i + 6)] = DT;
  Reb[0] = cos(psi) * cos(theta);
  Reb[1] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  Reb[2] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
  Reb[3] = sin(psi) * cos(theta);
  Reb[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
  Reb[5] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  Reb[6] = -sin(theta);
  Reb[7] = cos(theta) * sin(phi);
  Reb[8] = cos(theta) * cos(phi);
  transpose(Reb, Rebt, 3, 3);
  memset(Q, 0, sizeof Q);
  memset(t33, 0, sizeof t33);
  t33[0] = pow(VX_COV, 2.0);
  t33[4] = pow(VY_COV, 2.0);
  t33[8] = pow(VZ_COV, 2.0);
  gemm(Reb, t33, t33b, 3, 3, 3);
  gemm(t33b, Rebt, t33, 3, 3, 3);
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      Q[RS * i + j] = pow(DT, 3.0) * t33[3 * i + j];
      Q[RS * i + j + 6] = pow(DT, 2.0) * t33[3 * i + j];
      Q[RS * (j + 6) + i] = pow(DT, 2.0) * t33[3 * i + j];
      Q[RS * (i + 6) + j + 6] = pow(DT, 1.0) * t33[3 * i + j];
    }
  memset(t33, 0, sizeof t33);
  t33[0] = pow(WX_COV, 2.0);
  t33[4] = pow(WY_COV, 2.0);
  t33[8] = pow(WZ_COV, 2.0);
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      Q[RS * (i + 3) + j + 3] = pow(DT, 3.0) * t33[3 * i + j];
      Q[RS * (i + 3) + j + 9] = pow(DT, 2.0) * t33[3 * i + j];
      Q[RS * (j + 9) + i + 3] = pow(DT, 2.0) * t33[3 * i + j];
      Q[RS * (i + 9) + j + 9] = pow(DT, 1.0) * t33[3 * i + j];
    }
  for (i = 0; i < RS; ++i) {
    for (j = 0; j < RS; ++j) Prr[RS * i + j] = P[(size_t)i * N + j];
    for (j = RS; j < N; ++j) Pri[(size_t)i * NI + j - RS] = P[(size_t)i * N + j];
  }
  if (NI > 0) gemm(F, Pri, Pri2, RS, RS, NI);
  gemm(F, Prr, Prr2, RS, RS, RS);
  transpose(F, Ft, RS, RS);
  gemm(Prr2, Ft, Prr, RS, RS, RS);
  for (i = 0; i < RS * RS; ++i) Prr[i] += Q[i];
  for (i = 0; i < RS; ++i) {
    for (j = 0; j < RS; ++j) P[(size_t)i * N + j] = Prr[RS * i + j];
    for (j = RS; j < N; ++j) {
      P[(size_t)i * N + j] = Pri2[(size_t)i * NI + j - RS];
      P[(size_t)j * N + i] = Pri2[(size_t)i * NI + j - RS];
    }
  }
  x[0] += x[6] * DT;
  x[1] += x[7] * DT;
  x[2] += x[8] * DT;
  x[3] += x[9] * DT;
  x[4] += x[10] * DT;
  x[5] += x[11] * DT;
  free(Pri);
  free(Pri2);
}

/* ---- state augmentation ("next" rows of the scope table): AddFirstStates (kalmanClass.cpp:275-311),
 * AddNewCurve (:314-499), AddNewPoints (:502-666).  The reference re-allocates and copies all of P per call;
 * here x and P are caller-owned with capacity for the grown state, row-major with leading dimension `ld`.
 * Reference quirks kept: R1 = MEAS_COV*I and R_pts = PT_MEAS_COV*I are NOT squared here (constructor :91-98,
 * unlike the update's R); points added in one call get no cross-covariance among themselves (the cross block is
 * built from the covariance as it was before the call); AddNewCurve uses Gx*Prr^T, AddNewPoints Gx*Prr. */
#define VREAL 1.0

void orc_ekf_add_first_states(double *x, double *P, int ld, const double *z19)
{
  static const double pose_diag[12] = {0.0000001, 0.0000001, 0.1, 0.1, 0.1, 0.0000001, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1};
  int i, j;
  for (i = 0; i < 28; ++i) {
    x[i] = 0.0;
    for (j = 0; j < 28; ++j) P[(size_t)i * ld + j] = 0.0;
  }
  for (i = 0; i < 16; ++i) {
    P[(size_t)(i + 12) * ld + i + 12] = pow(MEAS_COV, 2.0);
    x[i + 12] = z19[i];
  }
  x[2] = z19[16];
  x[3] = z19[17];
  x[4] = z19[18];
  x[6] = VREAL;
  for (i = 0; i < 12; ++i) P[(size_t)i * ld + i] = pose_diag[i];
}

/* returns the new state dimension N + 8 */
int orc_ekf_add_new_curve(double *x, double *P, int ld, int N, const double *z8, const double *A)
{
  const double Tx = x[0], Ty = x[1], psi = x[5];
  double Rot[64], Rotd[64], Ainv[16], Hinv[64], t81[8], t81b[8], xnew[8], ones[4] = {1, 1, 1, 1}, a1[4];
  double Gz[64], Gx[8 * 12], Prr[144], PrrT[144], GxT[12 * 8], tmp[12 * 8], PN1N1[64], GzT[64], R1[64], t8[64], t8b[64];
  double PN1r[8 * 12];
  int i, j, k;
  memset(Rot, 0, sizeof Rot);
  memset(Rotd, 0, sizeof Rotd);
  for (i = 0; i < 4; ++i) {
    Rot[8 * i + i] = cos(psi);
    Rot[8 * (i + 4) + i + 4] = cos(psi);
    Rot[8 * (i + 4) + i] = sin(psi);
    Rot[8 * i + i + 4] = -sin(psi);
    Rotd[8 * i + i] = -sin(psi);
    Rotd[8 * (i + 4) + i + 4] = -sin(psi);
    Rotd[8 * (i + 4) + i] = cos(psi);
    Rotd[8 * i + i + 4] = -cos(psi);
  }
  orc_pinv_svd(A, Ainv, 4);
  memset(Hinv, 0, sizeof Hinv);
  for (i = 0; i < 4; ++i)
    for (j = 0; j < 4; ++j) {
      Hinv[8 * i + j] = Ainv[4 * i + j];
      Hinv[8 * (i + 4) + j + 4] = Ainv[4 * i + j];
    }
  gemm(Rot, z8, t81, 8, 8, 1);
  for (i = 0; i < 4; ++i) t81[i] += Tx;
  for (i = 4; i < 8; ++i) t81[i] += Ty;
  gemm(Hinv, t81, xnew, 8, 8, 1);
  gemm(Ainv, ones, a1, 4, 4, 1);
  gemm(Hinv, Rot, Gz, 8, 8, 8);
  gemm(Rotd, z8, t81, 8, 8, 1);
  gemm(Hinv, t81, t81b, 8, 8, 1);
  memset(Gx, 0, sizeof Gx);
  for (i = 0; i < 4; ++i) {
    Gx[12 * i + 0] = a1[i];
    Gx[12 * (i + 4) + 1] = a1[i];
    Gx[12 * i + 5] = t81b[i];
    Gx[12 * (i + 4) + 5] = t81b[i + 4];
  }
  for (i = 0; i < 12; ++i)
    for (j = 0; j < 12; ++j) Prr[12 * i + j] = P[(size_t)i * ld + j];
  transpose(Gx, GxT, 8, 12);
  gemm(Prr, GxT, tmp, 12, 12, 8);
  gemm(Gx, tmp, PN1N1, 8, 12, 8);
  memset(R1, 0, sizeof R1);
  for (i = 0; i < 8; ++i) R1[9 * i] = MEAS_COV;
  transpose(Gz, GzT, 8, 8);
  gemm(R1, GzT, t8, 8, 8, 8);
  gemm(Gz, t8, t8b, 8, 8, 8);
  for (i = 0; i < 64; ++i) PN1N1[i] += t8b[i];
  transpose(Prr, PrrT, 12, 12);
  gemm(Gx, PrrT, PN1r, 8, 12, 12);
  for (i = 0; i < 8; ++i) {
    x[N + i] = xnew[i];
    for (j = 0; j < 8; ++j) P[(size_t)(N + i) * ld + N + j] = PN1N1[8 * i + j];
    for (j = 12; j < N; ++j) { /* PN1i = Gx * Pri */
      double s = 0.0;
      for (k = 0; k < 12; ++k) s += Gx[12 * i + k] * P[(size_t)k * ld + j];
      P[(size_t)(N + i) * ld + j] = s;
      P[(size_t)j * ld + N + i] = s;
    }
    for (j = 0; j < 12; ++j) {
      P[(size_t)(N + i) * ld + j] = PN1r[12 * i + j];
      P[(size_t)j * ld + N + i] = PN1r[12 * i + j];
    }
  }
  return N + 8;
}

static void reb(double phi, double theta, double psi, double *R)
{ /* generate_Reb, common.h:189-201 */
  R[0] = cos(psi) * cos(theta);
  R[1] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi);
  R[2] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi);
  R[3] = sin(psi) * cos(theta);
  R[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi);
  R[5] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi);
  R[6] = -sin(theta);
  R[7] = cos(theta) * sin(phi);
  R[8] = cos(theta) * cos(phi);
}

static void reb_derivs(double phi, double theta, double psi, double *Rp, double *Rt, double *Rs)
{ /* get_Reb_derivs, kalmanClass.cpp:1230-1262 */
  S_(Rp, 0, 0, 0.0);
  S_(Rp, 0, 1, cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi));
  S_(Rp, 0, 2, -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi));
  S_(Rp, 1, 0, 0.0);
  S_(Rp, 1, 1, sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi));
  S_(Rp, 1, 2, -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi));
  S_(Rp, 2, 0, 0.0);
  S_(Rp, 2, 1, cos(theta) * cos(phi));
  S_(Rp, 2, 2, -cos(theta) * sin(phi));
  S_(Rt, 0, 0, -cos(psi) * sin(theta));
  S_(Rt, 0, 1, cos(psi) * cos(theta) * sin(phi));
  S_(Rt, 0, 2, cos(psi) * cos(theta) * cos(phi));
  S_(Rt, 1, 0, -sin(psi) * sin(theta));
  S_(Rt, 1, 1, sin(psi) * cos(theta) * sin(phi));
  S_(Rt, 1, 2, sin(psi) * cos(theta) * cos(phi));
  S_(Rt, 2, 0, -cos(theta));
  S_(Rt, 2, 1, -sin(theta) * sin(phi));
  S_(Rt, 2, 2, -sin(theta) * cos(phi));
  S_(Rs, 0, 0, -sin(psi) * cos(theta));
  S_(Rs, 0, 1, -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi));
  S_(Rs, 0, 2, -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi));
  S_(Rs, 1, 0, cos(psi) * cos(theta));
  S_(Rs, 1, 1, cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi));
  S_(Rs, 1, 2, cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi));
  S_(Rs, 2, 0, 0.0);
  S_(Rs, 2, 1, 0.0);
  S_(Rs, 2, 2, 0.0);
}

/* returns the new state dimension N + 3*n_pts */
int orc_ekf_add_new_points(double *x, double *P, int ld, int N, const double *meas, int n_pts)
{
  double R_eb[9], Rd[3][9], Prr[144];
  int n, i, j, k;
  if (n_pts <= 0) return N;
  reb(x[3], x[4], x[5], R_eb);
  reb_derivs(x[3], x[4], x[5], Rd[0], Rd[1], Rd[2]);
  for (i = 0; i < 12; ++i)
    for (j = 0; j < 12; ++j) Prr[12 * i + j] = P[(size_t)i * ld + j];
  for (n = 0; n < n_pts; ++n) {
    const double *xb = meas + 3 * n;
    const int o = N + 3 * n;
    double t3[3], Gx[36], GxT[36], tmp[36], PN1N1[9], Rp[9], GzT[9], t33[9], t33b[9], PN1r[36];
    gemm(R_eb, xb, t3, 3, 3, 1);
    for (j = 0; j < 3; ++j) x[o + j] = t3[j] + x[j];
    memset(Gx, 0, sizeof Gx);
    Gx[0] = 1.0; Gx[12 + 1] = 1.0; Gx[24 + 2] = 1.0;
    for (j = 0; j < 3; ++j) {
      gemm(Rd[j], xb, t3, 3, 3, 1);
      for (i = 0; i < 3; ++i) Gx[12 * i + j + 3] = t3[i];
    }
    transpose(Gx, GxT, 3, 12);
    gemm(Prr, GxT, tmp, 12, 12, 3);
    gemm(Gx, tmp, PN1N1, 3, 12, 3);
    memset(Rp, 0, sizeof Rp);
    for (i = 0; i < 3; ++i) Rp[4 * i] = PT_MEAS_COV;
    transpose(R_eb, GzT, 3, 3);
    gemm(Rp, GzT, t33, 3, 3, 3);
    gemm(R_eb, t33, t33b, 3, 3, 3);
    for (i = 0; i < 9; ++i) PN1N1[i] += t33b[i];
    gemm(Gx, Prr, PN1r, 3, 12, 12);
    for (i = 0; i < 3; ++i) {
      for (j = 0; j < 3; ++j) P[(size_t)(o + i) * ld + o + j] = PN1N1[3 * i + j];
      for (j = 12; j < N; ++j) { /* existing states as they were before the call */
        double s = 0.0;
        for (k = 0; k < 12; ++k) s += Gx[12 * i + k] * P[(size_t)k * ld + j];
        P[(size_t)(o + i) * ld + j] = s;
        P[(size_t)j * ld + o + i] = s;
      }
      for (j = 0; j < 12; ++j) {
        P[(size_t)(o + i) * ld + j] = PN1r[12 * i + j];
        P[(size_t)j * ld + o + i] = PN1r[12 * i + j];
      }
    }
  }
  return N + 3 * n_pts;
}
