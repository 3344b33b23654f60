/*
 * orc_gate.c -- CPU oracle of the Mahalanobis gate.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference contains no gating code (`grep -i mahal` over the reference finds nothing),
 * so this is the DEFINITION the CUDA kernel is checked against, built from the reference's point-landmark
 * measurement model: predicted measurement zhat_j = R_be (x_j - T) (kalmanClass.cpp:815-826), Jacobian rows
 * [-R_be | dR_be/dPsi (x_j - T) | ... R_be ...] (:836-851, with get_Rbe_derivs :1275-1298 as it is), noise
 * R_pts = PT_MEAS_COV^2 I (:975-984).  S_j = H_j P_sub H_j^T + R_pts over the 9 states (T, Psi, landmark j),
 * d2_ij = nu^T S_j^-1 nu, nu = z_i - zhat_j.  Arithmetic contract (bit-exact with the kernel): S_j^-1 by the
 * closed-form adjugate of the symmetrised 3x3, every expression evaluated as parenthesised with one rounding
 * per operator; the per-pair part uses only the packed (zhat, S^-1) values.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "orc_ekf.h"

#define PT_MEAS_COV 5.0

/* pack9[j] = {zhat0, zhat1, zhat2, i00, i01, i02, i11, i12, i22} from the landmark's 9x9 covariance sub-block
 * Ps over the states (T, Psi, landmark j) */
static void gate_pack_one(const double *x, int c, const double *R_be, double Rd[3][9], const double Ps[9][9],
                          double *pack)
{
  double xe[3], H[3][9], HP[3][9], S[3][3], det, idet;
  int a, b, k;
  for (a = 0; a < 3; ++a) xe[a] = x[c + a] - x[a];
  for (a = 0; a < 3; ++a) {
    double s = 0.0;
    for (k = 0; k < 3; ++k) s += R_be[3 * a + k] * xe[k];
    pack[a] = s;
    for (b = 0; b < 3; ++b) {
      double d = 0.0;
      H[a][b] = -R_be[3 * a + b];
      H[a][6 + b] = R_be[3 * a + b];
      for (k = 0; k < 3; ++k) d += Rd[b][3 * a + k] * xe[k];
      H[a][3 + b] = d;
    }
  }
  for (a = 0; a < 3; ++a)
    for (b = 0; b < 9; ++b) {
      double s = 0.0;
      for (k = 0; k < 9; ++k) s += H[a][k] * Ps[k][b];
      HP[a][b] = s;
    }
  for (a = 0; a < 3; ++a)
    for (b = 0; b < 3; ++b) {
      double s = 0.0;
      for (k = 0; k < 9; ++k) s += HP[a][k] * H[b][k];
      S[a][b] = s;
    }
  for (a = 0; a < 3; ++a) S[a][a] += PT_MEAS_COV * PT_MEAS_COV;
  { /* symmetrise, then adjugate inverse */
    const double s00 = S[0][0], s11 = S[1][1], s22 = S[2][2];
    const double s01 = 0.5 * (S[0][1] + S[1][0]), s02 = 0.5 * (S[0][2] + S[2][0]), s12 = 0.5 * (S[1][2] + S[2][1]);
    const double c00 = s11 * s22 - s12 * s12;
    const double c01 = s02 * s12 - s01 * s22;
    const double c02 = s01 * s12 - s02 * s11;
    const double c11 = s00 * s22 - s02 * s02;
    const double c12 = s01 * s02 - s00 * s12;
    const double c22 = s00 * s11 - s01 * s01;
    det = (s00 * c00 + s01 * c01) + s02 * c02;
    idet = 1.0 / det;
    pack[3] = c00 * idet;
    pack[4] = c01 * idet;
    pack[5] = c02 * idet;
    pack[6] = c11 * idet;
    pack[7] = c12 * idet;
    pack[8] = c22 * idet;
  }
}

void orc_gate_pack(const double *x, const double *P, int N, const int *point_col, int n_land, double *pack)
{
  double R_be[9], Rd[3][9];
  int j, a, b;
  orc_ekf_rbe(x[3], x[4], x[5], R_be);
  orc_ekf_rbe_derivs(x[3], x[4], x[5], Rd[0], Rd[1], Rd[2]);
  for (j = 0; j < n_land; ++j) {
    const int c = point_col[j];
    double Ps[9][9];
    int cols[9];
    for (a = 0; a < 6; ++a) cols[a] = a;
    for (a = 0; a < 3; ++a) cols[6 + a] = c + a;
    for (a = 0; a < 9; ++a)
      for (b = 0; b < 9; ++b) Ps[a][b] = P[(size_t)cols[a] * N + cols[b]];
    gate_pack_one(x, c, R_be, Rd, Ps, pack + 9 * j);
  }
}

/* The same pack from the pieces of P the gate reads, for maps whose dense P does not fit a host buffer (20k
 * landmarks: 28.8 GB): pose_rows = P[0..5][0..N-1] (6 x N), pose_cols = P[0..N-1][0..5] (N x 6), diag3[j] =
 * P[c_j..c_j+2][c_j..c_j+2] (3 x 3 per landmark). */
void orc_gate_pack_pieces(const double *x, const double *pose_rows, const double *pose_cols, const double *diag3, int N,
                          const int *point_col, int n_land, double *pack)
{
  double R_be[9], Rd[3][9];
  int j, a, b;
  orc_ekf_rbe(x[3], x[4], x[5], R_be);
  orc_ekf_rbe_derivs(x[3], x[4], x[5], Rd[0], Rd[1], Rd[2]);
  for (j = 0; j < n_land; ++j) {
    const int c = point_col[j];
    double Ps[9][9];
    for (a = 0; a < 6; ++a) {
      for (b = 0; b < 6; ++b) Ps[a][b] = pose_rows[(size_t)a * N + b];
      for (b = 0; b < 3; ++b) {
        Ps[a][6 + b] = pose_rows[(size_t)a * N + c + b];
        Ps[6 + b][a] = pose_cols[(size_t)(c + b) * 6 + a];
      }
    }
    for (a = 0; a < 3; ++a)
      for (b = 0; b < 3; ++b) Ps[6 + a][6 + b] = diag3[9 * (size_t)j + 3 * a + b];
    gate_pack_one(x, c, R_be, Rd, Ps, pack + 9 * j);
  }
}

typedef struct {
  const double *pack, *z;
  int n_land, nz, next;
  double gate;
  int *idx;
  double *d2;
} gate_job;

static void *gate_worker(void *arg)
{
  gate_job *J = (gate_job *)arg;
  for (;;) {
    const int lo = __atomic_fetch_add(&J->next, 64, __ATOMIC_RELAXED);
    int i, j;
    if (lo >= J->nz) break;
    for (i = lo; i < lo + 64 && i < J->nz; ++i) {
      const double z0 = J->z[3 * i], z1 = J->z[3 * i + 1], z2 = J->z[3 * i + 2];
      double best = J->gate;
      int bi = -1;
      for (j = 0; j < J->n_land; ++j) {
        const double *q = J->pack + 9 * j;
        const double n0 = z0 - q[0], n1 = z1 - q[1], n2 = z2 - q[2];
        const double t0 = (q[3] * n0 + q[4] * n1) + q[5] * n2;
        const double t1 = (q[4] * n0 + q[6] * n1) + q[7] * n2;
        const double t2 = (q[5] * n0 + q[7] * n1) + q[8] * n2;
        const double d2 = (n0 * t0 + n1 * t1) + n2 * t2;
        if (d2 < best) { best = d2; bi = j; } /* strict "<": ties keep the lowest j; NaN never passes */
      }
      J->idx[i] = bi;
      if (J->d2) J->d2[i] = (bi >= 0) ? best : INFINITY;
    }
  }
  return NULL;
}

static int gate_sweep(double *pack, int n_land, const double *z, int nz, double gate, int *idx_out, double *d2_out,
                      int nthreads)
{
  gate_job J;
  pthread_t th[256];
  int i, started = 0;
  J.pack = pack; J.z = z; J.n_land = n_land; J.nz = nz; J.next = 0; J.gate = gate; J.idx = idx_out; J.d2 = d2_out;
  if (nthreads > 256) nthreads = 256;
  for (i = 1; i < nthreads; ++i)
    if (pthread_create(&th[started], NULL, gate_worker, &J) == 0) ++started;
  gate_worker(&J);
  for (i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return 0;
}

int orc_gate_mahalanobis(const double *x, const double *P, int N, const int *point_col, int n_land,
                         const double *z, int nz, double gate, int *idx_out, double *d2_out, int nthreads)
{
  double *pack = (double *)malloc(sizeof(double) * 9 * (size_t)(n_land > 0 ? n_land : 1));
  if (!pack) return -1;
  orc_gate_pack(x, P, N, point_col, n_land, pack);
  gate_sweep(pack, n_land, z, nz, gate, idx_out, d2_out, nthreads);
  free(pack);
  return 0;
}

int orc_gate_mahalanobis_pieces(const double *x, const double *pose_rows, const double *pose_cols, const double *diag3,
                                int N, const int *point_col, int n_land, const double *z, int nz, double gate,
                                int *idx_out, double *d2_out, int nthreads)
{
  double *pack = (double *)malloc(sizeof(double) * 9 * (size_t)(n_land > 0 ? n_land : 1));
  if (!pack) return -1;
  orc_gate_pack_pieces(x, pose_rows, pose_cols, diag3, N, point_col, n_land, pack);
  gate_sweep(pack, n_land, z, nz, gate, idx_out, d2_out, nthreads);
  free(pack);
  return 0;
}
