/*
 * orc_kat.c -- known-answer problems for the LM oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The seven unconstrained problems the reference's own demo harness runs without LAPACK
 * (levmar-2.5/lmdemo.c:50-230 define them, :856-960 give the starting points, data and which LM
 * variant each one uses; expected outputs are tabulated in SURVEY.md A.5).  They are the classic
 * More'-Garbow-Hillstrom test functions, restated here with the same expression shapes so that the
 * iteration counts of the table are reproducible; tests/test_oracle_lm.py runs them through
 * orc_lm_der/orc_lm_dif AND through the reference levmar built in oracle/_ref and demands identical
 * bits from both.
 */
#include <math.h>
#include <string.h>

#include "orc_lm.h"

#define KAT_PI 3.14159265358979323846

static void f_ros(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; ++i)
    hx[i] = ((1.0 - p[0]) * (1.0 - p[0]) + 105.0 * (p[1] - p[0] * p[0]) * (p[1] - p[0] * p[0]));
}
static void j_ros(double *p, double *J, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; ++i) {
    J[2 * i] = (-2 + 2 * p[0] - 4 * 105.0 * (p[1] - p[0] * p[0]) * p[0]);
    J[2 * i + 1] = (2 * 105.0 * (p[1] - p[0] * p[0]));
  }
}

static void f_modros(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; i += 3) {
    hx[i] = 10 * (p[1] - p[0] * p[0]);
    hx[i + 1] = 1.0 - p[0];
    hx[i + 2] = 1E02;
  }
}
static void j_modros(double *p, double *J, int m, int n, void *a)
{
  int i, j = 0; (void)m; (void)a;
  for (i = 0; i < n; i += 3) {
    J[j++] = -20.0 * p[0]; J[j++] = 10.0;
    J[j++] = -1.0;         J[j++] = 0.0;
    J[j++] = 0.0;          J[j++] = 0.0;
  }
}

static void f_powell(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; i += 2) {
    hx[i] = p[0];
    hx[i + 1] = 10.0 * p[0] / (p[0] + 0.1) + 2 * p[1] * p[1];
  }
}
static void j_powell(double *p, double *J, int m, int n, void *a)
{
  int i, j = 0; (void)m; (void)a;
  for (i = 0; i < n; i += 2) {
    J[j++] = 1.0; J[j++] = 0.0;
    J[j++] = 1.0 / ((p[0] + 0.1) * (p[0] + 0.1));
    J[j++] = 4.0 * p[1];
  }
}

static void f_wood(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; i += 6) {
    hx[i] = 10.0 * (p[1] - p[0] * p[0]);
    hx[i + 1] = 1.0 - p[0];
    hx[i + 2] = sqrt(90.0) * (p[3] - p[2] * p[2]);
    hx[i + 3] = 1.0 - p[2];
    hx[i + 4] = sqrt(10.0) * (p[1] + p[3] - 2.0);
    hx[i + 5] = (p[1] - p[3]) / sqrt(10.0);
  }
}

static void f_meyer(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; ++i) {
    const double ui = 0.45 + 0.05 * i;
    hx[i] = p[0] * exp(10.0 * p[1] / (ui + p[2]) - 13.0);
  }
}

static void f_osborne(double *p, double *hx, int m, int n, void *a)
{
  int i; (void)m; (void)a;
  for (i = 0; i < n; ++i) {
    const double t = 10 * i;
    hx[i] = p[0] + p[1] * exp(-p[3] * t) + p[2] * exp(-p[4] * t);
  }
}
static void j_osborne(double *p, double *J, int m, int n, void *a)
{
  int i, j = 0; (void)m; (void)a;
  for (i = 0; i < n; ++i) {
    const double t = 10 * i;
    const double e1 = exp(-p[3] * t), e2 = exp(-p[4] * t);
    J[j++] = 1.0; J[j++] = e1; J[j++] = e2;
    J[j++] = -p[1] * t * e1;
    J[j++] = -p[2] * t * e2;
  }
}

static void f_helval(double *p, double *hx, int m, int n, void *a)
{
  double th; (void)m; (void)n; (void)a;
  if (p[0] < 0.0) th = atan(p[1] / p[0]) / (2.0 * KAT_PI) + 0.5;
  else if (0.0 < p[0]) th = atan(p[1] / p[0]) / (2.0 * KAT_PI);
  else th = (p[1] >= 0) ? 0.25 : -0.25;
  hx[0] = 10.0 * (p[2] - 10.0 * th);
  hx[1] = 10.0 * (sqrt(p[0] * p[0] + p[1] * p[1]) - 1.0);
  hx[2] = p[2];
}
static void j_helval(double *p, double *J, int m, int n, void *a)
{
  const double q = p[0] * p[0] + p[1] * p[1]; (void)m; (void)n; (void)a;
  J[0] = 50.0 * p[1] / (KAT_PI * q); J[1] = -50.0 * p[0] / (KAT_PI * q); J[2] = 10.0;
  J[3] = 10.0 * p[0] / sqrt(q);      J[4] = 10.0 * p[1] / sqrt(q);       J[5] = 0.0;
  J[6] = 0.0;                        J[7] = 0.0;                         J[8] = 1.0;
}

static const double meyer_x[16] = {34.780, 28.610, 23.650, 19.630, 16.370, 13.720, 11.540, 9.744,
                                   8.261,  7.030,  6.005,  5.147,  4.427,  3.820,  3.307,  2.872};
static const double osborne_x[33] = {
    8.44E-1, 9.08E-1, 9.32E-1, 9.36E-1, 9.25E-1, 9.08E-1, 8.81E-1, 8.5E-1,  8.18E-1, 7.84E-1, 7.51E-1,
    7.18E-1, 6.85E-1, 6.58E-1, 6.28E-1, 6.03E-1, 5.8E-1,  5.58E-1, 5.38E-1, 5.22E-1, 5.06E-1, 4.9E-1,
    4.78E-1, 4.67E-1, 4.57E-1, 4.48E-1, 4.38E-1, 4.31E-1, 4.24E-1, 4.2E-1,  4.14E-1, 4.11E-1, 4.06E-1};

/* Fills the starting point and data of problem `prob` (0..6).  Returns 0, or -1 if unknown.
 * *variant: 0 = analytic-Jacobian LM, 1 = secant LM (as the reference demo runs it). */
int orc_kat_problem(int prob, int *m, int *n, double *p0, double *x, int *variant, orc_func_t *func,
                    orc_func_t *jacf)
{
  int i;
  for (i = 0; i < 40; ++i) x[i] = 0.0;
  *variant = 0;
  switch (prob) {
    case 0: *m = 2; *n = 2; p0[0] = -1.2; p0[1] = 1.0; *func = f_ros; *jacf = j_ros; break;
    case 1: *m = 2; *n = 3; p0[0] = -1.2; p0[1] = 1.0; *func = f_modros; *jacf = j_modros; break;
    case 2: *m = 2; *n = 2; p0[0] = 3.0; p0[1] = 1.0; *func = f_powell; *jacf = j_powell; break;
    case 3: *m = 4; *n = 6; p0[0] = -3.0; p0[1] = -1.0; p0[2] = -3.0; p0[3] = -1.0;
            *func = f_wood; *jacf = 0; *variant = 1; break;
    case 4: *m = 3; *n = 16; p0[0] = 8.85; p0[1] = 4.0; p0[2] = 2.5;
            memcpy(x, meyer_x, sizeof meyer_x); *func = f_meyer; *jacf = 0; *variant = 1; break;
    case 5: *m = 5; *n = 33; p0[0] = 0.5; p0[1] = 1.5; p0[2] = -1.0; p0[3] = 1.0E-2; p0[4] = 2.0E-2;
            memcpy(x, osborne_x, sizeof osborne_x); *func = f_osborne; *jacf = j_osborne; break;
    case 6: *m = 3; *n = 3; p0[0] = -1.0; p0[1] = 0.0; p0[2] = 0.0; *func = f_helval; *jacf = j_helval; break;
    default: return -1;
  }
  return 0;
}
