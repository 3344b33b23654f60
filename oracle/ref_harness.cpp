/*
 * oracle/ref_harness.cpp -- a C-callable harness around the REFERENCE's own classes and free functions,
 * compiled UNCHANGED from /root/reference (kalmanClass.cpp, curveFittingClass.cpp, levmar-2.5) against the
 * minimal OpenCV-1.x / GSL stand-ins of oracle/shim/.  TEST INFRASTRUCTURE ONLY.  Output: oracle/_ref/libcps_ref.so
 * (git-ignored).  Used by tests/ to pin the oracle restatements (orc_ekf.c, orc_model.c, orc_closest.c ...) to the
 * reference's code and by tests/golden/gen_ref_golden.py to write the committed fixtures.
 *
 * Globals the reference expects from main.cpp: `binom` (main.cpp:56, filled at :166-178).
 */
#include "kalmanClass.h"
#include "curveFittingClass.h"

unsigned int binom[50][50];

/* DisplayClass is only ever passed around as a pointer on this path (curveFittingClass.cpp:414); its methods and
 * the feature detector are never linked. */

namespace {

struct BinomInit {
  BinomInit()
  {
    /* Pascal table exactly as main.cpp:166-178 builds it */
    for (int n = 0; n < 50; n++)
      for (int k = 0; k < 50; k++) {
        if (k == 0)
          binom[n][k] = 1;
        else if (n == 0)
          binom[n][k] = 0;
        else
          binom[n][k] = binom[n - 1][k] + binom[n - 1][k - 1];
      }
  }
} binom_init;

CvMat *mat_from(const double *src, int rows, int cols)
{
  CvMat *m = cvCreateMat(rows, cols, CV_64FC1);
  if (src) memcpy(m->data.db, src, (size_t)rows * cols * sizeof(double));
  return m;
}

CurveFittingClass *fitter()
{
  static CurveFittingClass *f = new CurveFittingClass();
  return f;
}

} // namespace

extern "C" {

/* ---- KalmanFilter (kalmanClass.h:40-163) ------------------------------------------------------------------ */
void *refh_kf_create(void) { return new KalmanFilter(); }
void refh_kf_destroy(void *h) { delete (KalmanFilter *)h; }

/* Replaces x, P and the landmark tables (the members are public: kalmanClass.h:116-118,152-162). */
void refh_kf_set(void *h, const double *x, const double *P, int N, const int *curve_inds, int nc, const int *point_inds,
                 int np)
{
  KalmanFilter *kf = (KalmanFilter *)h;
  cvReleaseMat(&kf->x);
  cvReleaseMat(&kf->P);
  kf->x = mat_from(x, N, 1);
  kf->P = mat_from(P, N, N);
  kf->num_curves = nc;
  kf->num_points = np;
  kf->curve_inds.assign(curve_inds, curve_inds + nc);
  kf->point_inds.assign(point_inds, point_inds + np);
}

int refh_kf_dim(void *h) { return ((KalmanFilter *)h)->x->rows; }
int refh_kf_num_curves(void *h) { return ((KalmanFilter *)h)->num_curves; }
int refh_kf_num_points(void *h) { return ((KalmanFilter *)h)->num_points; }

void refh_kf_get(void *h, double *x, double *P, int *curve_inds, int *point_inds)
{
  KalmanFilter *kf = (KalmanFilter *)h;
  const int N = kf->x->rows;
  if (x) memcpy(x, kf->x->data.db, (size_t)N * sizeof(double));
  if (P) memcpy(P, kf->P->data.db, (size_t)N * N * sizeof(double));
  if (curve_inds)
    for (size_t i = 0; i < kf->curve_inds.size(); ++i) curve_inds[i] = kf->curve_inds[i];
  if (point_inds)
    for (size_t i = 0; i < kf->point_inds.size(); ++i) point_inds[i] = kf->point_inds[i];
}

void refh_kf_predict(void *h) { ((KalmanFilter *)h)->PredictKF(); }

void refh_kf_add_first_states(void *h, const double *z19)
{
  CvMat *z = mat_from(z19, 19, 1);
  ((KalmanFilter *)h)->AddFirstStates(z);
  cvReleaseMat(&z);
}

void refh_kf_add_new_curve(void *h, const double *z8, const double *A16)
{
  CvMat *z = mat_from(z8, 8, 1), *A = mat_from(A16, 4, 4);
  ((KalmanFilter *)h)->AddNewCurve(z, A);
  cvReleaseMat(&z);
  cvReleaseMat(&A);
}

void refh_kf_add_new_points(void *h, const double *xyz, int n_pts)
{
  std::vector<double> m(xyz, xyz + 3 * (size_t)n_pts);
  ((KalmanFilter *)h)->AddNewPoints(m.data(), n_pts);
}

/* UpdateNCurvesAndPoints(z, n, &A, &curve_num, point_meas, point_nums, n_pts), kalmanClass.cpp:669 */
void refh_kf_update_curves_points(void *h, const double *z, int n, const double *A /* n x 16 */, const int *curve_num,
                                  const double *point_meas, const int *point_nums, int n_pts)
{
  const int M = 8 * n + 3 + 3 * n_pts;
  (void)M;
  CvMat *zm = mat_from(z, 8 * n + 3, 1);
  std::vector<CvMat *> As;
  std::vector<int> nums(curve_num, curve_num + n);
  for (int k = 0; k < n; ++k) As.push_back(mat_from(A + 16 * k, 4, 4));
  std::vector<double> pm(point_meas, point_meas + 3 * (size_t)n_pts);
  std::vector<int> pn(point_nums, point_nums + n_pts);
  pm.push_back(0.0);
  pn.push_back(0);
  ((KalmanFilter *)h)->UpdateNCurvesAndPoints(zm, n, &As, &nums, pm.data(), pn.data(), n_pts);
  for (int k = 0; k < n; ++k) cvReleaseMat(&As[k]);
  cvReleaseMat(&zm);
}

void refh_kf_update_points(void *h, const double *point_meas, const int *point_nums, int n_pts)
{
  std::vector<double> pm(point_meas, point_meas + 3 * (size_t)n_pts);
  std::vector<int> pn(point_nums, point_nums + n_pts);
  ((KalmanFilter *)h)->UpdatePoints(pm.data(), pn.data(), n_pts);
}

void refh_kf_split_matrices(void *h, double t, double *A1, double *A2)
{
  CvMat *a1 = mat_from(NULL, 4, 4), *a2 = mat_from(NULL, 4, 4);
  ((KalmanFilter *)h)->GetSplitMatrices(t, a1, a2);
  memcpy(A1, a1->data.db, 16 * sizeof(double));
  memcpy(A2, a2->data.db, 16 * sizeof(double));
  cvReleaseMat(&a1);
  cvReleaseMat(&a2);
}

void refh_kf_predicted_measurement(void *h, const double *A16, int num_curve, double *zhat8)
{
  KalmanFilter *kf = (KalmanFilter *)h;
  CvMat *zh = mat_from(NULL, 8, 1), *A = mat_from(A16, 4, 4);
  kf->GetPredictedMeasurement(zh, kf->x, A, num_curve);
  memcpy(zhat8, zh->data.db, 8 * sizeof(double));
  cvReleaseMat(&zh);
  cvReleaseMat(&A);
}

void refh_kf_rbe_derivs(void *h, double phi, double theta, double psi, double *dphi, double *dtheta, double *dpsi)
{
  CvMat *a = mat_from(NULL, 3, 3), *b = mat_from(NULL, 3, 3), *c = mat_from(NULL, 3, 3);
  ((KalmanFilter *)h)->get_Rbe_derivs(phi, theta, psi, a, b, c);
  memcpy(dphi, a->data.db, 72);
  memcpy(dtheta, b->data.db, 72);
  memcpy(dpsi, c->data.db, 72);
  cvReleaseMat(&a);
  cvReleaseMat(&b);
  cvReleaseMat(&c);
}

void refh_generate_rbe(double phi, double theta, double psi, double *R9)
{
  CvMat *r = mat_from(NULL, 3, 3);
  generate_Rbe(phi, theta, psi, r);
  memcpy(R9, r->data.db, 72);
  cvReleaseMat(&r);
}

/* ---- curve model and fit driver (curveFittingClass.h:33-75, 77-120) ------------------------------------------- */

/* modelfunc(cp, x, m, n, adata) with adata = struct mydata (curveFittingClass.h:22-31) */
void refh_modelfunc(const double *p19, double *hx, int n, const double *t, int n_l1, int n_r1, int n_l2, int n_r2)
{
  fitter();
  std::vector<double> p(p19, p19 + 19), tt(t, t + n / 2), orig(n, 0.0);
  std::vector<int> cn(n / 2, 0);
  struct mydata d;
  d.orig_meas = orig.data();
  d.t = tt.data();
  d.curve_num = cn.data();
  d.num_left1 = n_l1;
  d.num_right1 = n_r1;
  d.num_left2 = n_l2;
  d.num_right2 = n_r2;
  d.disp = NULL;
  modelfunc(p.data(), hx, 19, n, &d);
}

void refh_cp_earth2image(const double *c_earth8, double *c_left8, double *c_right8, const double *euler3,
                         const double *translation3)
{
  fitter();
  double ce[8], eu[3], tr[3];
  memcpy(ce, c_earth8, 64);
  memcpy(eu, euler3, 24);
  memcpy(tr, translation3, 24);
  cp_earth2image(ce, c_left8, c_right8, eu, tr);
}

void refh_earth2body(const double *xe3, double *xb3, const double *euler3, const double *translation3)
{
  fitter();
  double a[3], eu[3], tr[3];
  memcpy(a, xe3, 24);
  memcpy(eu, euler3, 24);
  memcpy(tr, translation3, 24);
  earth2body(a, xb3, eu, tr);
}

double refh_bezier_eval(const double *cp8, double t, const double *pt2, double *nearest2)
{
  double cp[8], pt[2];
  memcpy(cp, cp8, 64);
  memcpy(pt, pt2, 16);
  return closest_bezier_pt(cp, t, pt, nearest2);
}

void refh_closest_bezier_pt(const double *p8, const double *pt2, double *nearest2, double *out2)
{
  fitter();
  double p[8], pt[2];
  memcpy(p, p8, 64);
  memcpy(pt, pt2, 16);
  closest_bezier_pt(p, pt, nearest2, out2);
}

void refh_control2coeffs(const double *cp8, double *p8)
{
  double cp[8];
  memcpy(cp, cp8, 64);
  control2coeffs(cp, p8);
}

/* ---- the hook between fit_curve and levmar --------------------------------------------------------------------
 * oracle/Makefile compiles curveFittingClass.cpp with -Ddlevmar_dif=refh_dlevmar_dif_hook (a command-line macro: the
 * source stays unchanged), so the reference's one live call (curveFittingClass.cpp:681) lands here.  The hook
 * records what fit_curve hands to levmar -- the initial guess, the packed measurements, t, the segment sizes, opts,
 * itmax -- forwards to the reference's real dlevmar_dif, and records what came back before fit_curve's
 * post-normalisation touches it. */
#ifndef REFH_GPU
static struct {
  std::vector<double> p0, x, t, p_lm;
  int m, n, itmax, counts[4], ret, calls;
  double opts[5], info[10];
} g_hook;

int refh_dlevmar_dif_hook(void (*func)(double *p, double *hx, int m, int n, void *adata), double *p, double *x, int m,
                          int n, int itmax, double *opts, double *info, double *work, double *covar, void *adata)
{
  struct mydata *d = (struct mydata *)adata;
  g_hook.calls++;
  g_hook.m = m;
  g_hook.n = n;
  g_hook.itmax = itmax;
  g_hook.p0.assign(p, p + m);
  g_hook.x.assign(x, x + n);
  g_hook.t.assign(d->t, d->t + n / 2);
  g_hook.counts[0] = d->num_left1;
  g_hook.counts[1] = d->num_left2;
  g_hook.counts[2] = d->num_right1;
  g_hook.counts[3] = d->num_right2;
  for (int i = 0; i < 5; ++i) g_hook.opts[i] = opts[i];
  const int ret = dlevmar_dif(func, p, x, m, n, itmax, opts, info, work, covar, adata);
  g_hook.ret = ret;
  g_hook.p_lm.assign(p, p + m);
  for (int i = 0; i < 10; ++i) g_hook.info[i] = info[i];
  return ret;
}

/* what the last fit_curve handed to / got back from levmar; any pointer may be NULL.  Returns n. */
int refh_last_fit(double *p0, double *x, double *t, int *counts4, double *opts5, int *itmax, double *p_lm,
                  double *info10, int *ret)
{
  const int m = g_hook.m, n = g_hook.n;
  if (p0) memcpy(p0, g_hook.p0.data(), m * sizeof(double));
  if (x) memcpy(x, g_hook.x.data(), n * sizeof(double));
  if (t) memcpy(t, g_hook.t.data(), (n / 2) * sizeof(double));
  if (counts4) memcpy(counts4, g_hook.counts, 4 * sizeof(int));
  if (opts5) memcpy(opts5, g_hook.opts, 5 * sizeof(double));
  if (itmax) *itmax = g_hook.itmax;
  if (p_lm) memcpy(p_lm, g_hook.p_lm.data(), m * sizeof(double));
  if (info10) memcpy(info10, g_hook.info, 10 * sizeof(double));
  if (ret) *ret = g_hook.ret;
  return n;
}

#else  /* REFH_GPU */
/* The relinked build (oracle/Makefile, libcps_ref_gpu.so): curveFittingClass.cpp is compiled without any macro, its
 * dlevmar_dif call (:681) binds to libcps_levmar.so, i.e. to the CUDA path.  The one line a maintainer adds to main():
 * tell the library that the pointer `modelfunc` names the Stereo19 device model. */
extern "C" int cps_register_model(void (*func)(double *, double *, int, int, void *), int model_id);
int refh_register_gpu_model(void) { return cps_register_model(modelfunc, 19); }
#endif

/* fit_curve(params, L, R, display): the four contour segments as 2k x 1 CvMat of interleaved (x, y), exactly what
 * cleanup_and_group_edges hands over (curveFittingClass.cpp:1397-1414).  Returns fitting_error (= info[1]). */
double refh_fit_curve(double *params19, const double *l1, int n_l1, const double *l2, int n_l2, const double *r1,
                      int n_r1, const double *r2, int n_r2)
{
  CvMat *L[2] = {mat_from(l1, 2 * n_l1, 1), mat_from(l2, 2 * n_l2, 1)};
  CvMat *R[2] = {mat_from(r1, 2 * n_r1, 1), mat_from(r2, 2 * n_r2, 1)};
  CurveFittingClass *f = fitter();
  f->fit_curve(params19, L, R, NULL);
  for (int i = 0; i < 2; ++i) {
    cvReleaseMat(&L[i]);
    cvReleaseMat(&R[i]);
  }
  return f->fitting_error;
}

/* cleanup_and_group_edges(featuresL, featuresR, Lnew, Rnew, reset_data_assoc, left_y_cutoff, tracked_endpt_y),
 * curveFittingClass.cpp:1207.  pts: int (x, y) pairs of the four contours [L0 | L1 | R0 | R1], seg_off[5] in points.
 * out_counts[4] and out_meas (capacity cap doubles, packed [L0|L1|R0|R1]) receive the kept samples.  Returns the
 * reference's bool, or -1 when the call raised (cvCreateMat on an empty range). */
int refh_cleanup_and_group_edges(const int *pts, const long long *seg_off, int reset_data_assoc, int *left_y_cutoff2,
                                 double *tracked_endpt_y2, int *out_counts, double *out_meas, long long cap)
{
  std::vector<CvPoint> seg[4];
  for (int s = 0; s < 4; ++s)
    for (long long i = seg_off[s]; i < seg_off[s + 1]; ++i) {
      CvPoint q;
      q.x = pts[2 * i];
      q.y = pts[2 * i + 1];
      seg[s].push_back(q);
    }
  std::vector<CvPoint> *Lp[2] = {&seg[0], &seg[1]}, *Rp[2] = {&seg[2], &seg[3]};
  CvMat *Ln[2] = {NULL, NULL}, *Rn[2] = {NULL, NULL};
  bool ok;
  try {
    ok = fitter()->cleanup_and_group_edges(Lp, Rp, Ln, Rn, reset_data_assoc != 0, left_y_cutoff2, tracked_endpt_y2);
  } catch (...) { /* an empty kept range: OpenCV's cvCreateMat raises an error the reference does not handle */
    for (int s = 0; s < 4; ++s) out_counts[s] = 0;
    return -1;
  }
  long long w = 0;
  CvMat *order[4] = {Ln[0], Ln[1], Rn[0], Rn[1]};
  for (int s = 0; s < 4; ++s) {
    out_counts[s] = 0;
    if (ok && order[s]) {
      out_counts[s] = order[s]->rows / 2;
      for (int i = 0; i < order[s]->rows && w < cap; ++i) out_meas[w++] = order[s]->data.db[i];
    }
  }
  return ok ? 1 : 0;
}

} /* extern "C" */
