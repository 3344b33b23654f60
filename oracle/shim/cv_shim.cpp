/*
 * oracle/shim/cv_shim.cpp -- implementations behind oracle/shim/cxcore.h and gsl/gsl_poly.h.
 * TEST INFRASTRUCTURE ONLY (see cxcore.h): lets `make -C oracle ref` compile the reference's
 * kalmanClass.cpp and curveFittingClass.cpp unchanged.  Plain row-major double matrices, textbook loops:
 * products sum over ascending k, the SVD is the one-sided (Hestenes) Jacobi iteration, the CV_SVD inverse is the
 * pseudo-inverse V diag(1/w) U^T with singular values below 2 eps sum(w) dropped.  Built with
 * -ffp-contract=off like the rest of the checker.
 */
#include "cxcore.h"
#include "gsl/gsl_poly.h"

#include <stdexcept>
#include <vector>

extern "C" {

CvMat *cvCreateMat(int rows, int cols, int type)
{
  if (type != CV_64FC1) {
    fprintf(stderr, "cv_shim: only CV_64FC1 matrices exist here\n");
    abort();
  }
  if (rows <= 0 || cols <= 0) throw std::runtime_error("cv_shim: cvCreateMat with a non-positive size (OpenCV raises an error here)");
  CvMat *m = (CvMat *)calloc(1, sizeof(CvMat));
  m->type = type;
  m->rows = rows;
  m->cols = cols;
  m->step = cols * (int)sizeof(double);
  /* cvCreateMat leaves the payload uninitialised; a fixed pattern keeps runs reproducible */
  size_t cnt = (size_t)rows * cols;
  m->data.db = (double *)malloc((cnt ? cnt : 1) * sizeof(double));
  for (size_t i = 0; i < cnt; ++i) m->data.db[i] = 0.0;
  return m;
}

void cvReleaseMat(CvMat **m)
{
  if (m && *m) {
    free((*m)->data.db);
    free(*m);
    *m = NULL;
  }
}

CvMat *cvCloneMat(const CvMat *m)
{
  CvMat *c = cvCreateMat(m->rows, m->cols, m->type);
  memcpy(c->data.db, m->data.db, (size_t)m->rows * m->cols * sizeof(double));
  return c;
}

void cvSetZero(CvMat *m) { memset(m->data.db, 0, (size_t)m->rows * m->cols * sizeof(double)); }

static void need(bool ok, const char *what)
{
  if (!ok) {
    fprintf(stderr, "cv_shim: size mismatch in %s\n", what);
    abort();
  }
}

void cvCopy(const CvMat *src, CvMat *dst)
{
  need(src->rows == dst->rows && src->cols == dst->cols, "cvCopy");
  memmove(dst->data.db, src->data.db, (size_t)src->rows * src->cols * sizeof(double));
}

void cvGEMM(const CvMat *a, const CvMat *b, double alpha, const CvMat *c, double beta, CvMat *dst, int flags)
{
  need(flags == 0 && a->cols == b->rows && dst->rows == a->rows && dst->cols == b->cols, "cvGEMM");
  const int R = a->rows, K = a->cols, C = b->cols;
  std::vector<double> acc((size_t)R * C, 0.0);
  for (int i = 0; i < R; ++i) {
    double *o = &acc[(size_t)i * C];
    for (int k = 0; k < K; ++k) {
      const double aik = a->data.db[(size_t)i * K + k];
      const double *br = &b->data.db[(size_t)k * C];
      for (int j = 0; j < C; ++j) o[j] += aik * br[j];
    }
  }
  for (size_t i = 0; i < acc.size(); ++i) {
    double v = alpha * acc[i];
    if (c) v += beta * c->data.db[i];
    dst->data.db[i] = v;
  }
}

void cvTranspose(const CvMat *src, CvMat *dst)
{
  need(src->rows == dst->cols && src->cols == dst->rows, "cvTranspose");
  const int R = src->rows, C = src->cols;
  if (src->data.db == dst->data.db) {
    for (int i = 0; i < R; ++i)
      for (int j = i + 1; j < C; ++j) std::swap(dst->data.db[(size_t)i * C + j], dst->data.db[(size_t)j * C + i]);
    return;
  }
  const int T = 32;
  for (int i0 = 0; i0 < R; i0 += T)
    for (int j0 = 0; j0 < C; j0 += T)
      for (int i = i0; i < std::min(R, i0 + T); ++i)
        for (int j = j0; j < std::min(C, j0 + T); ++j) dst->data.db[(size_t)j * R + i] = src->data.db[(size_t)i * C + j];
}

void cps_shim_add(const CvMat *a, const CvMat *b, CvMat *dst, double sb)
{
  need(a->rows == b->rows && a->cols == b->cols && a->rows == dst->rows && a->cols == dst->cols, "cvAdd/cvSub");
  const size_t n = (size_t)a->rows * a->cols;
  if (sb < 0)
    for (size_t i = 0; i < n; ++i) dst->data.db[i] = a->data.db[i] - b->data.db[i];
  else
    for (size_t i = 0; i < n; ++i) dst->data.db[i] = a->data.db[i] + b->data.db[i];
}

void cvAddWeighted(const CvMat *a, double alpha, const CvMat *b, double beta, double gamma, CvMat *dst)
{
  need(a->rows == b->rows && a->cols == b->cols && a->rows == dst->rows && a->cols == dst->cols, "cvAddWeighted");
  const size_t n = (size_t)a->rows * a->cols;
  for (size_t i = 0; i < n; ++i) dst->data.db[i] = a->data.db[i] * alpha + b->data.db[i] * beta + gamma;
}

double cvNorm(const CvMat *a, const CvMat *a2, int norm_type, const CvMat *mask)
{
  need(a2 == NULL && mask == NULL && norm_type == CV_L2, "cvNorm");
  double s = 0.0;
  const size_t n = (size_t)a->rows * a->cols;
  for (size_t i = 0; i < n; ++i) s += a->data.db[i] * a->data.db[i];
  return sqrt(s);
}

double cvDotProduct(const CvMat *a, const CvMat *b)
{
  need(a->rows * a->cols == b->rows * b->cols, "cvDotProduct");
  double s = 0.0;
  const size_t n = (size_t)a->rows * a->cols;
  for (size_t i = 0; i < n; ++i) s += a->data.db[i] * b->data.db[i];
  return s;
}

/* One-sided Jacobi SVD of a (m x n, m >= n).  On return the n rows of `at` (n x m, = a^T on entry) hold
 * u_i * w_i, `vt` (n x n) holds the right singular vectors as rows, w descending. */
static void jacobi_svd(std::vector<double> &at, std::vector<double> &vt, std::vector<double> &w, int m, int n)
{
  const double eps = DBL_EPSILON * 10;
  vt.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) vt[(size_t)i * n + i] = 1.0;
  w.assign(n, 0.0);
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += at[(size_t)i * m + k] * at[(size_t)i * m + k];
    w[i] = s;
  }
  const int max_sweeps = std::max(m, 30);
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double *ai = &at[(size_t)i * m], *aj = &at[(size_t)j * m];
        double a = w[i], b = w[j], p = 0;
        for (int k = 0; k < m; ++k) p += ai[k] * aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = c * ai[k] + s * aj[k], t1 = -s * ai[k] + c * aj[k];
          ai[k] = t0;
          aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        w[i] = a;
        w[j] = b;
        changed = true;
        double *vi = &vt[(size_t)i * n], *vj = &vt[(size_t)j * n];
        for (int k = 0; k < n; ++k) {
          const double t0 = c * vi[k] + s * vj[k], t1 = -s * vi[k] + c * vj[k];
          vi[k] = t0;
          vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += at[(size_t)i * m + k] * at[(size_t)i * m + k];
    w[i] = sqrt(s);
  }
  /* selection sort, descending; rows of at and vt follow */
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (w[j] < w[k]) j = k;
    if (i != j) {
      std::swap(w[i], w[j]);
      for (int k = 0; k < m; ++k) std::swap(at[(size_t)i * m + k], at[(size_t)j * m + k]);
      for (int k = 0; k < n; ++k) std::swap(vt[(size_t)i * n + k], vt[(size_t)j * n + k]);
    }
  }
}

void cvSVD(CvMat *a, CvMat *w, CvMat *u, CvMat *v, int flags)
{
  const int m = a->rows, n = a->cols;
  need(flags == 0 && m >= n && w->rows * w->cols == n && u->rows == m && u->cols == n && v->rows == n && v->cols == n,
       "cvSVD");
  std::vector<double> at((size_t)n * m), vt, sv;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) at[(size_t)j * m + i] = a->data.db[(size_t)i * n + j];
  jacobi_svd(at, vt, sv, m, n);
  for (int j = 0; j < n; ++j) {
    w->data.db[j] = sv[j];
    const double inv = sv[j] > 0 ? 1.0 / sv[j] : 0.0;
    for (int i = 0; i < m; ++i) u->data.db[(size_t)i * n + j] = at[(size_t)j * m + i] * inv;
    for (int i = 0; i < n; ++i) v->data.db[(size_t)i * n + j] = vt[(size_t)j * n + i];
  }
}

double cvInvert(const CvMat *src, CvMat *dst, int method)
{
  const int n = src->rows;
  need(method == CV_SVD && src->cols == n && dst->rows == n && dst->cols == n, "cvInvert");
  std::vector<double> at((size_t)n * n), vt, sv;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) at[(size_t)j * n + i] = src->data.db[(size_t)i * n + j];
  jacobi_svd(at, vt, sv, n, n);
  double thresh = 0;
  for (int i = 0; i < n; ++i) thresh += sv[i];
  thresh *= 2 * DBL_EPSILON;
  /* dst = sum_k v_k (1/w_k) u_k^T ;  at row k = u_k w_k  =>  u_k / w_k = at_k / w_k^2 */
  std::vector<double> out((size_t)n * n, 0.0);
  for (int k = 0; k < n; ++k) {
    if (sv[k] <= thresh) continue;
    const double inv2 = 1.0 / (sv[k] * sv[k]);
    for (int i = 0; i < n; ++i) {
      const double vik = vt[(size_t)k * n + i] * inv2;
      for (int j = 0; j < n; ++j) out[(size_t)i * n + j] += vik * at[(size_t)k * n + j];
    }
  }
  memcpy(dst->data.db, out.data(), out.size() * sizeof(double));
  return (n > 0 && sv[0] > 0) ? sv[n - 1] / sv[0] : 0.0;
}

IplImage *cvCreateImage(CvSize, int, int)
{
  fprintf(stderr, "cv_shim: cvCreateImage is not part of the tested path\n");
  abort();
}

/* ---- gsl_poly_complex_solve: the root finder itself is the oracle's restatement of GSL's documented method
 * (oracle/orc_closest.c: balanced companion matrix, Francis double-shift QR); what compiling the reference
 * against it pins is the reference's code AROUND the solver (polynomial construction, root selection). ------- */
int orc_poly_roots(const double *a, int deg, double *zr, double *zi);

gsl_poly_complex_workspace *gsl_poly_complex_workspace_alloc(size_t n)
{
  gsl_poly_complex_workspace *w = (gsl_poly_complex_workspace *)calloc(1, sizeof(*w));
  w->nc = n ? n - 1 : 0;
  w->matrix = NULL;
  return w;
}

void gsl_poly_complex_workspace_free(gsl_poly_complex_workspace *w) { free(w); }

int gsl_poly_complex_solve(const double *a, size_t n, gsl_poly_complex_workspace *, gsl_complex_packed_ptr z)
{
  if (n < 2 || a[n - 1] == 0.0) return 1;
  const int deg = (int)n - 1;
  double zr[16], zi[16];
  const int rc = orc_poly_roots(a, deg, zr, zi);
  for (int i = 0; i < deg; ++i) {
    z[2 * i] = zr[i];
    z[2 * i + 1] = zi[i];
  }
  return rc;
}

} /* extern "C" */
