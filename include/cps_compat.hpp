// cps_compat.hpp -- header-only C++ mirror of the reference's operator surface for the hot path, over the C ABI.
//
// The reference is C++ on the OpenCV 1.x C API; neither OpenCV nor GSL headers exist in this image, so the
// reference classes cannot be compiled here.  This header restates just enough of their interface -- the same
// class names, method names, argument order and public members -- that main.cpp / curveFittingClass.cpp style
// call sites compile against libcps.so with a minimal CvMat stand-in:
//   KalmanFilter           kalmanClass.h:40-163   (PredictKF, UpdateNCurvesAndPoints, UpdatePoints, getState,
//                                                   GetSplitMatrices, num_curves, num_points, curve_inds, point_inds)
//   CurveFittingClass      curveFittingClass.h:77-121 (fit_curve's LM call and post-normalisation, fitting_error)
//   struct mydata, modelfunc tokens                 curveFittingClass.h:22-31
// Everything numeric happens in the CUDA kernels behind cps_*; nothing here computes on the CPU except the
// reference's post-fit sign/angle normalisation of 19 numbers (curveFittingClass.cpp:684-711).
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
#include "cps_ekf.h"
#include "cps_lm.h"
}

namespace cps_compat {

// the fields of OpenCV 1.x's CvMat the hot path touches (rows, cols, data.db), CV_64FC1 only
struct CvMat {
  int rows = 0, cols = 0;
  union {
    double* db;
  } data;
  std::vector<double> storage;
  CvMat(int r, int c) : rows(r), cols(c), storage((size_t)r * c, 0.0) { data.db = storage.data(); }
};
inline double cvmGet(const CvMat* m, int r, int c) { return m->data.db[(size_t)r * m->cols + c]; }
inline void cvmSet(CvMat* m, int r, int c, double v) { m->data.db[(size_t)r * m->cols + c] = v; }

inline void check(int rc, const char* what) {
  if (rc != CPS_OK) throw std::runtime_error(std::string(what) + ": " + cps_last_error());
}

// struct mydata (curveFittingClass.h:22-31) and its restatement in cps_lm.h must agree field by field: the
// same-signature levmar entry points read the reference's own funcdata through it
struct reference_mydata_layout {  // the reference's declaration, with DisplayClass* as an opaque pointer
  double* orig_meas;
  double* t;
  int* curve_num;
  int num_left1;
  int num_right1;
  int num_left2;
  int num_right2;
  void* disp;
};
static_assert(sizeof(cps_mydata) == sizeof(reference_mydata_layout), "cps_mydata must mirror struct mydata");
static_assert(offsetof(cps_mydata, t) == offsetof(reference_mydata_layout, t) &&
                  offsetof(cps_mydata, num_left1) == offsetof(reference_mydata_layout, num_left1) &&
                  offsetof(cps_mydata, num_right1) == offsetof(reference_mydata_layout, num_right1) &&
                  offsetof(cps_mydata, num_left2) == offsetof(reference_mydata_layout, num_left2) &&
                  offsetof(cps_mydata, num_right2) == offsetof(reference_mydata_layout, num_right2) &&
                  offsetof(cps_mydata, disp) == offsetof(reference_mydata_layout, disp),
              "cps_mydata must mirror struct mydata (curveFittingClass.h:22-31)");

class KalmanFilter {
 public:
  // The reference's KalmanFilter() takes no argument and re-allocates P on every augmentation
  // (kalmanClass.cpp:362-389); here P lives on the device inside a fixed capacity: the default holds the pose, 64
  // curves and 500 points (2 024 states, 33 MB) -- pass a capacity for larger maps.
  explicit KalmanFilter(int capacity = 12 + 8 * 64 + 3 * 500, int device = 0) : x(nullptr), num_curves(0), num_points(0) {
    check(cps_ekf_create(&dev_, device, capacity), "cps_ekf_create");
  }
  ~KalmanFilter() {
    cps_ekf_destroy(dev_);
    delete x;
  }
  KalmanFilter(const KalmanFilter&) = delete;
  KalmanFilter& operator=(const KalmanFilter&) = delete;

  // load a state (the reference builds it with AddFirstStates/AddNewCurve/AddNewPoints, kalmanClass.cpp:275-666)
  void SetState(const double* xs, const double* P, int N, const std::vector<int>& curves, const std::vector<int>& points) {
    check(cps_ekf_set_state(dev_, xs, P, N, curves.data(), (int)curves.size(), points.data(), (int)points.size()),
          "cps_ekf_set_state");
    curve_inds = curves;
    point_inds = points;
    num_curves = (int)curves.size();
    num_points = (int)points.size();
    delete x;
    x = new CvMat(N, 1);
  }
  // state augmentation, kalmanClass.h:43-47, same argument lists (kalmanClass.cpp:275-311, 314-499, 502-666)
  void AddFirstStates(CvMat* measurement) {
    check(cps_ekf_add_first_states(dev_, measurement->data.db), "cps_ekf_add_first_states");
    refresh_tables();
  }
  void AddNewCurve(CvMat* measurement, CvMat* A) {
    check(cps_ekf_add_new_curve(dev_, measurement->data.db, A->data.db), "cps_ekf_add_new_curve");
    refresh_tables();
  }
  void AddNewPoints(double* measurements, int n_pts) {
    check(cps_ekf_add_new_points(dev_, measurements, n_pts), "cps_ekf_add_new_points");
    refresh_tables();
  }
  void PredictKF() { check(cps_ekf_predict(dev_), "cps_ekf_predict"); }  // kalmanClass.cpp:161-272
  // kalmanClass.cpp:669-945, same argument list
  void UpdateNCurvesAndPoints(CvMat* measurement, int n, std::vector<CvMat*>* A, std::vector<int>* curve_num,
                              double* point_meas, int* point_nums, int n_pts) {
    std::vector<double> Aflat((size_t)16 * (n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) std::memcpy(Aflat.data() + 16 * k, A->at(k)->data.db, 16 * sizeof(double));
    check(cps_ekf_update_curves_points(dev_, measurement->data.db, n, Aflat.data(), curve_num->data(), point_meas,
                                       point_nums, n_pts),
          "cps_ekf_update_curves_points");
  }
  void UpdatePoints(double* point_meas, int* point_nums, int n_pts) {  // kalmanClass.cpp:948-1126
    check(cps_ekf_update_points(dev_, point_meas, point_nums, n_pts), "cps_ekf_update_points");
  }
  CvMat* getState() {  // kalmanClass.cpp:1138-1141
    check(cps_ekf_get_state(dev_, x->data.db), "cps_ekf_get_state");
    return x;
  }
  void GetSplitMatrices(double t, CvMat* A1, CvMat* A2) { cps_ekf_split_matrices(t, A1->data.db, A2->data.db); }
  void GetCovBlock(int r0, int c0, int nr, int nc, double* out) {
    check(cps_ekf_get_cov_block(dev_, r0, c0, nr, nc, out), "cps_ekf_get_cov_block");
  }
  int GateMahalanobis(const double* z, int nz, double gate, int* idx, double* d2) {
    return cps_gate_mahalanobis(dev_, z, nz, gate, idx, d2);
  }

  // the reference's public member P (kalmanClass.h:153) read through a block view: the covariance stays on the
  // device; P_block(r0, c0, nr, nc) returns the nr x nc block as a CvMat (displayClass.cpp reads the pose block)
  CvMat P_block(int r0, int c0, int nr, int nc) {
    CvMat m(nr, nc);
    GetCovBlock(r0, c0, nr, nc, m.data.db);
    return m;
  }
  int dim() {
    int N = 0;
    check(cps_ekf_dims(dev_, &N, nullptr, nullptr), "cps_ekf_dims");
    return N;
  }

  CvMat* x;  // host view refreshed by getState(); the covariance stays on the device
  int num_curves, num_points;
  std::vector<int> curve_inds, point_inds;

 private:
  void refresh_tables() {
    int N = 0, nc = 0, np = 0;
    check(cps_ekf_dims(dev_, &N, &nc, &np), "cps_ekf_dims");
    curve_inds.assign((size_t)nc, 0);
    point_inds.assign((size_t)np, 0);
    std::vector<int> ci((size_t)nc + 1), pi((size_t)np + 1);
    check(cps_ekf_get_index_tables(dev_, ci.data(), pi.data()), "cps_ekf_get_index_tables");
    for (int k = 0; k < nc; ++k) curve_inds[(size_t)k] = ci[(size_t)k];
    for (int k = 0; k < np; ++k) point_inds[(size_t)k] = pi[(size_t)k];
    num_curves = nc;
    num_points = np;
    delete x;
    x = new CvMat(N, 1);
  }
  cps_ekf* dev_ = nullptr;
};

class CurveFittingClass {
 public:
  explicit CurveFittingClass(int device = 0) : fitting_error(0.0) { check(cps_lm_create(&ctx_, device), "cps_lm_create"); }
  ~CurveFittingClass() { cps_lm_destroy(ctx_); }
  CurveFittingClass(const CurveFittingClass&) = delete;
  CurveFittingClass& operator=(const CurveFittingClass&) = delete;

  // fit_curve (curveFittingClass.cpp:408-724): like the reference it IGNORES params on entry -- the initial guess is
  // the stereo triangulation + plane fit of :516-651 (cps_fit_curve_init_batched) -- and returns the fitted,
  // normalised parameters; L/R are the 2k x 1 interleaved (x,y) contour matrices of the left / right image.  The
  // fourth argument of the reference (DisplayClass*, only stored) is accepted and ignored.
  void fit_curve(double* params, CvMat** featuresLeftImage, CvMat** featuresRightImage, void* /*display*/ = nullptr) {
    CvMat* seg[4] = {featuresLeftImage[0], featuresLeftImage[1], featuresRightImage[0], featuresRightImage[1]};
    std::vector<double> meas;
    int counts[4];
    for (int s = 0; s < 4; ++s) {  // measurement order [L c1 | L c2 | R c1 | R c2], :435-489
      counts[s] = seg[s]->rows / 2;
      meas.insert(meas.end(), seg[s]->data.db, seg[s]->data.db + 2 * counts[s]);
    }
    long long off[2] = {0, (long long)meas.size()};
    check(cps_fit_curve_init_batched(ctx_, 1, meas.data(), off, counts, params), "cps_fit_curve_init_batched");
    double opts[5] = {1E-10, 1E-10, 1E-10, 1E-20, 1E-06 * 1E1};  // :661-662
    double info[CPS_LM_INFO_SZ];
    int ret = 0;
    check(cps_dlevmar_dif_batched(ctx_, CPS_MODEL_STEREO19, 1, params, meas.data(), nullptr, off, counts, 1000, opts,
                                  info, &ret, nullptr),
          "cps_dlevmar_dif_batched");  // the live call, :681
    fitting_error = info[1];            // :683
    normalise(params);                  // :684-711
  }
  static void normalise(double* p) {
    const double PI = 3.1415926535;  // common.h:114
    if (p[18] > 0.0) {
      for (int i = 0; i < 8; i++) p[2 * i + 1] *= -1.0;
      p[16] *= -1.0;
      p[17] += PI;
      p[18] *= -1.0;
    }
    if (p[16] < -PI / 2.0) {
      for (int i = 0; i < 8; i++) p[2 * i + 1] *= -1.0;
      p[17] = (-PI - p[17]);
      p[18] += PI;
    }
    while (p[16] > PI) p[16] -= 2 * PI;
    while (p[16] < -PI) p[16] += 2 * PI;
    while (p[17] > PI) p[17] -= 2 * PI;
    while (p[17] < -PI) p[17] += 2 * PI;
  }
  double fitting_error;

 private:
  cps_lm_ctx* ctx_ = nullptr;
};

}  // namespace cps_compat
