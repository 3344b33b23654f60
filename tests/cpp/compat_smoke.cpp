// C++ caller compiled against include/cps_compat.hpp + libcps.so: exercises the reference-style call sites
// (KalmanFilter::PredictKF / UpdateNCurvesAndPoints / getState, CurveFittingClass::fit_curve) the way main.cpp
// would (main.cpp:388,445,965).  Reads a small binary case written by tests/test_cpp_compat.py and writes the
// results back for comparison with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cps_compat.hpp"

using namespace cps_compat;

static std::vector<double> read_doubles(FILE* f, size_t n) {
  std::vector<double> v(n);
  if (fread(v.data(), sizeof(double), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4];  // N, n_curves_state, n_meas_curves, K (points per contour segment)
  if (fread(hdr, sizeof(int), 4, f) != 4) return 2;
  const int N = hdr[0], nc = hdr[1], n = hdr[2], K = hdr[3];
  std::vector<double> x = read_doubles(f, N), P = read_doubles(f, (size_t)N * N), z = read_doubles(f, 8 * n + 3);
  std::vector<double> Aall = read_doubles(f, 16 * (size_t)n);
  std::vector<double> p0 = read_doubles(f, 19), meas = read_doubles(f, 8 * (size_t)K);
  fclose(f);

  // EKF: predict + update with the reference's argument list
  KalmanFilter kf(N);
  std::vector<int> curves(nc), points;
  for (int k = 0; k < nc; ++k) curves[k] = 12 + 8 * k;
  kf.SetState(x.data(), P.data(), N, curves, points);
  kf.PredictKF();
  CvMat zmat(8 * n + 3, 1);
  for (int i = 0; i < 8 * n + 3; ++i) zmat.data.db[i] = z[i];
  std::vector<CvMat*> A;
  std::vector<int> curve_num;
  for (int k = 0; k < n; ++k) {
    CvMat* a = new CvMat(4, 4);
    for (int i = 0; i < 16; ++i) a->data.db[i] = Aall[16 * k + i];
    A.push_back(a);
    curve_num.push_back(k);
  }
  kf.UpdateNCurvesAndPoints(&zmat, n, &A, &curve_num, nullptr, nullptr, 0);
  CvMat* xs = kf.getState();
  std::vector<double> Pout((size_t)N * N);
  kf.GetCovBlock(0, 0, N, N, Pout.data());

  // curve fit through the fit_curve mirror
  CurveFittingClass fitter;
  CvMat L0(2 * K, 1), L1(2 * K, 1), R0(2 * K, 1), R1(2 * K, 1);
  CvMat* Ls[2] = {&L0, &L1};
  CvMat* Rs[2] = {&R0, &R1};
  CvMat* segs[4] = {&L0, &L1, &R0, &R1};
  for (int s = 0; s < 4; ++s)
    for (int i = 0; i < 2 * K; ++i) segs[s]->data.db[i] = meas[(size_t)s * 2 * K + i];
  std::vector<double> params = p0;
  fitter.fit_curve(params.data(), Ls, Rs);

  FILE* o = fopen(argv[2], "wb");
  fwrite(xs->data.db, sizeof(double), N, o);
  fwrite(Pout.data(), sizeof(double), Pout.size(), o);
  fwrite(params.data(), sizeof(double), 19, o);
  fwrite(&fitter.fitting_error, sizeof(double), 1, o);
  fclose(o);
  for (CvMat* a : A) delete a;
  printf("compat_smoke ok\n");
  return 0;
}
