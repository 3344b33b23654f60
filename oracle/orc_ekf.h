/*
 * orc_ekf.h -- CPU oracle for the EKF hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Line-by-line restatement (not a copy) of KalmanFilter::PredictKF (kalmanClass.cpp:161-272),
 * UpdateNCurvesAndPoints (:669-945), UpdatePoints (:948-1126), GetPredictedMeasurement (:1162-1199),
 * GetSplitMatrices (:1144-1159), get_Rbe_derivs (:1275-1298), generate_Rbe (common.h:174-186) on plain
 * row-major dense matrices.  The OpenCV routines the reference calls (cvGEMM, cvInvert(CV_SVD), cvTranspose,
 * cvAddWeighted) are not vendored under /root/reference (version unpinned); their arithmetic is restated as: dot
 * products accumulated in ascending k, a one-sided Jacobi SVD pseudo-inverse for cvInvert(CV_SVD).
 * PARITY PINNING: the reference's kalmanClass.cpp is compiled UNCHANGED by `make -C oracle ref` (against the
 * OpenCV-1.x stand-ins of oracle/shim/) into oracle/_ref/libcps_ref.so; this restatement is checked against that
 * build live and against the fixtures it wrote (tests/golden/ref_cps.npz, tests/test_ref_pinning.py): predict, H
 * ingredients, predicted measurement, split matrices, rotation tables bit for bit; updates and augmentation within
 * 1e-12 (two different SVD routines).  Also kept: the cv2 4.13 transcription of the same lines
 * (tests/golden/gen_ekf_golden.py -> tests/golden/ekf_*.npz) and numeric-Jacobian checks of H like the reference's own
 * getHNumeric
 * (:1327-1377).  Quirks kept on purpose: the sign of the (0,2) term of generate_Rbe/Reb
 * (common.h:178,193), the typo in get_Rbe_derivs (R_be_theta(1,1) written twice, (2,1) never set,
 * kalmanClass.cpp:1283-1284), PI = 3.1415926535.
 *
 * Also: the Mahalanobis gate (orc_gate.c).  The reference has NO gating code: the gate is
 * defined by this project from the reference's point-measurement model (kalmanClass.cpp:815-851) and
 * R_pts (:975-984); parity for it is GPU-vs-this-oracle only ("parity unpinned" w.r.t. the reference).
 */
#ifndef ORC_EKF_H
#define ORC_EKF_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ROBOT_STATE 12

/* PredictKF: x [N], P [N*N] row-major, both in/out */
void orc_ekf_predict(double *x, double *P, int N);

/* UpdateNCurvesAndPoints.  z: 8n+3 curve/pose measurements; A: n 4x4 split matrices (row-major);
 * curve_col[k] = curve_inds[curve_num[k]] (first state index of the curve); point_meas: 3*n_pts;
 * point_col[k] = point_inds[point_nums[k]].  Optional outputs (may be NULL): H (M*N), S (M*M), K (N*M). */
int orc_ekf_update_curves_points(double *x, double *P, int N, const double *z, int n, const double *A,
                                 const int *curve_col, const double *point_meas, const int *point_col,
                                 int n_pts, double *H_out, double *S_out, double *K_out);

/* UpdatePoints */
int orc_ekf_update_points(double *x, double *P, int N, const double *point_meas, const int *point_col,
                          int n_pts);

/* state augmentation (AddFirstStates :275-311, AddNewCurve :314-499, AddNewPoints :502-666) on caller-owned
 * x / P with room for the grown state; P row-major with leading dimension ld; return the new dimension */
void orc_ekf_add_first_states(double *x, double *P, int ld, const double *z19);
int orc_ekf_add_new_curve(double *x, double *P, int ld, int N, const double *z8, const double *A);
int orc_ekf_add_new_points(double *x, double *P, int ld, int N, const double *meas, int n_pts);

/* helpers exported for unit tests */
void orc_ekf_split_matrices(double t, double *A1, double *A2);
void orc_ekf_predicted_curve(const double *x, const double *A, int curve_col, double *zhat8);
void orc_ekf_rbe(double phi, double theta, double psi, double *R9);
void orc_ekf_rbe_derivs(double phi, double theta, double psi, double *Rphi, double *Rtheta, double *Rpsi);
int orc_pinv_svd(const double *A, double *Ainv, int n);

/* Mahalanobis gate (a14): for every measurement z_i (body xyz) the index of the point landmark with the
 * smallest d2 = nu^T S_j^-1 nu among those with d2 < gate (ties -> lowest j), or -1.
 * x, P as above; point_col[j] = first state index of landmark j.  d2_out may be NULL. */
int orc_gate_mahalanobis(const double *x, const double *P, int N, const int *point_col, int n_land,
                         const double *z, int nz, double gate, int *idx_out, double *d2_out, int nthreads);
/* per-landmark pack used by both sides: zhat (3) and the 6 unique entries of S^-1 */
void orc_gate_pack(const double *x, const double *P, int N, const int *point_col, int n_land, double *pack9);

/* the same from the pieces of P the gate reads (pose rows 6 x N, pose columns N x 6, one 3 x 3 diagonal block per
 * landmark), for maps whose dense P does not fit a host buffer */
void orc_gate_pack_pieces(const double *x, const double *pose_rows, const double *pose_cols, const double *diag3, int N,
                          const int *point_col, int n_land, double *pack9);
int orc_gate_mahalanobis_pieces(const double *x, const double *pose_rows, const double *pose_cols, const double *diag3,
                                int N, const int *point_col, int n_land, const double *z, int nz, double gate,
                                int *idx_out, double *d2_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
