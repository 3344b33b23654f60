/*
 * cps_ekf.h -- C ABI of the EKF predict / measurement update and of the Mahalanobis gate (libcps.so).
 *
 * Replaces the arithmetic of the reference's KalmanFilter (kalmanClass.h:40-163) for the hot path:
 *   cps_ekf_predict               <- KalmanFilter::PredictKF              kalmanClass.cpp:161-272
 *   cps_ekf_update_curves_points  <- KalmanFilter::UpdateNCurvesAndPoints kalmanClass.cpp:669-945
 *   cps_ekf_update_points         <- KalmanFilter::UpdatePoints           kalmanClass.cpp:948-1126
 *   cps_ekf_get_state             <- KalmanFilter::getState               kalmanClass.cpp:1138-1141
 *   cps_ekf_split_matrices        <- KalmanFilter::GetSplitMatrices       kalmanClass.cpp:1144-1159
 *   cps_gate_mahalanobis          <- (no reference code; defined from the point measurement model
 *                                    kalmanClass.cpp:815-851 and R_pts :975-984 -- SURVEY.md D4 / a14)
 * The state x (N) and the dense joint covariance P (N x N, row-major) live in device memory between
 * calls (the reference keeps them in the public members `x`, `P`, kalmanClass.h:152-153); hosts read them
 * back with the get_* calls.  State layout as in the reference (kalmanClass.h:159-162): x[0..2] = T,
 * x[3..5] = (phi, theta, psi), x[6..8] = v, x[9..11] = omega, then 8 states per curve [x0..x3, y0..y3]
 * at curve_inds[k] and 3 per point landmark at point_inds[j].
 *
 * All entry points return CPS_OK (0) or a negative CPS_ERR_* code (cps_lm.h); there is no CPU path.
 * Calls on one handle are ordered on the handle's stream and are asynchronous unless they return data.
 */
#ifndef CPS_EKF_H
#define CPS_EKF_H

#include "cps_lm.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CPS_ROBOT_STATE_SIZE 12 /* ROBOT_STATE_SIZE, kalmanClass.h:36 */
#define CPS_EKF_MAX_MEAS 104    /* largest measurement vector of one update (the reference's live M is <= 35) */

typedef struct cps_ekf cps_ekf;

/* capacity = the largest state dimension the handle will ever hold (P is allocated capacity^2 doubles) */
int cps_ekf_create(cps_ekf **out, int device, int capacity);
void cps_ekf_destroy(cps_ekf *h);
int cps_ekf_sync(cps_ekf *h);

/* Load a state: x [N]; P [N*N] row-major host matrix or NULL (then P = 0); the landmark index tables
 * (the reference's curve_inds / point_inds vectors). */
int cps_ekf_set_state(cps_ekf *h, const double *x, const double *P, int N, const int *curve_inds, int num_curves,
                      const int *point_inds, int num_points);
/* P = scale * G G^T + diag * I formed on the device from a host G [N x r] (benchmark-size covariances that
 * do not fit a host buffer). */
int cps_ekf_set_cov_lowrank(cps_ekf *h, const double *G, int r, double scale, double diag);

int cps_ekf_dims(cps_ekf *h, int *N, int *num_curves, int *num_points);
int cps_ekf_get_state(cps_ekf *h, double *x_out);                                         /* [N]      */
/* The full symmetric matrix / any block of it.  Between updates only the 64 x 64 tiles on and above the diagonal tiles
   are kept current on the device (an update then moves 8 N^2 bytes instead of 12 N^2); a read that reaches below them
   gathers a small block (up to 2^20 entries) from the upper copies, and for a larger one first rebuilds the mirror images
   (one extra pass over P on the handle's stream, only when asked). */
int cps_ekf_get_cov(cps_ekf *h, double *P_out);                                           /* [N*N]    */
int cps_ekf_get_cov_block(cps_ekf *h, int r0, int c0, int nr, int nc, double *out);       /* [nr*nc]  */

int cps_ekf_predict(cps_ekf *h);

/* State augmentation, in place inside the capacity (the reference re-allocates and copies P per call):
 *   cps_ekf_add_first_states <- KalmanFilter::AddFirstStates kalmanClass.cpp:275-311  (z19 = 16 control-point
 *                               values of the first two curves, then height, phi, theta; N becomes 28)
 *   cps_ekf_add_new_curve    <- KalmanFilter::AddNewCurve    kalmanClass.cpp:314-499  (z8 body-frame control points,
 *                               A the 4x4 split matrix; appends 8 states, curve_inds grows by one)
 *   cps_ekf_add_new_points   <- KalmanFilter::AddNewPoints   kalmanClass.cpp:502-666  (3*n_pts body-frame xyz;
 *                               appends 3 states per point)
 * CPS_ERR_NOMEM when the capacity given to cps_ekf_create is exhausted. */
int cps_ekf_add_first_states(cps_ekf *h, const double *z19);
int cps_ekf_add_new_curve(cps_ekf *h, const double *z8, const double *A);
int cps_ekf_add_new_points(cps_ekf *h, const double *meas, int n_pts);
/* copies of the reference's curve_inds / point_inds vectors (sizes from cps_ekf_dims) */
int cps_ekf_get_index_tables(cps_ekf *h, int *curve_inds, int *point_inds);

/* z: 8n+3 values = per curve [x0..x3, y0..y3] in the body frame, then height, phi, theta (main.cpp:959-961);
 * A: n row-major 4x4 split matrices; curve_num[k] indexes curve_inds; point_meas: 3*n_pts body-frame xyz;
 * point_nums[k] indexes point_inds. */
int cps_ekf_update_curves_points(cps_ekf *h, const double *z, int n, const double *A, const int *curve_num,
                                 const double *point_meas, const int *point_nums, int n_pts);
int cps_ekf_update_points(cps_ekf *h, const double *point_meas, const int *point_nums, int n_pts);

/* device time (ms, CUDA events on the handle's stream) of the kernels of the last update / predict, the
 * time of its covariance kernel alone, and how many kernels the handle has launched so far */
int cps_ekf_last_timing(cps_ekf *h, float *total_ms, float *cov_kernel_ms, long long *launches);

/* de Casteljau split matrices of a cubic at t (pure host arithmetic, 32 doubles out) */
void cps_ekf_split_matrices(double t, double *A1, double *A2);

/* Mahalanobis gate against the point landmarks of the handle: for each measurement z_i (body-frame xyz)
 * idx_out[i] = argmin_j d2_ij subject to d2_ij < gate (ties: lowest j; none: -1), d2_out[i] = that d2
 * (+inf if none; may be NULL).  Host buffers; results are bit-identical to the CPU definition in oracle/. */
int cps_gate_mahalanobis(cps_ekf *h, const double *z, int nz, double gate, int *idx_out, double *d2_out);
/* device buffers (z: nz*3 doubles, idx: nz ints, d2: nz doubles or NULL), asynchronous on the handle's stream */
int cps_gate_mahalanobis_dev(cps_ekf *h, const double *d_z, int nz, double gate, int *d_idx, double *d_d2);
int cps_gate_last_timing(cps_ekf *h, float *pack_ms, float *sweep_ms);

#ifdef __cplusplus
}
#endif
#endif
