/*
 * cps_lm.h -- C ABI of the batched Bezier Levenberg-Marquardt fits (libcps.so, CUDA, sm_100a).
 *
 * Replaces, for B independent fits at a time, the reference's per-frame call
 *     dlevmar_dif(modelfunc, params, final_meas, m, num_meas, 1000, opts, info, NULL, NULL, &funcdata)
 * (curveFittingClass.cpp:681; ABI levmar-2.5/levmar.h:98-107) and its commented-out sibling
 * dlevmar_der (:682).  Device code cannot call host callbacks, so the model is named by an id instead
 * of a function pointer; everything else -- p in/out, opts[5], info[10], return value -- has the
 * reference's meaning (levmar.h:87-93, lm_core.c:396-409,422):
 *     info[0] ||e||^2 at the initial p     info[5] iterations
 *     info[1] ||e||^2 at the final p       info[6] stop reason 1..7
 *     info[2] ||J^T e||_inf                info[7] function evaluations
 *     info[3] ||dp||^2                     info[8] Jacobian evaluations
 *     info[4] mu / max diag(J^T J)         info[9] linear systems solved
 *     ret[f]  iterations, or -1 when the reason is 4 or 7 (LM_ERROR), or a CPS_FIT_* code below
 * There is no CPU fallback: without a CUDA device every entry point returns CPS_ERR_CUDA.
 */
#ifndef CPS_LM_H
#define CPS_LM_H

#ifdef __cplusplus
extern "C" {
#endif

/* model ids */
#define CPS_MODEL_STEREO19 19 /* modelfunc, curveFittingClass.cpp:82-119: 2 ground curves + theta, phi, z  */
#define CPS_MODEL_IMAGE8 8    /* Bernstein evaluation alone, curveFittingClass.cpp:1057-1073: one contour   */
/* LM variants */
#define CPS_LM_DER 0 /* analytic Jacobian, dlevmar_der lm_core.c:60-423 */
#define CPS_LM_DIF 1 /* secant / finite differences, dlevmar_dif lm_core.c:429-828 (the reference's live call) */

/* library status codes (return values of the entry points) */
#define CPS_OK 0
#define CPS_ERR_INVALID (-1)     /* bad argument                                            */
#define CPS_ERR_CUDA (-2)        /* CUDA runtime error or no device; see cps_last_error()   */
#define CPS_ERR_UNSUPPORTED (-3) /* callback / model the device path does not know          */
#define CPS_ERR_NOMEM (-4)

/* per-fit codes written to ret[] besides the reference's own values */
#define CPS_FIT_LM_ERROR (-1)  /* the reference's LM_ERROR: n < m, or stop reason 4 / 7 */
#define CPS_FIT_TOO_LARGE (-2) /* more measurements than n_max; Stereo19 dlevmar_der with the path FORCED to the
                                * warp-per-fit kernel (cps_lm_set_path MONOLITHIC): more than 1 280 measurements
                                * (the default path sends such batches to the staged kernels, which have no limit) */
#define CPS_FIT_BAD_COUNTS (-3)/* segment sizes do not add up to the measurement count  */
#define CPS_FIT_SKIPPED (-4)   /* cps_fit_frames_batched: cleanup_and_group_edges returned false for the frame */

/* integer pixel coordinates of cps_fit_frames_batched (CvPoint is two ints; a 320 x 240 image fits int16) */
#define CPS_PTS_INT32 0
#define CPS_PTS_INT16 1

#define CPS_LM_INFO_SZ 10 /* LM_INFO_SZ, levmar.h:85 */
#define CPS_LM_OPTS_SZ 5  /* LM_OPTS_SZ, levmar.h:84 */

typedef struct cps_lm_ctx cps_lm_ctx; /* owns a stream, the fit queue and the device workspace; one per thread */

int cps_lm_create(cps_lm_ctx **out, int device);
void cps_lm_destroy(cps_lm_ctx *ctx);
const char *cps_last_error(void);

/*
 * B fits with HOST buffers (copies in, kernel, copies out, synchronous).
 *   p      [B][m]     in: initial parameters, out: fitted parameters
 *   meas   [off[B]]   concatenated (x,y) contour samples; fit f owns doubles off[f]..off[f+1];
 *                     Stereo19 segment order [left c1 | left c2 | right c1 | right c2] (fit_curve :435-489)
 *   t      [off[B]/2] curve parameter of every sample, or NULL: derive it from the image rows exactly
 *                     like fit_curve (:451-511)
 *   counts [B][4] (Stereo19) or [B][1] (Image8): points per segment
 *   opts   [5] {tau, eps1, eps2, eps3, delta} or NULL for the reference defaults (levmar.h:87-93);
 *          delta < 0 selects central differences (dif only)
 *   info   [B][10], ret [B], extra [B][2] = (J^T J rebuilds, secant updates) or NULL
 */
int cps_dlevmar_der_batched(cps_lm_ctx *ctx, int model, int B, double *p, const double *meas, const double *t,
                            const long long *off, const int *counts, int itmax, const double *opts,
                            double *info, int *ret, int *extra);
int cps_dlevmar_dif_batched(cps_lm_ctx *ctx, int model, int B, double *p, const double *meas, const double *t,
                            const long long *off, const int *counts, int itmax, const double *opts,
                            double *info, int *ret, int *extra);

/*
 * The same fits with the samples as INTEGER pixel coordinates -- what the reference's detector produces (CvPoint
 * contours, featureDetector.cpp:213-288, packed by fit_curve :435-489) -- in the measurement order and with the `off`
 * (in measurements) and `counts` of the entries above.  The samples cross PCIe as 4 (CPS_PTS_INT32) or 2
 * (CPS_PTS_INT16) bytes per coordinate instead of 8 and become doubles on the device (exactly: a pixel coordinate is a
 * double); t is derived from the image rows as fit_curve does.  variant = CPS_LM_DER / CPS_LM_DIF.
 */
int cps_dlevmar_batched_pix(cps_lm_ctx *ctx, int model, int variant, int B, double *p, const void *pix, int pix_type,
                            const long long *off, const int *counts, int itmax, const double *opts, double *info,
                            int *ret, int *extra);

/*
 * Same with DEVICE buffers already resident in HBM, on `stream` (a cudaStream_t passed as void*; NULL = the CUDA
 * default stream, as in any CUDA call).  n_max = an even upper bound on off[f+1]-off[f].  The monolithic shape (see
 * cps_lm_set_path) only enqueues work and returns; the staged shape returns when the fits are done (its host loop
 * reads the number of fits still iterating after every round).  Either way the outputs are ordered on `stream`.
 * A context owns one fit queue and one work area: calls made on different streams are ordered one after the other (the
 * later call's stream waits for the earlier call's kernels); use one context per stream for concurrent batches.
 */
int cps_lm_fit_batched_dev(cps_lm_ctx *ctx, int model, int variant, int B, double *d_p, const double *d_meas,
                           const double *d_t, const long long *d_off, const int *d_counts, int n_max, int itmax,
                           const double *opts, double *d_info, int *d_ret, int *d_extra, void *stream);
int cps_lm_sync(cps_lm_ctx *ctx);
/*
 * Asynchronous device entry.  With on != 0, cps_lm_fit_batched_dev only queues the call and returns, for EVERY
 * execution shape: a worker thread owned by the context runs it (the staged shapes' round loop included) on a private
 * stream that first waits for the work already enqueued on the caller's `stream`, and `stream` itself is made to wait
 * on the device (cuStreamWaitValue32 on a sequence number the worker publishes) until the call has finished -- so the
 * caller's next kernels, copies and events on `stream` are ordered after the fits without the host ever blocking.
 * Calls run one at a time in the order they were made.  An error of a queued call is returned by the next
 * cps_lm_sync (which also waits for the queue to drain).  Off by default (the call then returns when the staged rounds
 * are done, see above).
 */
int cps_lm_set_async(cps_lm_ctx *ctx, int on);

/*
 * fit_curve's initial guess for B Stereo19 fits (curveFittingClass.cpp:516-651): stereo triangulation of the start /
 * end points of both curves, plane fit, rotation into the ground frame, straight-line control points; writes
 * p[B][19], the `params` the reference hands to levmar.  The reference's plane normal comes from cvSVD and its sign
 * is implementation-defined; here it is oriented so that the height parameter is <= 0 (the branch fit_curve's own
 * post-normalisation :684-702 maps every fit to).  Fits with an empty segment get NaN parameters.
 */
int cps_fit_curve_init_batched(cps_lm_ctx *ctx, int B, const double *meas, const long long *off, const int *counts,
                               double *p_out);
int cps_fit_curve_init_batched_dev(cps_lm_ctx *ctx, int B, const double *d_meas, const long long *d_off,
                                   const int *d_counts, double *d_p, void *stream);

/*
 * CurveFittingClass::cleanup_and_group_edges (curveFittingClass.cpp:1207-1418) for F frames: trims the four integer
 * edge sequences of every frame (left image curves 0/1, right image curves 0/1, each bottom-to-top) so that the two
 * views of a curve cover the same rows, no curve spans more than LENGTH_EACH_CURVE = 100 rows, nothing extends above
 * the tracked end point (unless reset_data_assoc), and both curves of an image end on the same row; then packs the
 * kept points as doubles in the fit's measurement order.
 *   pts            concatenated (x,y) int pairs (the reference's std::vector<CvPoint>)
 *   seg_off        [4F+1] point offsets, per frame [L0, L1, R0, R1]
 *   tracked_endpt_y[F][2]
 *   meas_out       room for 2 * (number of input points) doubles; off_out [F+1]; counts_out [F][4] -- exactly the
 *                  meas / off / counts arguments of cps_dlevmar_*_batched
 *   valid_out      [F] 1, or 0 where the reference returns false (such frames get zero counts)
 */
int cps_cleanup_and_group_edges_batched(cps_lm_ctx *ctx, int F, const int *pts, const long long *seg_off,
                                        int reset_data_assoc, const double *tracked_endpt_y, double *meas_out,
                                        long long *off_out, int *counts_out, int *valid_out);
/*
 * The per-frame sequence of main.cpp:425-445 as ONE call for F frames -- cleanup_and_group_edges
 * (curveFittingClass.cpp:1207-1418), then fit_curve (:408-724): measurement packing, t-parameterisation, initial
 * guess by stereo triangulation + plane fit (the reference ignores `params` on entry), the LM fit (variant
 * CPS_LM_DIF = the reference's live dlevmar_dif call :681, or CPS_LM_DER with the derived Jacobian) and the sign /
 * angle post-normalisation (:684-711).  The integer contours cross PCIe once (8 or 4 bytes per point instead of
 * 16 bytes of doubles); every stage runs on the device and nothing returns to the host in between.
 *   pts            concatenated (x,y) pixel pairs, int32 (CPS_PTS_INT32: the reference's CvPoint) or int16
 *   seg_off        [4F+1] point offsets, per frame [L0, L1, R0, R1]; tracked_endpt_y [F][2]
 *   itmax, opts    as dlevmar_*; the reference passes 1000 and {1e-10,1e-10,1e-10,1e-20,1e-5}
 *   p_out          [F][19] fitted, post-normalised parameters
 *   info / ret     [F][10] / [F] as cps_dlevmar_*_batched; ret = CPS_FIT_SKIPPED where the trimming rejected the frame
 *   valid          [F] the bool cleanup_and_group_edges returns; fitting_error [F] = info[1] (may be NULL)
 */
int cps_fit_frames_batched(cps_lm_ctx *ctx, int variant, int F, const void *pts, int pts_type, const long long *seg_off,
                           int reset_data_assoc, const double *tracked_endpt_y, int itmax, const double *opts,
                           double *p_out, double *info, int *ret, int *valid, double *fitting_error);
/*
 * Execution shape of dlevmar_der on Stereo19 batches.  Both shapes perform the reference's operation sequence
 * and give bit-identical results; they differ in how the work is spread over the GPU:
 *   MONOLITHIC  one persistent kernel, a warp per fit (best for a frame's worth of fits: one launch)
 *   STAGED      rounds of three phase kernels with a thread per fit in the Jacobian / J^T J phase (best for
 *               batches of tens of thousands of fits)
 *   AUTO        (default) STAGED from 8 192 fits per call upwards (the measured crossover on B200 is ~6 k)
 */
#define CPS_LM_PATH_AUTO 0
#define CPS_LM_PATH_MONOLITHIC 1
#define CPS_LM_PATH_STAGED 2
int cps_lm_set_path(cps_lm_ctx *ctx, int path);
/*
 * Arithmetic of the staged dlevmar_der kernels (Stereo19 batches that take the staged shape; everything else is always
 * exact).  CPS_ARITH_EXACT (default): the reference's floating-point operation sequence, every multiply and add issued
 * separately -- results equal the reference levmar's bit for bit.  CPS_ARITH_FMA: the Jacobian / J^T J and evaluation
 * kernels use fused multiply-adds (half the FP64 instructions in the sums); results then differ from the reference's
 * in the last bits of every sum, i.e. the way two CPU builds of the reference differ from each other (same stop reason
 * and ||e||^2 within 1e-10 on every fit, parameters within 1e-9 on about 95 % of the fits: bench.py extras.lm_der_fma,
 * tests/test_lm_staged_gpu.py).  The last fits of a batch (the warp-per-fit tail kernel) and the solve stay exact.
 */
#define CPS_ARITH_EXACT 0
#define CPS_ARITH_FMA 1
int cps_lm_set_arith(cps_lm_ctx *ctx, int arith);
/*
 * Per-kernel timing of the staged shape: with profiling on, every launch of its three kernels is bracketed by CUDA
 * events on the launching stream; cps_lm_get_profile returns the accumulated milliseconds and launch counts of
 * {solve, eval, jac, tail} (and clears them when reset != 0); tail is the warp-per-fit kernel that takes the last
 * fits of a batch over from the rounds.
 */
int cps_lm_set_profiling(cps_lm_ctx *ctx, int on);
int cps_lm_get_profile(cps_lm_ctx *ctx, double *ms4, long long *launches4, int reset);
/*
 * Work behind those times (accumulated while profiling is on): the Jacobian evaluations staged_jac_kernel performed
 * (a fit counts once per accepted step it computed there) and the fits the tail kernel took over.
 */
int cps_lm_get_profile_work(cps_lm_ctx *ctx, long long *jac_evals, long long *tail_fits, int reset);
/*
 * The staged dlevmar_der solves (J^T J + mu I) dp = J^T e on the block-arrow form of the Stereo19 matrix (the two
 * 8 x 8 blocks of structural zeros are not stored) and hands a system to the dense solve when the reference's scaled
 * pivot search would leave a curve's own rows or a factor is not finite (Axb_core.c:1044-1132 is followed either
 * way).  Number of systems of the last staged call that took that detour (0 on every batch measured so far).
 */
int cps_lm_get_solve_handovers(cps_lm_ctx *ctx, long long *systems);
/* geometry of the last launch and the number of kernel launches issued through this context */
int cps_lm_launch_stats(cps_lm_ctx *ctx, int *grid, int *block, int *smem_bytes, long long *launches);

/*
 * Same-signature shims for code that calls levmar one fit at a time (curveFittingClass.cpp:679-682).
 * `func` must be cps_modelfunc (m=19) or cps_modelfunc_image8 (m=8) and, for _der, `jacf` must be
 * cps_jacmodelfunc / cps_jacmodelfunc_image8: these are tokens that select the device model, they are
 * never executed.  Any other callback returns CPS_ERR_UNSUPPORTED (there is no CPU path).  `adata`
 * points to the reference's struct mydata (curveFittingClass.h:22-31), restated below.  `work` and
 * `covar` must be NULL (the reference passes NULL for both).
 */
struct cps_mydata {
  double *orig_meas;
  double *t;
  int *curve_num;
  int num_left1;
  int num_right1;
  int num_left2;
  int num_right2;
  void *disp;
};
/*
 * Registers a host function pointer as a NAME for a device model, so that existing code which passes its own model
 * function -- the reference's `modelfunc` (curveFittingClass.cpp:82, declared curveFittingClass.h:37) -- keeps
 * compiling and linking unchanged: cps_register_model(modelfunc, CPS_MODEL_STEREO19) once at start-up, after which
 * dlevmar_dif(modelfunc, ...) / cps_dlevmar_dif(modelfunc, ...) run the Stereo19 device model.  The pointer is only
 * compared, never called.  The derived analytic Jacobian has no counterpart in the reference (its jacmodelfunc is
 * stale); cps_register_jacobian names it for dlevmar_der callers.
 * libcps_levmar.so (csrc/cps_levmar_shim.c) exports dlevmar_der / dlevmar_dif under levmar's own names
 * (levmar.h:98-107) on top of these entry points: relink against it instead of liblevmar.a.
 */
int cps_register_model(void (*func)(double *p, double *hx, int m, int n, void *adata), int model_id);
int cps_register_jacobian(void (*jacf)(double *p, double *j, int m, int n, void *adata), int model_id);
void cps_modelfunc(double *p, double *hx, int m, int n, void *adata);
void cps_jacmodelfunc(double *p, double *jac, int m, int n, void *adata);
void cps_modelfunc_image8(double *p, double *hx, int m, int n, void *adata);
void cps_jacmodelfunc_image8(double *p, double *jac, int m, int n, void *adata);
int cps_dlevmar_der(void (*func)(double *, double *, int, int, void *),
                    void (*jacf)(double *, double *, int, int, void *), double *p, double *x, int m, int n,
                    int itmax, double *opts, double *info, double *work, double *covar, void *adata);
int cps_dlevmar_dif(void (*func)(double *, double *, int, int, void *), double *p, double *x, int m, int n,
                    int itmax, double *opts, double *info, double *work, double *covar, void *adata);

/*
 * Closest point on a planar cubic for Q (curve, point) queries: closest_bezier_pt(p, pt, nearest_pt, out),
 * curveFittingClass.cpp:936-1043 (+ control2coeffs :737-749 when is_control_points != 0); callers
 * dataAssocClass.cpp:272,366.  curves: Q x 8 doubles, or 1 x 8 shared by every query when shared_curve != 0.
 * out[q] = {distance, t}; nearest[q] = the point (0,0 when no real root qualifies: the reference leaves it untouched);
 * nroots[q] (may be NULL) = number of roots the solver returned, -1 if it did not converge.  The reference's root
 * finder is GSL's gsl_poly_complex_solve; its algorithm (balanced companion matrix + Francis QR) is restated.
 */
int cps_closest_bezier_pt_batched(int device, int Q, const double *curves, int is_control_points, int shared_curve,
                                  const double *pts, double *nearest, double *out, int *nroots);

/* FP64 roofline denominators measured on the device: a DFMA stream and a DMUL+DADD stream (TFLOP/s,
 * a multiply-add pair counted as 2 flop in both). */
int cps_measure_fp64_peak(int device, double *tflops_fma, double *tflops_muladd);
/* FP64 tensor-core peak (mma.sync.m8n8k4.f64 stream), TFLOP/s */
int cps_measure_dmma_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
