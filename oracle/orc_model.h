/*
 * orc_model.h -- CPU oracle for the curve models the LM fits.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates (does not copy) the measurement model of the reference:
 *   m=19 stereo-pair model  : curveFittingClass.cpp:82-119 (modelfunc), :804-817
 *                             (cp_earth2image), :831-847 (earth2body), common.h:174-186
 *                             (generate_Rbe), curveFittingClass.cpp:1057-1073 (Bernstein eval),
 *                             camera constants common.h:70-102.
 *   m=8 image-space model   : curveFittingClass.cpp:1057-1073 alone.
 *   fit_curve driver pieces : measurement packing and t-parameterisation :435-511, LM options
 *                             :661-662, post-fit sign/angle normalisation :684-711.
 * The analytic Jacobian of the live m=19 model does not exist in the reference (its
 * jacmodelfunc, :122-405, differentiates an older model); the one here is derived
 *  and pinned with the reference's own dlevmar_chkjac (tests/test_oracle_lm.py).
 *
 * PARITY PINNING: the reference's curveFittingClass.cpp is compiled UNCHANGED by `make -C oracle ref` (OpenCV-1.x /
 * GSL stand-ins under oracle/shim/) into oracle/_ref/libcps_ref.so.  In ORC_MATH_LIBM mode this restatement is
 * bit-identical to that build: modelfunc on random parameters, the m = 8 Bernstein evaluation, fit_curve end to end
 * (packing, t, the live dlevmar_dif call, post-normalisation: parameters, info[0..9], return code), and
 * cleanup_and_group_edges frame by frame (tests/test_ref_pinning.py, fixtures tests/golden/ref_cps.npz).  The
 * initial guess agrees to 1e-13 (the reference's cvSVD is external to it).  ORC_MATH_DET is what the CUDA kernels
 * compute (fixed sin/cos sequence, pow as multiplies): within 4 ulp of the reference per model value.
 */
#ifndef ORC_MODEL_H
#define ORC_MODEL_H

#include "orc_lm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* camera constants, spelled with the same arithmetic as common.h:70-87,102 so that the
 * compiler folds them to the same doubles */
#define ORC_FX ((521.22 / 2.0 + 527.03 / 2.0) / 2.0)
#define ORC_FY ((518.32 / 2.0 + 524.73 / 2.0) / 2.0)
#define ORC_CX ((296.08 / 2.0 - 1.0 + 320.95 / 2.0 + 1.0) / 2.0)
#define ORC_CY ((236.17 / 2.0 + 228.81 / 2.0) / 2.0)
#define ORC_BASELINE 0.535
#define ORC_PI 3.1415926535 /* common.h:114 (truncated on purpose, as the reference has it) */

enum { ORC_MATH_DET = 0, ORC_MATH_LIBM = 1 };
enum { ORC_MODEL_STEREO19 = 19, ORC_MODEL_IMAGE8 = 8 };
enum { ORC_LM_DER = 0, ORC_LM_DIF = 1 };

/* per-fit "adata" (the reference's struct mydata, curveFittingClass.h:22-31, minus the GUI
 * pointer and the unused fields) */
typedef struct {
  const double *t; /* one parameter value per sample point */
  int count[4];    /* stereo19: points in [L c1, L c2, R c1, R c2]; image8: count[0] only */
  int math_mode;   /* ORC_MATH_DET or ORC_MATH_LIBM */
} orc_fit_data;

void orc_stereo19_func(double *p, double *hx, int m, int n, void *adata);
void orc_stereo19_jac(double *p, double *jac, int m, int n, void *adata);
void orc_image8_func(double *p, double *hx, int m, int n, void *adata);
void orc_image8_jac(double *p, double *jac, int m, int n, void *adata);

/* Bernstein weights of a cubic at t, in the reference's operation order */
void orc_bernstein3(double t, int math_mode, double w[4]);

/* t-parameterisation of one contour segment from its image rows (fit_curve:451-511) */
void orc_segment_t(const double *xy, int npts, double *t_out);

/* initial guess of fit_curve (curveFittingClass.cpp:516-651): stereo triangulation of the start / end points of
 * both curves, plane fit (normal = right singular vector of the smallest singular value of the centred 4x3
 * point matrix, the reference's cvSVD), rotation into the ground frame, straight-line control points.  The sign
 * of the singular vector is implementation-defined in the reference; `sign` = +1 orients the
 * normal so that d = n.c >= 0 (height parameter <= 0, the branch fit_curve's post-normalisation maps everything
 * to), -1 gives the opposite orientation.  meas / counts as in orc_fit_batch. */
void orc_stereo19_init_guess(const double *meas, const int *counts, double *p19, int math_mode, int sign);

/* cleanup_and_group_edges (curveFittingClass.cpp:1207-1418): kept index ranges of the four edge sequences of a
 * frame; returns 0 where the reference returns false */
int orc_cleanup_edges(const int *pts, const long long *seg_off, int reset_data_assoc, const double *tracked_endpt_y,
                      int range[4][2]);

/* post-fit normalisation of the 19 parameters (fit_curve:684-711) */
void orc_stereo19_postfit(double *p);

/*
 * Batched driver: B independent fits, laid out exactly like the product's batched C ABI
 * (include/cps_lm.h) so the parity tests can hand both the same buffers.
 *   meas     : concatenated (x,y) samples, fit f occupying doubles [off[f], off[f+1])
 *   t        : concatenated parameter values, fit f occupying [off[f]/2, off[f+1]/2)
 *   counts   : B x 4 (stereo19) or B x 1 (image8) segment sizes in points
 * Returns 0, or -1 on a bad argument.  nthreads > 1 uses a pthread work queue over fits.
 */
int orc_fit_batch(int model, int variant, int math_mode, int B, double *p, const double *meas,
                  const double *t, const long long *off, const int *counts, int itmax,
                  const double *opts, double *info, int *ret, int *extra /* B x 2 or NULL */,
                  int nthreads);

/* Same driver with the LM core supplied by the caller: oracle/orc_refshim.c passes adapters around
 * the reference's own dlevmar_der/dlevmar_dif so that "reference levmar + restated model" runs on
 * exactly the same buffers (used for pinning and as bench.py's --impl reference arm). */
typedef int (*orc_lm_der_fn)(orc_func_t func, orc_func_t jacf, double *p, const double *x, int m, int n,
                             int itmax, const double *opts, double *info, void *adata,
                             orc_lm_extra *extra);
typedef int (*orc_lm_dif_fn)(orc_func_t func, double *p, const double *x, int m, int n, int itmax,
                             const double *opts, double *info, void *adata, orc_lm_extra *extra);
int orc_fit_batch_impl(orc_lm_der_fn der, orc_lm_dif_fn dif, int model, int variant, int math_mode,
                       int B, double *p, const double *meas, const double *t, const long long *off,
                       const int *counts, int itmax, const double *opts, double *info, int *ret,
                       int *extra, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
