/*
 * orc_lm.h -- CPU oracle for the Levenberg-Marquardt hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a from-scratch restatement (not a copy) of the algorithm the reference runs in
 * levmar-2.5 (reference files cited per function in orc_lm.c).  It exists so that the CUDA
 * path can be checked bit-for-bit on a CPU; it is never linked into, imported by, or executed
 * from the product library (curvepointslam_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Pinning: orc_lm_der / orc_lm_dif are checked bit-for-bit against the reference's own
 * levmar-2.5 sources compiled unchanged into oracle/_ref/ (tests/test_oracle_lm.py), and
 * against the lmdemo known-answer table.
 */
#ifndef ORC_LM_H
#define ORC_LM_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_LM_INFO_SZ 10
#define ORC_LM_ERROR (-1)

typedef void (*orc_func_t)(double *p, double *hx, int m, int n, void *adata);

/* extra per-fit counters that do not fit the reference's 10 info slots */
typedef struct {
  int n_jtj;     /* number of times J^T J / J^T e were rebuilt */
  int n_broyden; /* number of rank-1 secant updates (dif only) */
} orc_lm_extra;

int orc_lm_der(orc_func_t func, orc_func_t jacf, double *p, const double *x, int m, int n,
               int itmax, const double *opts4, double *info, void *adata, orc_lm_extra *extra);

int orc_lm_dif(orc_func_t func, double *p, const double *x, int m, int n, int itmax,
               const double *opts5, double *info, void *adata, orc_lm_extra *extra);

/* building blocks, exported so the tests can pin them one by one */
double orc_residual_sqnorm(double *e, const double *x, const double *y, int n);
void orc_ata_blocked(const double *a, double *b, int n, int m);
int orc_lu_solve(const double *A, const double *B, double *x, int m);
void orc_fdjac_forward(orc_func_t func, double *p, const double *hx, double *hxx, double delta,
                       double *jac, int m, int n, void *adata);
void orc_fdjac_central(orc_func_t func, double *p, double *hxm, double *hxp, double delta,
                       double *jac, int m, int n, void *adata);

#ifdef __cplusplus
}
#endif
#endif
