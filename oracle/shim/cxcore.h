/*
 * oracle/shim/cxcore.h -- a minimal stand-in for the slice of the OpenCV 1.x C API that the reference's
 * kalmanClass.cpp / curveFittingClass.cpp touch.  TEST INFRASTRUCTURE ONLY: it exists so that
 * `make -C oracle ref` can compile those reference sources UNCHANGED, from where they lie under
 * /root/reference, into oracle/_ref/libcps_ref.so and the oracle restatements can be pinned against the
 * reference's own code.  Nothing here is OpenCV code; the functions are written from the documented
 * semantics of the calls (dst = a*b, dst = a^T, pseudo-inverse through an SVD, ...).
 *
 * Calls the reference makes (grep of the two sources): cvCreateMat cvReleaseMat cvCloneMat cvSetZero cvmGet
 * cvmSet cvMatMul cvSub cvAdd cvTranspose cvInvert(CV_SVD) cvAddWeighted cvNorm cvCopy cvSVD cvDotProduct.
 * Types that only appear in declarations (IplImage, CvVideoWriter, CvCapture ...) are opaque.
 */
#ifndef CPS_ORACLE_SHIM_CXCORE_H
#define CPS_ORACLE_SHIM_CXCORE_H

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#include <algorithm>
#include <deque>
#include <iostream>
#include <string>
#include <vector>
#endif

typedef unsigned char uchar;

#define CV_64FC1 6
#define CV_32FC1 5
#define CV_SVD 1
#define CV_LU 0
#define CV_L2 4
#define CV_PI 3.1415926535897932384626433832795

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

typedef struct CvMat {
  int type;
  int step; /* bytes per row */
  int *refcount;
  int hdr_refcount;
  union {
    uchar *ptr;
    double *db;
  } data;
  int rows;
  int cols;
} CvMat;

typedef void CvArr;
typedef struct CvPoint { int x, y; } CvPoint;
typedef struct CvPoint2D32f { float x, y; } CvPoint2D32f;
typedef struct CvPoint3D32f { float x, y, z; } CvPoint3D32f;
typedef struct CvSize { int width, height; } CvSize;
typedef struct CvRect { int x, y, width, height; } CvRect;
typedef struct _IplImage IplImage;       /* opaque: declarations only */
typedef struct CvVideoWriter CvVideoWriter;
typedef struct CvCapture CvCapture;

#ifdef __cplusplus
extern "C" {
#endif

CvMat *cvCreateMat(int rows, int cols, int type);
void cvReleaseMat(CvMat **m);
CvMat *cvCloneMat(const CvMat *m);
void cvSetZero(CvMat *m);
void cvCopy(const CvMat *src, CvMat *dst);
/* dst = alpha * a * b + beta * c (c may be NULL); dst may alias a or b */
void cvGEMM(const CvMat *a, const CvMat *b, double alpha, const CvMat *c, double beta, CvMat *dst, int flags);
void cvTranspose(const CvMat *src, CvMat *dst); /* dst may alias src for square matrices */
void cvAddWeighted(const CvMat *a, double alpha, const CvMat *b, double beta, double gamma, CvMat *dst);
/* ||a||_2 over all entries (a2 and mask must be NULL: the only form the reference uses) */
double cvNorm(const CvMat *a, const CvMat *a2, int norm_type, const CvMat *mask);
double cvDotProduct(const CvMat *a, const CvMat *b);
/* CV_SVD: Moore-Penrose pseudo-inverse through the singular value decomposition; returns the inverse
 * condition number (0 when singular).  dst may alias src. */
double cvInvert(const CvMat *src, CvMat *dst, int method);
/* a = u diag(w) v^T with flags == 0: w descending, u rows x cols, v cols x cols (columns = right
 * singular vectors). */
void cvSVD(CvMat *a, CvMat *w, CvMat *u, CvMat *v, int flags);
IplImage *cvCreateImage(CvSize size, int depth, int channels); /* never called on the tested path */
void cps_shim_add(const CvMat *a, const CvMat *b, CvMat *dst, double sb);

#ifdef __cplusplus
}
#endif

static inline double cvmGet(const CvMat *m, int row, int col) { return m->data.db[(size_t)row * m->cols + col]; }
static inline void cvmSet(CvMat *m, int row, int col, double v) { m->data.db[(size_t)row * m->cols + col] = v; }
#define cvMatMul(a, b, dst) cvGEMM((a), (b), 1.0, NULL, 0.0, (dst), 0)
#define cvSub(a, b, dst) cps_shim_add((a), (b), (dst), -1.0)
#define cvAdd(a, b, dst) cps_shim_add((a), (b), (dst), 1.0)

#endif
