/* oracle R shim: prototypes of the BLAS/LAPACK routines the reference calls, bound to the
 * LP64 OpenBLAS bundled with scipy (symbols carry a scipy_ prefix).  Fortran hidden
 * string-length arguments are omitted exactly as R >= 3.6.2's headers allow for C callers. */
#ifndef PAREBEN_ORACLE_RSHIM_LAPACK_H
#define PAREBEN_ORACLE_RSHIM_LAPACK_H
#include "../R.h"
double scipy_ddot_(const int *n, const double *x, const int *incx, const double *y, const int *incy);
double scipy_dasum_(const int *n, const double *x, const int *incx);
void scipy_dcopy_(const int *n, const double *x, const int *incx, double *y, const int *incy);
void scipy_daxpy_(const int *n, const double *a, const double *x, const int *incx, double *y, const int *incy);
void scipy_dscal_(const int *n, const double *a, double *x, const int *incx);
void scipy_dgemv_(const char *trans, const int *m, const int *n, const double *alpha, const double *a,
                  const int *lda, const double *x, const int *incx, const double *beta, double *y, const int *incy);
void scipy_dgemm_(const char *ta, const char *tb, const int *m, const int *n, const int *k, const double *alpha,
                  const double *a, const int *lda, const double *b, const int *ldb, const double *beta,
                  double *c, const int *ldc);
void scipy_dpotrf_(const char *uplo, const int *n, double *a, const int *lda, int *info);
void scipy_dpotri_(const char *uplo, const int *n, double *a, const int *lda, int *info);
void scipy_dgelsy_(const int *m, const int *n, const int *nrhs, double *a, const int *lda, double *b,
                   const int *ldb, int *jpvt, const double *rcond, int *rank, double *work,
                   const int *lwork, int *info);
#endif
