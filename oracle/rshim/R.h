/* Test infrastructure: minimal stand-in for R's <R.h>, just enough to compile the
 * reference EBEN C sources (/root/reference/EBEN_orig/src/*.c) UNMODIFIED into
 * oracle/_ref/libeben_ref.so without an R installation.  Not part of the product. */
#ifndef PAREBEN_ORACLE_RSHIM_R_H
#define PAREBEN_ORACLE_RSHIM_R_H
#include <stdlib.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>

/* R_alloc memory is reclaimed by R at the end of .C(); the only user on the path is the
 * unused lambda-max helper, so a plain (leaking) calloc is an adequate stand-in. */
static inline char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }
#define Calloc(n, t) ((t *)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(t)))
#define Free(p) do { free(p); (p) = NULL; } while (0)

/* Progress text goes to stderr, and only when EBEN_REF_VERBOSE is set (the reference prints
 * "out of Memory" diagnostics even with verbose = 0). */
static inline void Rprintf(const char *fmt, ...) {
    static int on = -1;
    if (on < 0) on = getenv("EBEN_REF_VERBOSE") != NULL;
    if (!on) return;
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
#define F77_CALL(x) scipy_##x##_
#define F77_NAME(x) scipy_##x##_
#endif
