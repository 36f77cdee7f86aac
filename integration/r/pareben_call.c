/* pareben_call.c -- the `.Call` shim an R package maintainer adds under parEBEN/src/.
 * Thin by design: it only converts SEXPs to the plain pointers of include/pareben.h.
 * Build inside the R package with  PKG_LIBS = -L<dir of libpareben.so> -lpareben
 * (cannot be compiled in the build container: R headers are absent; the C-ABI it calls is
 * exercised from Python in tests/). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "pareben.h"

/* .Call("pareben_cv_grid_call", BASIS, Target, foldId, nFolds, alpha, lambda, epis, prior, device)
 * returns list(fold_err = matrix[nFolds x nGrid], status = integer matrix, n_selected = integer matrix);
 * column g of fold_err holds the nFolds hold-out errors of grid row g, so as.vector(fold_err)
 * is already in Results.Detail order (grid row major, fold minor). */
SEXP pareben_cv_grid_call(SEXP BASIS, SEXP Target, SEXP foldId, SEXP nFolds, SEXP alpha, SEXP lambda,
                          SEXP epis, SEXP prior, SEXP device)
{
    SEXP dim = getAttrib(BASIS, R_DimSymbol);
    if (!isReal(BASIS) || isNull(dim) || LENGTH(dim) != 2) error("BASIS must be a double matrix");
    const int n = INTEGER(dim)[0], k = INTEGER(dim)[1];
    const int nf = asInteger(nFolds), ng = LENGTH(alpha);
    if (LENGTH(Target) != n || LENGTH(foldId) != n || LENGTH(lambda) != ng) error("argument lengths disagree");
    SEXP err = PROTECT(allocMatrix(REALSXP, nf, ng));
    SEXP st = PROTECT(allocMatrix(INTSXP, nf, ng));
    SEXP ns = PROTECT(allocMatrix(INTSXP, nf, ng));
    const int rc = pareben_cv_grid(REAL(BASIS), n, k, REAL(Target), INTEGER(foldId), nf, REAL(alpha), REAL(lambda),
                                   ng, asInteger(epis), asInteger(prior), asInteger(device), 0, 1,
                                   REAL(err), INTEGER(st), INTEGER(ns));
    if (rc != PAREBEN_OK) { UNPROTECT(3); error("pareben_cv_grid failed (%d): %s", rc, pareben_last_error()); }
    SEXP out = PROTECT(allocVector(VECSXP, 3)), nm = PROTECT(allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, err); SET_VECTOR_ELT(out, 1, st); SET_VECTOR_ELT(out, 2, ns);
    SET_STRING_ELT(nm, 0, mkChar("fold_err")); SET_STRING_ELT(nm, 1, mkChar("status")); SET_STRING_ELT(nm, 2, mkChar("n_selected"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"pareben_cv_grid_call", (DL_FUNC)&pareben_cv_grid_call, 9},
    {NULL, NULL, 0}
};

void R_init_parEBEN(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
