/* Minimal stand-in for R's headers: the declarations r/src/shim.c needs where R is absent. NOT R.
 * r/tests/mock/rmock.c implements them (heap-allocated vectors with dim/names attributes) so the shim can be executed. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define CHARSXP 9
extern SEXP R_NilValue, R_NamesSymbol;
int Rf_nrows(SEXP); int Rf_ncols(SEXP); int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP);
R_xlen_t XLENGTH(SEXP); double* REAL(SEXP); int* LOGICAL(SEXP); int* INTEGER(SEXP);
SEXP Rf_allocMatrix(int, int, int); SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_xlengthgets(SEXP, R_xlen_t);
SEXP Rf_ScalarReal(double); SEXP Rf_ScalarInteger(int); SEXP Rf_mkChar(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...); void Rf_warning(const char*, ...); void Rf_onintr(void);
void Rprintf(const char*, ...);
void R_CheckUserInterrupt(void); Rboolean R_ToplevelExec(void (*)(void*), void*);
void GetRNGstate(void); void PutRNGstate(void); double unif_rand(void);
#endif
