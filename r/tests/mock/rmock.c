/* rmock.c — a miniature stand-in for the part of R's C API that r/src/shim.c uses, so the shim can be EXECUTED where R is
 * absent (r/tests/shim_driver.c). NOT R: vectors are heap blocks with a type, a length, optional dim and names; nothing is
 * garbage collected; Rf_error() longjmps to the driver; unif_rand() is a fixed splitmix64 stream. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC { int type; R_xlen_t len; int nrow, ncol; void* data; SEXP names; };

static struct SEXPREC nil_rec = {0, 0, 0, 0, NULL, NULL}, names_rec = {0, 0, 0, 0, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_rec;

jmp_buf rmock_error_jmp;
char rmock_error_msg[1024];
char rmock_warning_msg[1024];
int rmock_interrupt_after = -1;     /* >= 0: the n-th R_CheckUserInterrupt() call "interrupts" */
int rmock_interrupt_calls = 0;
int rmock_onintr_called = 0;
const R_CallMethodDef* rmock_routines = NULL;
static uint64_t rng_state = 0x1234567ull;

static size_t elt_size(int type)
{
    switch (type) { case REALSXP: return sizeof(double); case LGLSXP: case INTSXP: return sizeof(int);
                    case STRSXP: case VECSXP: return sizeof(SEXP); case CHARSXP: return 1; default: return 1; }
}

SEXP Rf_allocVector(int type, R_xlen_t len)
{
    SEXP s = (SEXP)calloc(1, sizeof *s);
    s->type = type; s->len = len; s->nrow = (int)len; s->ncol = 1;
    s->data = calloc(len > 0 ? (size_t)len : 1, elt_size(type));
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_xlengthgets(SEXP x, R_xlen_t len)
{
    SEXP s = Rf_allocVector(x->type, len);
    memcpy(s->data, x->data, (size_t)(len < x->len ? len : x->len) * elt_size(x->type));
    return s;
}
int Rf_nrows(SEXP x) { return x->nrow; }
int Rf_ncols(SEXP x) { return x->ncol; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
double* REAL(SEXP x) { return (double*)x->data; }
int* LOGICAL(SEXP x) { return (int*)x->data; }
int* INTEGER(SEXP x) { return (int*)x->data; }
int Rf_asInteger(SEXP x) { return x->type == REALSXP ? (int)REAL(x)[0] : INTEGER(x)[0]; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0]; }
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_mkChar(const char* c) { SEXP s = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(c) + 1); strcpy((char*)s->data, c); return s; }
SEXP Rf_setAttrib(SEXP x, SEXP sym, SEXP val) { if (sym == R_NamesSymbol) x->names = val; return val; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
const char* rmock_name(SEXP list, R_xlen_t i) { return list->names ? (const char*)((SEXP*)list->names->data)[i]->data : ""; }
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int n) { (void)n; }
void Rf_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(rmock_error_msg, sizeof rmock_error_msg, fmt, ap); va_end(ap);
    longjmp(rmock_error_jmp, 1);
}
void Rf_warning(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(rmock_warning_msg, sizeof rmock_warning_msg, fmt, ap); va_end(ap);
}
void Rprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
void Rf_onintr(void) { rmock_onintr_called = 1; }
static jmp_buf toplevel_jmp;
static int in_toplevel = 0;
void R_CheckUserInterrupt(void)
{
    if (rmock_interrupt_after >= 0 && rmock_interrupt_calls++ >= rmock_interrupt_after && in_toplevel) longjmp(toplevel_jmp, 1);
}
Rboolean R_ToplevelExec(void (*fn)(void*), void* data)
{
    in_toplevel = 1;
    if (setjmp(toplevel_jmp)) { in_toplevel = 0; return FALSE; }
    fn(data);
    in_toplevel = 0;
    return TRUE;
}
void GetRNGstate(void) {}
void PutRNGstate(void) {}
double unif_rand(void)
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
void rmock_seed(uint64_t s) { rng_state = s; }
int R_registerRoutines(DllInfo* d, const void* c, const R_CallMethodDef* call, const void* f, const void* e)
{
    (void)d; (void)c; (void)f; (void)e; rmock_routines = call; return 1;
}
int R_useDynamicSymbols(DllInfo* d, int v) { (void)d; (void)v; return 0; }
