/* shim.c — the `.Call` boundary of the NNLM R package re-pointed at libnnlm_b200 (C ABI: include/nnlm_b200.h).
 *
 * Replaces src/RcppExports.cpp of the reference (the Rcpp-generated marshalling, lines 9-65): same registered symbols
 * (`_NNLM_c_nnmf` with 17 arguments, `_NNLM_c_nnlm` with 9), same argument order and R types, same returned lists, so
 * R/RcppExports.R, R/nnmf.R and R/nnlm.R of the reference run unmodified on top of it.
 *
 * Plain C against R's own API (no Rcpp, no Armadillo): build with  R CMD SHLIB shim.c -L<dir> -lnnlm_b200 .
 * R is not present in the image this repository is developed in: the file is compile-checked against the declarations of
 * r/tests/mock/Rinternals.h, and EXECUTED on the GPU box against the miniature runtime r/tests/mock/rmock.c, which
 * implements just the R API calls used here (r/tests/shim_driver.c, tests/test_r_shim.py); see INTEGRATION.md.
 */
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "nnlm_b200.h"

/* Rcpp::checkUserInterrupt() (src/nnmf.cpp:111) without a longjmp across live device state: the core polls this between
 * iterations and unwinds normally; the shim then raises the R interrupt condition. */
static void chk_intr(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int interrupted(void* user) { (void)user; return R_ToplevelExec(chk_intr, NULL) == FALSE; }

/* verbose == 2 table (src/nnmf.cpp:100-104,155-156,194-198) goes to R's console, not to the process's stdout */
static void print_to_r(void* user, const char* text) { (void)user; Rprintf("%s", text); }

/* Knobs without a counterpart in the reference's R signature come from the environment, so R/nnmf.R stays untouched:
 *   NNLM_B200_PRECISION = auto | exact | fast      (include/nnlm_b200.h, default auto)
 *   NNLM_B200_GPUS      = N                        (shard one nnmf() call over N GPUs; read by the library itself)
 *   NNLM_B200_MKL_TRACE = all | final              (default all = the reference's behaviour)
 *   NNLM_B200_DEVICE    = CUDA ordinal             (default: current device) */
static void options_from_env(nnlm_options* opt)
{
    memset(opt, 0, sizeof *opt);
    opt->device = -1;
    const char* e = getenv("NNLM_B200_PRECISION");
    if (e && strcmp(e, "exact") == 0) opt->precision = NNLM_PREC_EXACT;
    else if (e && strcmp(e, "fast") == 0) opt->precision = NNLM_PREC_FAST;
    e = getenv("NNLM_B200_MKL_TRACE");
    if (e && strcmp(e, "final") == 0) opt->mkl_trace = 1;
    e = getenv("NNLM_B200_DEVICE");
    if (e && *e) opt->device = atoi(e);
    opt->print = print_to_r;
}

static const int32_t* mask_or_null(SEXP m) { return XLENGTH(m) > 0 ? (const int32_t*)LOGICAL(m) : NULL; }

/* c_nnmf(A, k, W, H, Wm, Hm, alpha, beta, max_iter, rel_tol, n_threads, verbose, show_warning, inner_max_iter,
 *        inner_rel_tol, method, trace)  — src/nnmf.cpp:4-8, src/RcppExports.cpp:30 */
SEXP _NNLM_c_nnmf(SEXP A, SEXP kS, SEXP W0, SEXP H0, SEXP Wm, SEXP Hm, SEXP alpha, SEXP beta, SEXP max_iterS, SEXP rel_tolS,
                  SEXP n_threadsS, SEXP verboseS, SEXP show_warningS, SEXP inner_max_iterS, SEXP inner_rel_tolS, SEXP methodS,
                  SEXP traceS)
{
    const R_xlen_t n = Rf_nrows(A), m = Rf_ncols(A);
    const int K = Rf_asInteger(kS);
    const unsigned max_iter = (unsigned)Rf_asInteger(max_iterS);
    unsigned trace = (unsigned)Rf_asInteger(traceS);
    if (trace < 1) trace = 1;                                                      /* src/nnmf.cpp:53 */
    const unsigned cap = (unsigned)((max_iter + trace - 1) / trace) + 1;           /* :54 */
    const int32_t* wm = mask_or_null(Wm);
    const int32_t* hm = mask_or_null(Hm);

    SEXP W = PROTECT(Rf_allocMatrix(REALSXP, (int)n, K));
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, K, (int)m));
    double *w = REAL(W), *h = REAL(H);
    /* explicit init for the core. Default init of src/nnmf.cpp:82-98: W.randu(k, n) * 0.01 (filled k-fastest, then used
     * transposed), masked entries zero; drawn from R's RNG like Rcpp::RNGScope (src/RcppExports.cpp:33). */
    GetRNGstate();
    if (XLENGTH(W0) > 0) {
        memcpy(w, REAL(W0), sizeof(double) * (size_t)n * K);
    } else {
        for (R_xlen_t i = 0; i < n; i++)
            for (int c = 0; c < K; c++) {
                const double u = 0.01 * unif_rand();
                w[i + n * c] = (wm && wm[i + n * c]) ? 0.0 : u;
            }
    }
    if (XLENGTH(H0) > 0) {
        memcpy(h, REAL(H0), sizeof(double) * (size_t)K * m);
    } else {
        for (R_xlen_t e = 0; e < (R_xlen_t)K * m; e++) {
            const double u = 0.01 * unif_rand();
            h[e] = (hm && hm[e]) ? 0.0 : u;
        }
    }
    PutRNGstate();

    SEXP mse = PROTECT(Rf_allocVector(REALSXP, cap)), mkl = PROTECT(Rf_allocVector(REALSXP, cap));
    SEXP tgt = PROTECT(Rf_allocVector(REALSXP, cap)), ep = PROTECT(Rf_allocVector(REALSXP, cap));
    uint32_t n_err = 0, n_iter = 0;
    int32_t converged = 1;
    char err[512];
    nnlm_options opt;
    options_from_env(&opt);
    const int rc = nnlm_nnmf(REAL(A), (int64_t)n, (int64_t)m, K, w, h, wm, hm, REAL(alpha), REAL(beta), max_iter,
                             Rf_asReal(rel_tolS), Rf_asInteger(n_threadsS), Rf_asInteger(verboseS),
                             (unsigned)Rf_asInteger(inner_max_iterS), Rf_asReal(inner_rel_tolS), Rf_asInteger(methodS), trace,
                             REAL(mse), REAL(mkl), REAL(tgt), REAL(ep), cap, &n_err, &n_iter, &converged,
                             interrupted, NULL, &opt, NULL, err, sizeof err);
    if (rc == NNLM_E_INTERRUPT) { UNPROTECT(6); Rf_onintr(); return R_NilValue; }
    if (rc != NNLM_OK) { UNPROTECT(6); Rf_error("%s", err); }
    if (Rf_asLogical(show_warningS) && !converged)
        Rf_warning("Target tolerance not reached. Try a larger max.iter.");         /* src/nnmf.cpp:208-209 */

    /* List(W, H, mse_error, mkl_error, target_error, average_epoch, n_iteration) — src/nnmf.cpp:211-219 */
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 7));
    static const char* nm[7] = {"W", "H", "mse_error", "mkl_error", "target_error", "average_epoch", "n_iteration"};
    SET_VECTOR_ELT(out, 0, W);
    SET_VECTOR_ELT(out, 1, H);
    SET_VECTOR_ELT(out, 2, Rf_xlengthgets(mse, n_err));                             /* :200-206 */
    SET_VECTOR_ELT(out, 3, Rf_xlengthgets(mkl, n_err));
    SET_VECTOR_ELT(out, 4, Rf_xlengthgets(tgt, n_err));
    SET_VECTOR_ELT(out, 5, Rf_xlengthgets(ep, n_err));
    SET_VECTOR_ELT(out, 6, Rf_ScalarReal((double)n_iter));
    for (int i = 0; i < 7; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(8);
    return out;
}

/* c_nnlm(x, y, alpha, mask, beta0, max_iter, rel_tol, n_threads, method) — src/nnlm.cpp:4-5, src/RcppExports.cpp:11 */
SEXP _NNLM_c_nnlm(SEXP x, SEXP y, SEXP alpha, SEXP mask, SEXP beta0, SEXP max_iterS, SEXP rel_tolS, SEXP n_threadsS, SEXP methodS)
{
    const R_xlen_t n = Rf_nrows(x), p = Rf_ncols(x), q = Rf_ncols(y);
    SEXP coef = PROTECT(Rf_allocMatrix(REALSXP, (int)p, (int)q));
    double* b = REAL(coef);
    if (XLENGTH(beta0) > 0) {
        memcpy(b, REAL(beta0), sizeof(double) * (size_t)p * q);
    } else {                                                                        /* beta.randu(), src/nnlm.cpp:38-39 */
        GetRNGstate();
        for (R_xlen_t e = 0; e < p * q; e++) b[e] = unif_rand();
        PutRNGstate();
    }
    int64_t nstep = 0;
    char err[512];
    nnlm_options opt;
    options_from_env(&opt);
    const int rc = nnlm_nnlm(REAL(x), REAL(y), (int64_t)n, (int64_t)p, (int64_t)q, b, mask_or_null(mask), REAL(alpha),
                             (unsigned)Rf_asInteger(max_iterS), Rf_asReal(rel_tolS), Rf_asInteger(n_threadsS),
                             Rf_asInteger(methodS), &nstep, &opt, NULL, err, sizeof err);
    if (rc != NNLM_OK) { UNPROTECT(1); Rf_error("%s", err); }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, coef);
    SET_VECTOR_ELT(out, 1, Rf_ScalarInteger((int)nstep));                            /* src/nnlm.cpp:49-52 */
    SET_STRING_ELT(names, 0, Rf_mkChar("coefficient"));
    SET_STRING_ELT(names, 1, Rf_mkChar("n_iteration"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(3);
    return out;
}

/* registration table: src/RcppExports.cpp:56-65 */
static const R_CallMethodDef CallEntries[] = {
    {"_NNLM_c_nnlm", (DL_FUNC)&_NNLM_c_nnlm, 9},
    {"_NNLM_c_nnmf", (DL_FUNC)&_NNLM_c_nnmf, 17},
    {NULL, NULL, 0}
};

void R_init_NNLM(DllInfo* dll)
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
