/* shim_driver.c — executes r/src/shim.c end to end without R: the miniature runtime r/tests/mock/rmock.c stands in for
 * libR, this program plays the part of R/nnmf.R + R/RcppExports.R (it looks the two routines up in the table R_init_NNLM
 * registers and `.Call`s them with the 17 / 9 arguments of the reference, R/nnmf.R:177-182, R/nnlm.R:119-120).
 *
 *   shim_driver nnmf <in.bin> <out.bin>     in:  int64 n, m, K, max_iter, trace, method, inner, has_init, has_na(unused); double rel_tol;
 *                                                double A[n*m] (+ W[n*K], H[K*m] if has_init)
 *                                           out: int64 n_err, n_iter(as double->int), warned; double W[n*K], H[K*m], mse[], mkl[], target[], epochs[]
 *   shim_driver nnlm <in.bin> <out.bin>     in:  int64 n, p, q, max_iter, method; double rel_tol; double x[n*p], y[n*q], beta0[p*q]
 *                                           out: int64 n_iteration; double coef[p*q]
 *   shim_driver interrupt                   an nnmf call whose 3rd interrupt poll fires: must unwind and call Rf_onintr()
 * tests/test_r_shim.py builds this with gcc on the GPU box and compares the outputs with the ctypes path. */
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <Rinternals.h>
#include <R_ext/Rdynload.h>

extern jmp_buf rmock_error_jmp;
extern char rmock_error_msg[], rmock_warning_msg[];
extern int rmock_interrupt_after, rmock_onintr_called;
extern const R_CallMethodDef* rmock_routines;
void R_init_NNLM(DllInfo*);
const char* rmock_name(SEXP, R_xlen_t);

typedef SEXP (*call17)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*call9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static DL_FUNC lookup(const char* name, int nargs)
{
    for (const R_CallMethodDef* r = rmock_routines; r && r->name; r++)
        if (strcmp(r->name, name) == 0 && r->numArgs == nargs) return r->fun;
    fprintf(stderr, "routine %s/%d is not registered\n", name, nargs);
    exit(2);
}
static SEXP lgl0(int nr, int nc) { return Rf_allocMatrix(LGLSXP, nr, nc); }
static void rd(FILE* f, void* p, size_t bytes) { if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); } }

static int run_nnmf(const char* in, const char* out)
{
    FILE* f = fopen(in, "rb");
    if (!f) { perror(in); return 2; }
    int64_t hd[9]; double rel_tol;
    rd(f, hd, sizeof hd); rd(f, &rel_tol, sizeof rel_tol);
    const int n = (int)hd[0], m = (int)hd[1], K = (int)hd[2];
    SEXP A = Rf_allocMatrix(REALSXP, n, m);
    rd(f, REAL(A), sizeof(double) * (size_t)n * m);
    SEXP W0, H0;
    if (hd[7]) {
        W0 = Rf_allocMatrix(REALSXP, n, K); H0 = Rf_allocMatrix(REALSXP, K, m);
        rd(f, REAL(W0), sizeof(double) * (size_t)n * K); rd(f, REAL(H0), sizeof(double) * (size_t)K * m);
    } else { W0 = Rf_allocMatrix(REALSXP, n, 0); H0 = Rf_allocMatrix(REALSXP, 0, m); }        /* R/misc.R: empty init */
    fclose(f);
    SEXP alpha = Rf_allocVector(REALSXP, 3), beta = Rf_allocVector(REALSXP, 3);
    call17 fn = (call17)lookup("_NNLM_c_nnmf", 17);
    if (setjmp(rmock_error_jmp)) { fprintf(stderr, "Rf_error: %s\n", rmock_error_msg); return 3; }
    SEXP r = fn(A, Rf_ScalarInteger(K), W0, H0, lgl0(n, 0), lgl0(0, m), alpha, beta, Rf_ScalarInteger((int)hd[3]),
                Rf_ScalarReal(rel_tol), Rf_ScalarInteger(1), Rf_ScalarInteger(0), Rf_ScalarInteger(1) /*show.warning*/,
                Rf_ScalarInteger((int)hd[6]), Rf_ScalarReal(1e-9), Rf_ScalarInteger((int)hd[5]), Rf_ScalarInteger((int)hd[4]));
    static const char* want[7] = {"W", "H", "mse_error", "mkl_error", "target_error", "average_epoch", "n_iteration"};
    if (XLENGTH(r) != 7) { fprintf(stderr, "result list has %ld elements\n", (long)XLENGTH(r)); return 4; }
    for (int i = 0; i < 7; i++) if (strcmp(rmock_name(r, i), want[i])) { fprintf(stderr, "element %d is named %s\n", i, rmock_name(r, i)); return 4; }
    SEXP W = VECTOR_ELT(r, 0), H = VECTOR_ELT(r, 1);
    if (Rf_nrows(W) != n || Rf_ncols(W) != K || Rf_nrows(H) != K || Rf_ncols(H) != m) { fprintf(stderr, "bad output dims\n"); return 4; }
    const int64_t ne = XLENGTH(VECTOR_ELT(r, 2));
    int64_t oh[3] = {ne, (int64_t)Rf_asReal(VECTOR_ELT(r, 6)), rmock_warning_msg[0] ? 1 : 0};
    FILE* g = fopen(out, "wb");
    fwrite(oh, sizeof oh, 1, g);
    fwrite(REAL(W), sizeof(double), (size_t)n * K, g); fwrite(REAL(H), sizeof(double), (size_t)K * m, g);
    for (int i = 2; i <= 5; i++) fwrite(REAL(VECTOR_ELT(r, i)), sizeof(double), (size_t)ne, g);
    fclose(g);
    if (rmock_warning_msg[0]) fprintf(stderr, "Rf_warning: %s\n", rmock_warning_msg);
    return 0;
}

static int run_nnlm(const char* in, const char* out)
{
    FILE* f = fopen(in, "rb");
    if (!f) { perror(in); return 2; }
    int64_t hd[5]; double rel_tol;
    rd(f, hd, sizeof hd); rd(f, &rel_tol, sizeof rel_tol);
    const int n = (int)hd[0], p = (int)hd[1], q = (int)hd[2];
    SEXP x = Rf_allocMatrix(REALSXP, n, p), y = Rf_allocMatrix(REALSXP, n, q), b0 = Rf_allocMatrix(REALSXP, p, q);
    rd(f, REAL(x), sizeof(double) * (size_t)n * p); rd(f, REAL(y), sizeof(double) * (size_t)n * q); rd(f, REAL(b0), sizeof(double) * (size_t)p * q);
    fclose(f);
    call9 fn = (call9)lookup("_NNLM_c_nnlm", 9);
    if (setjmp(rmock_error_jmp)) { fprintf(stderr, "Rf_error: %s\n", rmock_error_msg); return 3; }
    SEXP r = fn(x, y, Rf_allocVector(REALSXP, 3), lgl0(0, 0), b0, Rf_ScalarInteger((int)hd[3]), Rf_ScalarReal(rel_tol),
                Rf_ScalarInteger(1), Rf_ScalarInteger((int)hd[4]));
    if (XLENGTH(r) != 2 || strcmp(rmock_name(r, 0), "coefficient") || strcmp(rmock_name(r, 1), "n_iteration")) return 4;
    int64_t nit = Rf_asInteger(VECTOR_ELT(r, 1));
    FILE* g = fopen(out, "wb");
    fwrite(&nit, sizeof nit, 1, g);
    fwrite(REAL(VECTOR_ELT(r, 0)), sizeof(double), (size_t)p * q, g);
    fclose(g);
    return 0;
}

static int run_interrupt(void)
{
    const int n = 60, m = 40, K = 3;
    SEXP A = Rf_allocMatrix(REALSXP, n, m);
    for (int e = 0; e < n * m; e++) REAL(A)[e] = 1.0 + (double)(e % 7);
    call17 fn = (call17)lookup("_NNLM_c_nnmf", 17);
    rmock_interrupt_after = 2;
    if (setjmp(rmock_error_jmp)) { fprintf(stderr, "Rf_error: %s\n", rmock_error_msg); return 3; }
    SEXP r = fn(A, Rf_ScalarInteger(K), Rf_allocMatrix(REALSXP, n, 0), Rf_allocMatrix(REALSXP, 0, m), lgl0(n, 0), lgl0(0, m),
                Rf_allocVector(REALSXP, 3), Rf_allocVector(REALSXP, 3), Rf_ScalarInteger(50), Rf_ScalarReal(-1.0), Rf_ScalarInteger(1),
                Rf_ScalarInteger(0), Rf_ScalarInteger(0), Rf_ScalarInteger(10), Rf_ScalarReal(1e-9), Rf_ScalarInteger(1), Rf_ScalarInteger(1));
    if (r != R_NilValue || !rmock_onintr_called) { fprintf(stderr, "interrupt was not propagated\n"); return 5; }
    printf("interrupt propagated\n");
    return 0;
}

int main(int argc, char** argv)
{
    R_init_NNLM(NULL);
    if (argc >= 4 && !strcmp(argv[1], "nnmf")) return run_nnmf(argv[2], argv[3]);
    if (argc >= 4 && !strcmp(argv[1], "nnlm")) return run_nnlm(argv[2], argv[3]);
    if (argc >= 2 && !strcmp(argv[1], "interrupt")) return run_interrupt();
    fprintf(stderr, "usage: shim_driver nnmf|nnlm in out | interrupt\n");
    return 2;
}
