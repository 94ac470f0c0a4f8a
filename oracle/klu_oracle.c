/*
 * klu_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic kvxopt's klu wrapper (reference src/C/klu.c) obtains from
 * SuiteSparse KLU 7.8.2 (pinned at reference .ci/config/versions.env:7; not in the reference tree, so
 * the published algorithm is restated: Davis & Palamadai Natarajan, "Algorithm 907: KLU", ACM TOMS
 * 2010): left-looking Gilbert-Peierls sparse LU with threshold partial pivoting that prefers the
 * diagonal (klu_defaults: tol = 0.001), on the row-scaled matrix (scale = 2: divide each row by its
 * max |entry|), followed by klu_refactor semantics (same pattern, same pivots, no pivot search).
 *
 *   oracle_klu_factor   <- klu_l_factor   (klu.c:142,337)  with a GIVEN row/column pre-ordering: the
 *                          BTF/AMD ordering is not arithmetic; tests pass the product's P,Q or identity.
 *                          The matrix is treated as ONE block (a BTF permutation changes the operation
 *                          order, not the solution).
 *   oracle_klu_refactor <- klu_l_refactor (documented at klu.c:296-301, never called by the reference)
 *   oracle_klu_solve    <- klu_l_solve / klu_l_tsolve (klu.c:651-657)
 *   oracle_klu_det      <- get_det (klu.c:764-813)
 *
 * Pins: doc known answers (reference doc/source/spsolvers.rst:333-345, 420-439), det = 114
 * (reference tests/test_sparse_solvers.py:298-313) and scipy SuperLU cross-checks in tests/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

typedef struct {
    i64 n;
    i64 *Lp, *Li, *Up, *Ui;     /* CSC; row indices are PIVOTAL positions; L unit diagonal first, U diagonal last */
    double *Lx, *Ux;
    i64 *Pnum;                  /* Pnum[k] = original row of pivotal row k                 */
    i64 *Q;                     /* Q[k] = original column of column k                      */
    double* Rs;                 /* row scale factors in pivotal order                      */
    i64 lcap, ucap;
    double flops;
    double *ws_rs, *ws_x;       /* refactor workspace, allocated once (klu_refactor keeps its workspace in the Numeric object too) */
    i64* ws_pinv;
    int usorted;                /* U columns already sorted by row (done by the first refactor) */
} oracle_klu;

void oracle_klu_free(oracle_klu* F) {
    if (!F) return;
    free(F->Lp); free(F->Li); free(F->Up); free(F->Ui); free(F->Lx); free(F->Ux); free(F->Pnum); free(F->Q); free(F->Rs);
    free(F->ws_rs); free(F->ws_x); free(F->ws_pinv);
    free(F);
}

static void grow(i64** idx, double** val, i64* cap, i64 need) {
    if (need <= *cap) return;
    i64 nc = *cap * 2 > need ? *cap * 2 : need;
    *idx = (i64*)realloc(*idx, sizeof(i64) * (size_t)nc);
    *val = (double*)realloc(*val, sizeof(double) * (size_t)nc);
    *cap = nc;
}

/* P0: row pre-ordering (P0[k] = original row at position k) or NULL; Q: column ordering or NULL.
 * Returns NULL when the matrix is singular (*singular_col = failing column). */
oracle_klu* oracle_klu_factor(i64 n, const i64* Ap, const i64* Ai, const double* Ax, const i64* P0, const i64* Q,
                              double tol, i64* singular_col) {
    oracle_klu* F = (oracle_klu*)calloc(1, sizeof *F);
    F->n = n;
    F->Lp = (i64*)calloc((size_t)n + 1, sizeof(i64));
    F->Up = (i64*)calloc((size_t)n + 1, sizeof(i64));
    F->Pnum = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    F->Q = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    F->Rs = (double*)malloc(sizeof(double) * ((size_t)n + 1));
    F->lcap = F->ucap = Ap[n] * 4 + 16;
    F->Li = (i64*)malloc(sizeof(i64) * (size_t)F->lcap); F->Lx = (double*)malloc(sizeof(double) * (size_t)F->lcap);
    F->Ui = (i64*)malloc(sizeof(i64) * (size_t)F->ucap); F->Ux = (double*)malloc(sizeof(double) * (size_t)F->ucap);
    if (singular_col) *singular_col = -1;
    double* rs = (double*)calloc((size_t)n + 1, sizeof(double));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) { double a = fabs(Ax[p]); if (a > rs[Ai[p]]) rs[Ai[p]] = a; }
    for (i64 i = 0; i < n; i++) if (!(rs[i] > 0.0)) rs[i] = 1.0;
    i64* pos0 = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));   /* original row -> pre-ordered position */
    for (i64 k = 0; k < n; k++) { i64 r = P0 ? P0[k] : k; pos0[r] = k; F->Q[k] = Q ? Q[k] : k; }
    i64* pivpos = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));  /* pre-ordered row -> pivotal position or -1 */
    i64* rowat = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    for (i64 k = 0; k < n; k++) pivpos[k] = -1;
    double* x = (double*)calloc((size_t)n + 1, sizeof(double));
    i64* mark = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    i64* visit = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    for (i64 k = 0; k < n; k++) { mark[k] = -1; visit[k] = -1; }
    i64* xi = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    i64* reach = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    i64* stk = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    i64* pst = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    int ok = 1;
    for (i64 k = 0; k < n && ok; k++) {
        const i64 col = F->Q[k];
        i64 nx = 0, nreach = 0;
        for (i64 p = Ap[col]; p < Ap[col + 1]; p++) {
            i64 r = pos0[Ai[p]];
            x[r] = Ax[p] / rs[Ai[p]];
            if (mark[r] != k) { mark[r] = k; xi[nx++] = r; }
        }
        i64 nx0 = nx;
        for (i64 q = 0; q < nx0; q++) {                 /* depth-first search through the columns of L */
            i64 r0 = xi[q];
            if (pivpos[r0] < 0 || visit[r0] == k) continue;
            i64 top = 0;
            stk[0] = r0; pst[0] = F->Lp[pivpos[r0]] + 1; visit[r0] = k;
            while (top >= 0) {
                i64 r = stk[top], jc = pivpos[r];
                int desc = 0;
                for (; pst[top] < F->Lp[jc + 1]; pst[top]++) {
                    i64 rr = F->Li[pst[top]];
                    if (mark[rr] != k) { mark[rr] = k; xi[nx++] = rr; x[rr] = 0.0; }
                    if (pivpos[rr] >= 0 && visit[rr] != k) {
                        visit[rr] = k;
                        pst[top]++;
                        top++;
                        stk[top] = rr; pst[top] = F->Lp[pivpos[rr]] + 1;
                        desc = 1;
                        break;
                    }
                }
                if (!desc) { reach[nreach++] = r; top--; }
            }
        }
        for (i64 t = nreach - 1; t >= 0; t--) {         /* x = L \ x in topological order */
            i64 r = reach[t], jc = pivpos[r];
            double xj = x[r];
            for (i64 p = F->Lp[jc] + 1; p < F->Lp[jc + 1]; p++) x[F->Li[p]] -= F->Lx[p] * xj;
            F->flops += 2.0 * (double)(F->Lp[jc + 1] - F->Lp[jc] - 1);
        }
        double amax = -1.0; i64 prow = -1;
        for (i64 q = 0; q < nx; q++) { i64 r = xi[q]; if (pivpos[r] < 0) { double a = fabs(x[r]); if (a > amax) { amax = a; prow = r; } } }
        if (pivpos[k] < 0 && mark[k] == k && fabs(x[k]) >= tol * amax && x[k] != 0.0) prow = k;   /* diagonal preference */
        if (prow < 0 || !(amax > 0.0) || x[prow] == 0.0) { if (singular_col) *singular_col = k; ok = 0; break; }
        double piv = x[prow];
        pivpos[prow] = k; rowat[k] = prow;
        grow(&F->Ui, &F->Ux, &F->ucap, F->Up[k] + nx + 1);
        grow(&F->Li, &F->Lx, &F->lcap, F->Lp[k] + nx + 1);
        i64 up = F->Up[k], lp = F->Lp[k];
        for (i64 q = 0; q < nx; q++) { i64 r = xi[q]; if (pivpos[r] >= 0 && r != prow) { F->Ui[up] = pivpos[r]; F->Ux[up++] = x[r]; } }
        F->Ui[up] = k; F->Ux[up++] = piv;
        F->Li[lp] = prow; F->Lx[lp++] = 1.0;
        for (i64 q = 0; q < nx; q++) { i64 r = xi[q]; if (pivpos[r] < 0) { F->Li[lp] = r; F->Lx[lp++] = x[r] / piv; } }
        F->Up[k + 1] = up; F->Lp[k + 1] = lp;
        for (i64 q = 0; q < nx; q++) x[xi[q]] = 0.0;
    }
    if (ok) {
        for (i64 p = 0; p < F->Lp[n]; p++) F->Li[p] = pivpos[F->Li[p]];      /* pre-ordered rows -> pivotal positions */
        for (i64 k = 0; k < n; k++) { F->Pnum[k] = P0 ? P0[rowat[k]] : rowat[k]; F->Rs[k] = rs[F->Pnum[k]]; }
    }
    free(rs); free(pos0); free(pivpos); free(rowat); free(x); free(mark); free(visit); free(xi); free(reach); free(stk); free(pst);
    if (!ok) { oracle_klu_free(F); return NULL; }
    return F;
}

/* klu_refactor: same pattern and pivot order, new values.  Returns 0, or 2 when a pivot is zero. */
int oracle_klu_refactor(oracle_klu* F, const i64* Ap, const i64* Ai, const double* Ax) {
    const i64 n = F->n;
    if (!F->ws_rs) {
        F->ws_rs = (double*)calloc((size_t)n + 1, sizeof(double));
        F->ws_pinv = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
        F->ws_x = (double*)calloc((size_t)n + 1, sizeof(double));
    }
    double* rs = F->ws_rs;
    i64* pinv = F->ws_pinv;
    double* x = F->ws_x;          /* all zero on entry and on exit */
    memset(rs, 0, sizeof(double) * ((size_t)n + 1));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) { double a = fabs(Ax[p]); if (a > rs[Ai[p]]) rs[Ai[p]] = a; }
    for (i64 i = 0; i < n; i++) if (!(rs[i] > 0.0)) rs[i] = 1.0;
    for (i64 k = 0; k < n; k++) { pinv[F->Pnum[k]] = k; F->Rs[k] = rs[F->Pnum[k]]; }
    int status = 0;
    for (i64 k = 0; k < n; k++) {
        const i64 col = F->Q[k];
        for (i64 p = Ap[col]; p < Ap[col + 1]; p++) x[pinv[Ai[p]]] = Ax[p] / rs[Ai[p]];
        /* U(:,k) is stored in the order the pivoting factorization produced; the dependency order is by row */
        /* process pivotal rows in ascending order: gather, sort by row index */
        i64 u0 = F->Up[k], u1 = F->Up[k + 1] - 1;
        if (!F->usorted) for (i64 a = u0 + 1; a < u1; a++) {             /* insertion sort of (Ui,Ux) by Ui, stable pattern */
            i64 ri = F->Ui[a]; double rx = F->Ux[a]; i64 b = a - 1;
            while (b >= u0 && F->Ui[b] > ri) { F->Ui[b + 1] = F->Ui[b]; F->Ux[b + 1] = F->Ux[b]; b--; }
            F->Ui[b + 1] = ri; F->Ux[b + 1] = rx;
        }
        for (i64 p = u0; p < u1; p++) {
            i64 j = F->Ui[p];
            double xj = x[j];
            F->Ux[p] = xj;
            x[j] = 0.0;
            for (i64 q = F->Lp[j] + 1; q < F->Lp[j + 1]; q++) x[F->Li[q]] -= F->Lx[q] * xj;
        }
        double piv = x[k];
        x[k] = 0.0;
        F->Ux[u1] = piv;
        if (!(fabs(piv) > 0.0)) status = 2;
        for (i64 q = F->Lp[k] + 1; q < F->Lp[k + 1]; q++) { F->Lx[q] = x[F->Li[q]] / piv; x[F->Li[q]] = 0.0; }
    }
    F->usorted = 1;
    return status;
}

/* trans = 0: A X = B; trans = 1: A^T X = B.  B is n x nrhs column-major, overwritten. */
void oracle_klu_solve(const oracle_klu* F, int trans, double* B, i64 nrhs, i64 ldB) {
    const i64 n = F->n;
    double* x = (double*)malloc(sizeof(double) * ((size_t)n + 1));
    for (i64 c = 0; c < nrhs; c++) {
        double* b = B + c * ldB;
        if (!trans) {
            for (i64 k = 0; k < n; k++) x[k] = b[F->Pnum[k]] / F->Rs[k];
            for (i64 k = 0; k < n; k++) { double xk = x[k]; for (i64 p = F->Lp[k] + 1; p < F->Lp[k + 1]; p++) x[F->Li[p]] -= F->Lx[p] * xk; }
            for (i64 k = n - 1; k >= 0; k--) {
                i64 u1 = F->Up[k + 1] - 1;
                double xk = x[k] / F->Ux[u1];
                x[k] = xk;
                for (i64 p = F->Up[k]; p < u1; p++) x[F->Ui[p]] -= F->Ux[p] * xk;
            }
            for (i64 k = 0; k < n; k++) b[F->Q[k]] = x[k];
        } else {
            for (i64 k = 0; k < n; k++) x[k] = b[F->Q[k]];
            for (i64 k = 0; k < n; k++) {
                i64 u1 = F->Up[k + 1] - 1;
                double acc = x[k];
                for (i64 p = F->Up[k]; p < u1; p++) acc -= F->Ux[p] * x[F->Ui[p]];
                x[k] = acc / F->Ux[u1];
            }
            for (i64 k = n - 1; k >= 0; k--) { double acc = x[k]; for (i64 p = F->Lp[k] + 1; p < F->Lp[k + 1]; p++) acc -= F->Lx[p] * x[F->Li[p]]; x[k] = acc; }
            for (i64 k = 0; k < n; k++) b[F->Pnum[k]] = x[k] / F->Rs[k];
        }
    }
    free(x);
}

/* determinant: prod(Udiag * Rs) * sign(P) * sign(Q), parity by cycle-sorting (klu.c:764-813) */
double oracle_klu_det(const oracle_klu* F) {
    const i64 n = F->n;
    double det = 1.0;
    for (i64 k = 0; k < n; k++) det *= F->Ux[F->Up[k + 1] - 1] * F->Rs[k];
    i64* w = (i64*)malloc(sizeof(i64) * ((size_t)n + 1));
    i64 npiv = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (i64 i = 0; i < n; i++) w[i] = pass ? F->Q[i] : F->Pnum[i];
        for (i64 i = 0; i < n; i++)
            while (w[i] != i) { i64 t = w[w[i]]; w[w[i]] = w[i]; w[i] = t; npiv++; }
    }
    free(w);
    return (npiv % 2) ? -det : det;
}
i64 oracle_klu_nnz(const oracle_klu* F, int which) { return which ? F->Up[F->n] : F->Lp[F->n]; }
double oracle_klu_flops(const oracle_klu* F) { return F->flops; }
/* the pivot sequence the factorization chose: out[k] = original row of pivotal row k (tests compare it with the product's) */
void oracle_klu_pnum(const oracle_klu* F, i64* out) { for (i64 k = 0; k < F->n; k++) out[k] = F->Pnum[k]; }
