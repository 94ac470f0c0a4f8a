/*
 * chol_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic that kvxopt's cholmod wrapper (reference src/C/cholmod.c) obtains
 * from SuiteSparse CHOLMOD 7.8.2 (pinned in the reference at .ci/config/versions.env:7; the library
 * itself is NOT in the reference tree and not installable here, so its published algorithm is
 * restated): supernodal LEFT-looking LL^T (Ng & Peyton 1993; Chen, Davis, Hager, Rajamanickam,
 * "Algorithm 887: CHOLMOD", ACM TOMS 2008) with BLAS-3 dsyrk/dgemm/dpotrf/dtrsm per supernode, and
 * the solve systems of cholmod_l_solve as numbered by the reference (src/C/cholmod.c:437-439).
 *
 *   oracle_chol_analyze   <- pack (cholmod.c:132-181) + cholmod_l_analyze_p (cholmod.c:269) with a
 *                            GIVEN permutation (the ordering itself is not arithmetic; tests pass the
 *                            product's permutation or identity so both sides factor the same P A P')
 *   oracle_chol_factorize <- cholmod_l_factorize (cholmod.c:362,677): left-looking supernodal
 *   oracle_chol_solve     <- cholmod_l_solve per column (cholmod.c:481-493), sys 0..8
 *   oracle_chol_diag      <- diag() (cholmod.c:900-945)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this.  Parity pins: doc known answers (reference doc/source/spsolvers.rst:555-563,580-585,759-772)
 * and LAPACK dense solves computed by the reference's own lapack module (tests/golden/*.npz).
 *
 * BLAS: links to scipy's bundled OpenBLAS (symbols scipy_d*_ , LP64) unless ORACLE_NAIVE_BLAS.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

#ifdef ORACLE_NAIVE_BLAS
static void dsyrk_ln(int n, int k, const double* A, int lda, double* C, int ldc) { /* C = A A' (lower) */
    for (int j = 0; j < n; j++)
        for (int i = j; i < n; i++) {
            double s = 0;
            for (int p = 0; p < k; p++) s += A[i + (i64)p * lda] * A[j + (i64)p * lda];
            C[i + (i64)j * ldc] = s;
        }
}
static void dgemm_nt(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            double s = 0;
            for (int p = 0; p < k; p++) s += A[i + (i64)p * lda] * B[j + (i64)p * ldb];
            C[i + (i64)j * ldc] = s;
        }
}
static int dpotrf_l(int n, double* A, int lda) {
    for (int j = 0; j < n; j++) {
        double d = A[j + (i64)j * lda];
        for (int p = 0; p < j; p++) d -= A[j + (i64)p * lda] * A[j + (i64)p * lda];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        A[j + (i64)j * lda] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i + (i64)j * lda];
            for (int p = 0; p < j; p++) s -= A[i + (i64)p * lda] * A[j + (i64)p * lda];
            A[i + (i64)j * lda] = s / d;
        }
    }
    return 0;
}
static void dtrsm_rltn(int m, int n, const double* L, int ldl, double* B, int ldb) { /* B := B L^-T */
    for (int j = 0; j < n; j++) {
        for (int p = 0; p < j; p++) {
            double l = L[j + (i64)p * ldl];
            for (int i = 0; i < m; i++) B[i + (i64)j * ldb] -= B[i + (i64)p * ldb] * l;
        }
        double d = L[j + (i64)j * ldl];
        for (int i = 0; i < m; i++) B[i + (i64)j * ldb] /= d;
    }
}
int oracle_blas_threads(int n) { (void)n; return 1; }
#else
extern void scipy_dsyrk_(const char*, const char*, const int*, const int*, const double*, const double*, const int*,
                         const double*, double*, const int*);
extern void scipy_dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                         const int*, const double*, const int*, const double*, double*, const int*);
extern void scipy_dpotrf_(const char*, const int*, double*, const int*, int*);
extern void scipy_dtrsm_(const char*, const char*, const char*, const char*, const int*, const int*, const double*,
                         const double*, const int*, double*, const int*);
extern void scipy_openblas_set_num_threads(int);
extern int scipy_openblas_get_num_threads(void);
static void dsyrk_ln(int n, int k, const double* A, int lda, double* C, int ldc) {
    double one = 1.0, zero = 0.0;
    scipy_dsyrk_("L", "N", &n, &k, &one, A, &lda, &zero, C, &ldc);
}
static void dgemm_nt(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    double one = 1.0, zero = 0.0;
    scipy_dgemm_("N", "T", &m, &n, &k, &one, A, &lda, B, &ldb, &zero, C, &ldc);
}
static int dpotrf_l(int n, double* A, int lda) {
    int info = 0;
    scipy_dpotrf_("L", &n, A, &lda, &info);
    return info;
}
static void dtrsm_rltn(int m, int n, const double* L, int ldl, double* B, int ldb) {
    double one = 1.0;
    scipy_dtrsm_("R", "L", "T", "N", &m, &n, &one, L, &ldl, B, &ldb);
}
int oracle_blas_threads(int n) {
    if (n > 0) scipy_openblas_set_num_threads(n);
    return scipy_openblas_get_num_threads();
}
#endif

typedef struct {
    i64 n, nsuper, nnzL, xsize;
    i64* perm;      /* perm[k] = original index of permuted row k                    */
    i64* iperm;
    i64* super;     /* nsuper+1: first column of each supernode                      */
    i64* pi;        /* nsuper+1: offsets into s[]                                    */
    i64* px;        /* nsuper+1: offsets into x[]                                    */
    i64* s;         /* row indices of every supernode (first nscol = its columns)    */
    double* x;      /* dense column-major panels, leading dimension = rows of supernode (cholmod.c:927-943) */
    i64* Cp;        /* lower triangle of P A P' as CCS pattern + map to input entries */
    i64* Ci;
    i64* Cmap;      /* index into the caller's value array                            */
    i64 minor;
    int is_numeric;
    double flops;
} oracle_chol;

static int cmp_i64(const void* a, const void* b) {
    i64 x = *(const i64*)a, y = *(const i64*)b;
    return (x > y) - (x < y);
}

void oracle_chol_free(oracle_chol* F) {
    if (!F) return;
    free(F->perm); free(F->iperm); free(F->super); free(F->pi); free(F->px); free(F->s); free(F->x);
    free(F->Cp); free(F->Ci); free(F->Cmap);
    free(F);
}

/* Symbolic analysis with a given permutation (NULL = identity).  Relaxed supernodes use the CHOLMOD
 * default thresholds nrelax = {4,16,48}, zrelax = {0.8,0.1,0.05}.  Returns NULL on invalid input. */
oracle_chol* oracle_chol_analyze(i64 n, const i64* colptr, const i64* rowind, char uplo, const i64* perm) {
    oracle_chol* F = (oracle_chol*)calloc(1, sizeof *F);
    F->n = n;
    F->minor = n;
    F->perm = (i64*)malloc(sizeof(i64) * (n + 1));
    F->iperm = (i64*)malloc(sizeof(i64) * (n + 1));
    for (i64 k = 0; k < n; k++) F->iperm[k] = -1;
    for (i64 k = 0; k < n; k++) {
        i64 v = perm ? perm[k] : k;
        if (v < 0 || v >= n || F->iperm[v] != -1) { oracle_chol_free(F); return NULL; }
        F->perm[k] = v;
        F->iperm[v] = k;
    }
    /* ---- pack: keep the uplo triangle (rows >= j for 'L', rows <= j for 'U'), permute, store lower */
    const int lower = (uplo == 'L' || uplo == 'l');
    i64* cnt = (i64*)calloc(n + 1, sizeof(i64));
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (lower ? i < j : i > j) continue;
            i64 r = F->iperm[i], c = F->iperm[j];
            if (r < c) c = r;
            cnt[c + 1]++;
        }
    for (i64 j = 0; j < n; j++) cnt[j + 1] += cnt[j];
    i64 cnz = cnt[n];
    F->Cp = (i64*)malloc(sizeof(i64) * (n + 1));
    memcpy(F->Cp, cnt, sizeof(i64) * (n + 1));
    i64* packed = (i64*)malloc(sizeof(i64) * 2 * (cnz + 1));   /* (row, source index) pairs, sorted per column */
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (lower ? i < j : i > j) continue;
            i64 r = F->iperm[i], c = F->iperm[j];
            if (r < c) { i64 t = r; r = c; c = t; }
            packed[2 * cnt[c]] = r;
            packed[2 * cnt[c] + 1] = k;
            cnt[c]++;
        }
    F->Ci = (i64*)malloc(sizeof(i64) * (cnz + 1));
    F->Cmap = (i64*)malloc(sizeof(i64) * (cnz + 1));
    for (i64 j = 0; j < n; j++) {
        qsort(packed + 2 * F->Cp[j], (size_t)(F->Cp[j + 1] - F->Cp[j]), 2 * sizeof(i64), cmp_i64);
        for (i64 p = F->Cp[j]; p < F->Cp[j + 1]; p++) { F->Ci[p] = packed[2 * p]; F->Cmap[p] = packed[2 * p + 1]; }
    }
    free(packed);
    free(cnt);
    /* ---- elimination tree + full column structures of L (sorted), children merged into parents */
    i64* parent = (i64*)malloc(sizeof(i64) * (n + 1));
    i64** st = (i64**)calloc(n + 1, sizeof(i64*));   /* structure of column j strictly below the diagonal */
    i64* len = (i64*)calloc(n + 1, sizeof(i64));
    i64* mark = (i64*)malloc(sizeof(i64) * (n + 1));
    i64* head = (i64*)malloc(sizeof(i64) * (n + 1));
    i64* next = (i64*)malloc(sizeof(i64) * (n + 1));
    for (i64 j = 0; j < n; j++) { mark[j] = -1; head[j] = -1; next[j] = -1; parent[j] = -1; }
    i64* tmp = (i64*)malloc(sizeof(i64) * (n + 1));
    i64 nnzL = 0;
    for (i64 j = 0; j < n; j++) {
        i64 m = 0;
        mark[j] = j;
        for (i64 p = F->Cp[j]; p < F->Cp[j + 1]; p++) {
            i64 i = F->Ci[p];
            if (i > j && mark[i] != j) { mark[i] = j; tmp[m++] = i; }
        }
        for (i64 c = head[j]; c != -1; c = next[c]) {
            for (i64 q = 0; q < len[c]; q++) {
                i64 i = st[c][q];
                if (i > j && mark[i] != j) { mark[i] = j; tmp[m++] = i; }
            }
        }
        qsort(tmp, (size_t)m, sizeof(i64), cmp_i64);
        st[j] = (i64*)malloc(sizeof(i64) * (m + 1));
        memcpy(st[j], tmp, sizeof(i64) * m);
        len[j] = m;
        nnzL += m + 1;
        if (m > 0) { parent[j] = tmp[0]; next[j] = head[tmp[0]]; head[tmp[0]] = j; }
        /* children's structures are no longer needed once merged -- keep only what supernode detection needs */
        for (i64 c = head[j]; c != -1; c = next[c]) { /* keep len[c]; free the list itself */ free(st[c]); st[c] = NULL; }
    }
    /* note: the permutation is used as given; parent[j] > j always holds, columns are processed in order */
    /* ---- supernodes: j+1 joins j when parent[j] == j+1 and len[j+1] == len[j]-1, then relaxed amalgamation */
    i64* sfirst = (i64*)malloc(sizeof(i64) * (n + 1));
    i64 ns = 0;
    for (i64 j = 0; j < n; j++)
        if (!(j > 0 && parent[j - 1] == j && len[j] == len[j - 1] - 1)) sfirst[ns++] = j;
    sfirst[ns] = n;
    /* relaxed: merge supernode t into the following supernode u when u starts at parent(last col of t) */
    {
        static const double nrelax[3] = {4, 16, 48}, zrelax[3] = {0.8, 0.1, 0.05};
        double* zeros = (double*)calloc(ns + 1, sizeof(double));
        double* hgt = (double*)malloc(sizeof(double) * (ns + 1));
        i64* first = (i64*)malloc(sizeof(i64) * (ns + 1));
        char* dead = (char*)calloc(ns + 1, 1);
        for (i64 t = 0; t < ns; t++) { first[t] = sfirst[t]; hgt[t] = (double)len[sfirst[t]] + 1.0; }
        for (i64 t = 0; t + 1 < ns; t++) {
            i64 last = sfirst[t + 1] - 1, u = t + 1;
            if (parent[last] != sfirst[u]) continue;
            double nt = (double)(last - first[t] + 1), nu = (double)(sfirst[u + 1] - sfirst[u]);
            double ntot = nt + nu, hnew = nt + hgt[u];
            double newz = nt * (hnew - hgt[t]);
            double z = zeros[t] + zeros[u] + newz;
            double lnz = ntot * hnew - ntot * (ntot - 1) / 2;
            int merge;
            if (ntot <= nrelax[0]) merge = 1;
            else if (ntot <= nrelax[1]) merge = z / lnz < zrelax[0];
            else if (ntot <= nrelax[2]) merge = z / lnz < zrelax[1];
            else merge = z / lnz < zrelax[2];
            if (newz == 0) merge = 1;
            if (!merge) continue;
            dead[t] = 1;
            first[u] = first[t];
            zeros[u] = z;
            hgt[u] = hnew;
        }
        i64 m = 0;
        for (i64 t = 0; t < ns; t++) if (!dead[t]) sfirst[m++] = first[t];
        sfirst[m] = n;
        ns = m;
        free(zeros); free(hgt); free(first); free(dead);
    }
    F->nsuper = ns;
    F->super = (i64*)malloc(sizeof(i64) * (ns + 1));
    memcpy(F->super, sfirst, sizeof(i64) * (ns + 1));
    free(sfirst);
    /* ---- row structure of each supernode: its columns, then the union of rows below (A and children) */
    i64* snode = (i64*)malloc(sizeof(i64) * (n + 1));
    for (i64 t = 0; t < ns; t++) for (i64 j = F->super[t]; j < F->super[t + 1]; j++) snode[j] = t;
    i64** srows = (i64**)calloc(ns + 1, sizeof(i64*));
    i64* slen = (i64*)calloc(ns + 1, sizeof(i64));
    i64* shead = (i64*)malloc(sizeof(i64) * (ns + 1));
    i64* snext = (i64*)malloc(sizeof(i64) * (ns + 1));
    for (i64 t = 0; t < ns; t++) { shead[t] = -1; snext[t] = -1; }
    for (i64 j = 0; j < n; j++) mark[j] = -1;
    F->pi = (i64*)malloc(sizeof(i64) * (ns + 1));
    F->px = (i64*)malloc(sizeof(i64) * (ns + 1));
    i64 ssize = 0, xsize = 0;
    double flops = 0;
    for (i64 t = 0; t < ns; t++) {
        i64 c0 = F->super[t], c1 = F->super[t + 1], m = 0;
        for (i64 j = c0; j < c1; j++)
            for (i64 p = F->Cp[j]; p < F->Cp[j + 1]; p++) {
                i64 i = F->Ci[p];
                if (i >= c1 && mark[i] != t) { mark[i] = t; tmp[m++] = i; }
            }
        for (i64 c = shead[t]; c != -1; c = snext[c]) {
            i64 cc = F->super[c + 1] - F->super[c];
            for (i64 q = cc; q < slen[c]; q++) {
                i64 i = srows[c][q];
                if (i >= c1 && mark[i] != t) { mark[i] = t; tmp[m++] = i; }
            }
        }
        qsort(tmp, (size_t)m, sizeof(i64), cmp_i64);
        i64 nscol = c1 - c0, nsrow = nscol + m;
        srows[t] = (i64*)malloc(sizeof(i64) * (nsrow + 1));
        for (i64 j = 0; j < nscol; j++) srows[t][j] = c0 + j;
        memcpy(srows[t] + nscol, tmp, sizeof(i64) * m);
        slen[t] = nsrow;
        if (m > 0) { i64 pt = snode[tmp[0]]; snext[t] = shead[pt]; shead[pt] = t; }
        F->pi[t] = ssize; F->px[t] = xsize;
        ssize += nsrow; xsize += nsrow * nscol;
        double c = (double)nscol, r = (double)m;
        flops += c * c * c / 3.0 + c * c * r + c * r * r;
    }
    F->pi[ns] = ssize; F->px[ns] = xsize;
    F->s = (i64*)malloc(sizeof(i64) * (ssize + 1));
    for (i64 t = 0; t < ns; t++) { memcpy(F->s + F->pi[t], srows[t], sizeof(i64) * slen[t]); free(srows[t]); }
    F->xsize = xsize;
    F->x = (double*)malloc(sizeof(double) * (xsize + 1));
    F->nnzL = nnzL;
    F->flops = flops;
    for (i64 j = 0; j < n; j++) free(st[j]);
    free(st); free(len); free(mark); free(head); free(next); free(parent); free(tmp);
    free(snode); free(srows); free(slen); free(shead); free(snext);
    return F;
}

i64 oracle_chol_nnzL(const oracle_chol* F) { return F->nnzL; }
i64 oracle_chol_nsuper(const oracle_chol* F) { return F->nsuper; }
double oracle_chol_flops(const oracle_chol* F) { return F->flops; }
i64 oracle_chol_minor(const oracle_chol* F) { return F->minor; }

/* Left-looking supernodal numeric factorization.  val indexes the caller's CCS arrays given to analyze.
 * Returns 0 on success, 1 when not positive definite (F->minor = failing column, permuted order). */
int oracle_chol_factorize(oracle_chol* F, const double* val) {
    const i64 n = F->n, ns = F->nsuper;
    F->is_numeric = 0;
    F->minor = n;
    if (n == 0) { F->is_numeric = 1; return 0; }
    i64* Map = (i64*)malloc(sizeof(i64) * (n + 1));
    i64* Head = (i64*)malloc(sizeof(i64) * (ns + 1));     /* descendants that still have to update supernode t */
    i64* Next = (i64*)malloc(sizeof(i64) * (ns + 1));
    i64* Lpos = (i64*)calloc(ns + 1, sizeof(i64));        /* first row of supernode d not yet consumed       */
    i64* snode = (i64*)malloc(sizeof(i64) * (n + 1));
    i64 maxrow = 0;
    for (i64 t = 0; t < ns; t++) {
        Head[t] = -1; Next[t] = -1;
        for (i64 j = F->super[t]; j < F->super[t + 1]; j++) snode[j] = t;
        i64 r = F->pi[t + 1] - F->pi[t];
        if (r > maxrow) maxrow = r;
    }
    double* Cbuf = (double*)malloc(sizeof(double) * ((size_t)maxrow * (size_t)maxrow + 1));
    int status = 0;
    for (i64 t = 0; t < ns && !status; t++) {
        const i64 c0 = F->super[t], c1 = F->super[t + 1], nscol = c1 - c0;
        const i64* rows = F->s + F->pi[t];
        const i64 nsrow = F->pi[t + 1] - F->pi[t];
        double* Lx = F->x + F->px[t];
        for (i64 q = 0; q < nsrow; q++) Map[rows[q]] = q;
        memset(Lx, 0, sizeof(double) * (size_t)(nsrow * nscol));
        for (i64 j = c0; j < c1; j++)
            for (i64 p = F->Cp[j]; p < F->Cp[j + 1]; p++) Lx[Map[F->Ci[p]] + (j - c0) * nsrow] += val[F->Cmap[p]];
        /* apply every pending descendant d: rows of d in [c0,c1) select the columns it updates */
        i64 d = Head[t];
        while (d != -1) {
            i64 dnext = Next[d];
            const i64* drows = F->s + F->pi[d];
            const i64 ndrow = F->pi[d + 1] - F->pi[d], ndcol = F->super[d + 1] - F->super[d];
            const double* Dx = F->x + F->px[d];
            i64 k1 = Lpos[d], k2 = k1;
            while (k2 < ndrow && drows[k2] < c1) k2++;
            const i64 m1 = k2 - k1, m2 = ndrow - k1;          /* C is m2 x m1: rows k1.. of d times rows k1..k2 */
            dsyrk_ln((int)m1, (int)ndcol, Dx + k1, (int)ndrow, Cbuf, (int)m2);
            if (m2 > m1) dgemm_nt((int)(m2 - m1), (int)m1, (int)ndcol, Dx + k2, (int)ndrow, Dx + k1, (int)ndrow, Cbuf + m1, (int)m2);
            for (i64 cj = 0; cj < m1; cj++) {
                double* dst = Lx + (drows[k1 + cj] - c0) * nsrow;
                for (i64 ci = cj; ci < m2; ci++) dst[Map[drows[k1 + ci]]] -= Cbuf[ci + cj * m2];
            }
            Lpos[d] = k2;
            if (k2 < ndrow) { i64 u = snode[drows[k2]]; Next[d] = Head[u]; Head[u] = d; }
            d = dnext;
        }
        int info = dpotrf_l((int)nscol, Lx, (int)nsrow);
        if (info != 0) { F->minor = c0 + info - 1; status = 1; break; }
        if (nsrow > nscol) dtrsm_rltn((int)(nsrow - nscol), (int)nscol, Lx, (int)nsrow, Lx + nscol, (int)nsrow);
        Lpos[t] = nscol;
        if (nsrow > nscol) { i64 u = snode[rows[nscol]]; Next[t] = Head[u]; Head[u] = t; }
    }
    free(Map); free(Head); free(Next); free(Lpos); free(snode); free(Cbuf);
    F->is_numeric = !status;
    return status;
}

static void fwd_col(const oracle_chol* F, double* y) {   /* L y = b in permuted coordinates */
    for (i64 t = 0; t < F->nsuper; t++) {
        const i64 c0 = F->super[t], nscol = F->super[t + 1] - c0, nsrow = F->pi[t + 1] - F->pi[t];
        const i64* rows = F->s + F->pi[t];
        const double* Lx = F->x + F->px[t];
        for (i64 j = 0; j < nscol; j++) {
            double v = y[c0 + j] / Lx[j + j * nsrow];
            y[c0 + j] = v;
            for (i64 i = j + 1; i < nsrow; i++) y[rows[i]] -= Lx[i + j * nsrow] * v;
        }
    }
}
static void bwd_col(const oracle_chol* F, double* y) {   /* L' x = y */
    for (i64 t = F->nsuper - 1; t >= 0; t--) {
        const i64 c0 = F->super[t], nscol = F->super[t + 1] - c0, nsrow = F->pi[t + 1] - F->pi[t];
        const i64* rows = F->s + F->pi[t];
        const double* Lx = F->x + F->px[t];
        for (i64 j = nscol - 1; j >= 0; j--) {
            double v = y[c0 + j];
            for (i64 i = j + 1; i < nsrow; i++) v -= Lx[i + j * nsrow] * y[rows[i]];
            y[c0 + j] = v / Lx[j + j * nsrow];
        }
    }
}

/* sys numbering of the reference (cholmod.c:437-439): 0 A x=b, 1 LDL'x=b, 2 LD x=b, 3 DL'x=b, 4 L x=b,
 * 5 L'x=b, 6 D x=b, 7 x=P b, 8 x=P'b; the factor is LL' so D = I.  One column at a time like :481-493. */
int oracle_chol_solve(const oracle_chol* F, int sys, double* B, i64 nrhs, i64 ldB) {
    const i64 n = F->n;
    if (sys < 0 || sys > 8) return -4;
    if (n == 0 || nrhs == 0) return 0;
    if (!F->is_numeric) return -4;
    double* y = (double*)malloc(sizeof(double) * (size_t)n);
    for (i64 c = 0; c < nrhs; c++) {
        double* b = B + c * ldB;
        switch (sys) {
            case 0:
                for (i64 k = 0; k < n; k++) y[k] = b[F->perm[k]];
                fwd_col(F, y); bwd_col(F, y);
                for (i64 k = 0; k < n; k++) b[F->perm[k]] = y[k];
                break;
            case 1: fwd_col(F, b); bwd_col(F, b); break;
            case 2: case 4: fwd_col(F, b); break;
            case 3: case 5: bwd_col(F, b); break;
            case 6: break;
            case 7:
                for (i64 k = 0; k < n; k++) y[k] = b[F->perm[k]];
                memcpy(b, y, sizeof(double) * (size_t)n);
                break;
            case 8:
                for (i64 k = 0; k < n; k++) y[F->perm[k]] = b[k];
                memcpy(b, y, sizeof(double) * (size_t)n);
                break;
        }
    }
    free(y);
    return 0;
}

/* diagonal of L: strided copy with stride nsrow+1 inside each supernode (cholmod.c:927-943) */
void oracle_chol_diag(const oracle_chol* F, double* d) {
    for (i64 t = 0; t < F->nsuper; t++) {
        const i64 c0 = F->super[t], nscol = F->super[t + 1] - c0, nsrow = F->pi[t + 1] - F->pi[t];
        for (i64 j = 0; j < nscol; j++) d[c0 + j] = F->x[F->px[t] + j * (nsrow + 1)];
    }
}
void oracle_chol_get_perm(const oracle_chol* F, i64* p) { memcpy(p, F->perm, sizeof(i64) * (size_t)F->n); }

/* L as dense column-major n x n (small n only; for entrywise comparison in tests) */
void oracle_chol_dense_L(const oracle_chol* F, double* Ld) {
    const i64 n = F->n;
    memset(Ld, 0, sizeof(double) * (size_t)(n * n));
    for (i64 t = 0; t < F->nsuper; t++) {
        const i64 c0 = F->super[t], nscol = F->super[t + 1] - c0, nsrow = F->pi[t + 1] - F->pi[t];
        const i64* rows = F->s + F->pi[t];
        for (i64 j = 0; j < nscol; j++)
            for (i64 i = j; i < nsrow; i++) Ld[rows[i] + (c0 + j) * n] = F->x[F->px[t] + i + j * nsrow];
    }
}
