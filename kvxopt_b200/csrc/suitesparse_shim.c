/*
 * suitesparse_shim.c -- the CHOLMOD / KLU entry points that kvxopt's wrappers src/C/cholmod.c and src/C/klu.c call,
 * implemented on libb200sparse.so (include/b200sparse.h).  With include/suitesparse_shim/{cholmod,klu}.h this lets the
 * reference's extension-module sources compile unmodified and run their numeric work on the B200
 * (tools/build_kvxopt_ext.sh).  Plain C, host only; every function states which C-ABI call serves it.
 *
 * Ownership follows SuiteSparse: objects returned here are released by the matching *_free_* call, which the wrappers'
 * capsule destructors make (cholmod.c:210-214, klu.c:51-72).
 */
#include "../../include/suitesparse_shim/cholmod.h"
#include "../../include/suitesparse_shim/klu.h"
#include "../../include/b200sparse.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int chol_status_of(b200s_status st) {
    switch (st) {
        case B200S_OK: return CHOLMOD_OK;
        case B200S_NOT_POSDEF: return CHOLMOD_NOT_POSDEF;
        case B200S_OUT_OF_MEMORY: return CHOLMOD_OUT_OF_MEMORY;
        case B200S_TOO_LARGE: return CHOLMOD_TOO_LARGE;
        case B200S_NO_DEVICE:
        case B200S_CUDA_ERROR:
            fprintf(stderr, "kvxopt.cholmod (B200): %s: %s\n", b200s_strerror(st), b200s_last_error());
            return CHOLMOD_GPU_PROBLEM;
        default: return CHOLMOD_INVALID;
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * CHOLMOD
 * ------------------------------------------------------------------------------------------------------------- */
int cholmod_l_start(cholmod_common* cm) { return cholmod_l_defaults(cm); }
int cholmod_l_finish(cholmod_common* cm) { (void)cm; return 1; }
int cholmod_l_defaults(cholmod_common* cm) {
    if (!cm) return 0;
    cm->status = CHOLMOD_OK;
    cm->print = 3; cm->supernodal = 1; cm->nmethods = 0; cm->postorder = 1; cm->dbound = 0.0;   /* CHOLMOD's defaults */
    return 1;
}

cholmod_sparse* cholmod_l_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted, int packed, int stype, int xtype,
                                          cholmod_common* cm) {
    cholmod_sparse* A = (cholmod_sparse*)calloc(1, sizeof *A);
    const size_t es = xtype == CHOLMOD_COMPLEX ? 2 * sizeof(double) : sizeof(double);
    const size_t nz1 = nzmax > 0 ? nzmax : 1;
    if (A) {
        A->nrow = nrow; A->ncol = ncol; A->nzmax = nz1; A->sorted = sorted; A->packed = packed; A->stype = stype; A->xtype = xtype;
        A->p = calloc(ncol + 1, sizeof(int64_t));
        A->i = malloc(nz1 * sizeof(int64_t));
        A->x = xtype == CHOLMOD_PATTERN ? NULL : malloc(nz1 * es);
        A->nz = packed ? NULL : calloc(ncol > 0 ? ncol : 1, sizeof(int64_t));
    }
    if (!A || !A->p || !A->i || (xtype != CHOLMOD_PATTERN && !A->x) || (!packed && !A->nz)) {
        if (A) { free(A->p); free(A->i); free(A->x); free(A->nz); free(A); }
        if (cm) cm->status = CHOLMOD_OUT_OF_MEMORY;
        return NULL;
    }
    return A;
}
int cholmod_l_free_sparse(cholmod_sparse** A, cholmod_common* cm) {
    (void)cm;
    if (A && *A) { free((*A)->p); free((*A)->i); free((*A)->x); free((*A)->nz); free((*A)->z); free(*A); *A = NULL; }
    return 1;
}
cholmod_dense* cholmod_l_allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype, cholmod_common* cm) {
    cholmod_dense* X = (cholmod_dense*)calloc(1, sizeof *X);
    const size_t es = xtype == CHOLMOD_COMPLEX ? 2 * sizeof(double) : sizeof(double);
    if (X) {
        X->nrow = nrow; X->ncol = ncol; X->d = d; X->nzmax = d * ncol > 0 ? d * ncol : 1; X->xtype = xtype;
        X->x = calloc(X->nzmax, es);
    }
    if (!X || !X->x) { free(X); if (cm) cm->status = CHOLMOD_OUT_OF_MEMORY; return NULL; }
    return X;
}
int cholmod_l_free_dense(cholmod_dense** X, cholmod_common* cm) {
    (void)cm;
    if (X && *X) { free((*X)->x); free(*X); *X = NULL; }
    return 1;
}
int cholmod_l_check_perm(void* Perm, size_t len, size_t n, cholmod_common* cm) {
    (void)cm;
    const int64_t* p = (const int64_t*)Perm;
    if (!p && len > 0) return 0;
    unsigned char* seen = (unsigned char*)calloc(n > 0 ? n : 1, 1);
    int ok = seen != NULL;
    for (size_t k = 0; ok && k < len; k++) {
        if (p[k] < 0 || (size_t)p[k] >= n || seen[p[k]]) ok = 0; else seen[p[k]] = 1;
    }
    free(seen);
    return ok;
}

/* cholmod.c:274,663,811 -> b200s_chol_analyze.  A is the packed `uplo` triangle (pack, cholmod.c:132-181: stype -1 lower, +1 upper) */
cholmod_factor* cholmod_l_analyze_p(cholmod_sparse* A, void* UserPerm, void* fset, size_t fsize, cholmod_common* cm) {
    (void)fset; (void)fsize;
    if (!A || !cm || A->nrow != A->ncol || A->stype == 0) { if (cm) cm->status = CHOLMOD_INVALID; return NULL; }
    b200s_chol_opts o;
    b200s_chol_default_opts(&o);
    o.supernodal = cm->supernodal; o.nmethods = cm->nmethods; o.postorder = cm->postorder; o.dbound = cm->dbound;
    b200s_chol* F = NULL;
    /* 'z' matrices (pack() builds the cholmod_sparse with CHOLMOD_COMPLEX, cholmod.c:144,153): the factor object of the real
     * embedding, b200s_chol_analyze_z */
    b200s_status st = A->xtype == CHOLMOD_COMPLEX
        ? b200s_chol_analyze_z((b200s_int)A->nrow, (const b200s_int*)A->p, (const b200s_int*)A->i, A->stype < 0 ? 'L' : 'U',
                               (const b200s_int*)UserPerm, &o, &F)
        : b200s_chol_analyze((b200s_int)A->nrow, (const b200s_int*)A->p, (const b200s_int*)A->i, A->stype < 0 ? 'L' : 'U',
                             (const b200s_int*)UserPerm, &o, &F);
    if (st != B200S_OK) { cm->status = chol_status_of(st); return NULL; }
    cholmod_factor* L = (cholmod_factor*)calloc(1, sizeof *L);
    if (!L) { b200s_chol_free(F); cm->status = CHOLMOD_OUT_OF_MEMORY; return NULL; }
    L->n = A->nrow; L->minor = A->nrow; L->xtype = CHOLMOD_PATTERN; L->b200s = F; L->stype = A->stype;
    L->is_ll = cm->supernodal != 0;      /* supernodal = 0 asks for LDL' (cholmod.c:60-64); 1 and 2 are LL' in this engine */
    L->is_super = L->is_ll;
    L->zfactor = A->xtype == CHOLMOD_COMPLEX;
    cm->status = CHOLMOD_OK;
    return L;
}

/* cholmod.c:362,677,824 -> b200s_chol_factorize: A's own (colptr, rowind, values) travel, the pattern is checked there */
int cholmod_l_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common* cm) {
    if (!A || !L || !cm) { if (cm) cm->status = CHOLMOD_INVALID; return 0; }
    const int z = A->xtype == CHOLMOD_COMPLEX;
    if ((A->xtype != CHOLMOD_REAL && !z) || z != L->zfactor) { cm->status = CHOLMOD_INVALID; return 0; }
    b200s_int minor = (b200s_int)L->n;
    b200s_status st = z
        ? b200s_chol_factorize_z((b200s_chol*)L->b200s, (const b200s_int*)A->p, (const b200s_int*)A->i, (const double*)A->x, &minor)
        : b200s_chol_factorize((b200s_chol*)L->b200s, (const b200s_int*)A->p, (const b200s_int*)A->i, (const double*)A->x, &minor);
    cm->status = chol_status_of(st);
    if (st == B200S_OK || st == B200S_NOT_POSDEF) {
        L->xtype = z ? CHOLMOD_COMPLEX : CHOLMOD_REAL;
        L->minor = st == B200S_OK ? L->n : (size_t)minor;
    } else {
        L->xtype = CHOLMOD_PATTERN;
        return 0;
    }
    if (st == B200S_OK && L->is_ll && L->n > 0) {
        /* the members diag() reads (cholmod.c:927-943): n supernodes of one column and one row whose "panel" is the
         * diagonal entry -- b200s_chol_diag gathers it on the device, n doubles come back */
        const size_t n = L->n;
        if (!L->super) {
            int64_t* idx = (int64_t*)malloc((n + 1) * sizeof(int64_t));
            if (idx) for (size_t k = 0; k <= n; k++) idx[k] = (int64_t)k;
            L->super = idx; L->pi = idx; L->px = idx;
            L->x = malloc(n * (z ? 2 : 1) * sizeof(double));
        }
        if (!L->super || !L->x) { cm->status = CHOLMOD_OUT_OF_MEMORY; return 0; }
        if ((z ? b200s_chol_diag_z((b200s_chol*)L->b200s, (double*)L->x) : b200s_chol_diag((b200s_chol*)L->b200s, (double*)L->x)) != B200S_OK) {
            cm->status = CHOLMOD_GPU_PROBLEM; return 0;
        }
        L->nsuper = n;
    }
    return 1;
}

/* cholmod.c:483,735 -> b200s_chol_solve (one column per call, as the wrapper's loop asks) */
cholmod_dense* cholmod_l_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common* cm) {
    if (!L || !B || !cm) { if (cm) cm->status = CHOLMOD_INVALID; return NULL; }
    if (B->xtype != L->xtype || (L->xtype != CHOLMOD_REAL && L->xtype != CHOLMOD_COMPLEX)) { cm->status = CHOLMOD_INVALID; return NULL; }
    /* a complex vector is its own real embedding: (re, im) pairs = 2n real unknowns per column */
    const size_t w = L->xtype == CHOLMOD_COMPLEX ? 2 : 1;
    cholmod_dense* X = cholmod_l_allocate_dense(L->n, B->ncol, L->n, L->xtype, cm);
    if (!X) return NULL;
    for (size_t c = 0; c < B->ncol; c++) memcpy((double*)X->x + c * w * L->n, (const double*)B->x + c * w * B->d, w * L->n * sizeof(double));
    b200s_status st = b200s_chol_solve((b200s_chol*)L->b200s, sys, (double*)X->x, (b200s_int)B->ncol, (b200s_int)(w * L->n));
    cm->status = chol_status_of(st);
    return X;
}

/* cholmod.c:567,858 -> b200s_chol_spsolve */
cholmod_sparse* cholmod_l_spsolve(int sys, cholmod_factor* L, cholmod_sparse* B, cholmod_common* cm) {
    if (!L || !B || !cm) { if (cm) cm->status = CHOLMOD_INVALID; return NULL; }
    if (B->xtype != L->xtype || (L->xtype != CHOLMOD_REAL && L->xtype != CHOLMOD_COMPLEX)) { cm->status = CHOLMOD_INVALID; return NULL; }
    b200s_int *xp = NULL, *xi = NULL;
    double* xx = NULL;
    b200s_status st = L->xtype == CHOLMOD_COMPLEX
        ? b200s_chol_spsolve_z((b200s_chol*)L->b200s, sys, (b200s_int)B->nrow, (b200s_int)B->ncol, (const b200s_int*)B->p,
                               (const b200s_int*)B->i, (const double*)B->x, &xp, &xi, &xx)
        : b200s_chol_spsolve((b200s_chol*)L->b200s, sys, (b200s_int)B->nrow, (b200s_int)B->ncol, (const b200s_int*)B->p,
                             (const b200s_int*)B->i, (const double*)B->x, &xp, &xi, &xx);
    cm->status = chol_status_of(st);
    if (st != B200S_OK) return NULL;
    cholmod_sparse* X = (cholmod_sparse*)calloc(1, sizeof *X);
    if (!X) { b200s_free(xp); b200s_free(xi); b200s_free(xx); cm->status = CHOLMOD_OUT_OF_MEMORY; return NULL; }
    X->nrow = B->nrow; X->ncol = B->ncol; X->nzmax = (size_t)xp[B->ncol]; X->p = xp; X->i = xi; X->x = xx;
    X->sorted = 1; X->packed = 1; X->xtype = L->xtype;
    return X;
}

/* cholmod.c:969 -> b200s_chol_get_L (the device factor stays as it is; CHOLMOD converts L in place) */
cholmod_sparse* cholmod_l_factor_to_sparse(cholmod_factor* L, cholmod_common* cm) {
    if (!L || !cm) { if (cm) cm->status = CHOLMOD_INVALID; return NULL; }
    b200s_int *lp = NULL, *li = NULL;
    double* lx = NULL;
    b200s_status st = L->xtype == CHOLMOD_COMPLEX ? b200s_chol_get_L_z((b200s_chol*)L->b200s, &lp, &li, &lx)
                                                  : b200s_chol_get_L((b200s_chol*)L->b200s, &lp, &li, &lx);
    cm->status = chol_status_of(st);
    if (st != B200S_OK) return NULL;
    cholmod_sparse* S = (cholmod_sparse*)calloc(1, sizeof *S);
    if (!S) { b200s_free(lp); b200s_free(li); b200s_free(lx); cm->status = CHOLMOD_OUT_OF_MEMORY; return NULL; }
    S->nrow = L->n; S->ncol = L->n; S->nzmax = L->n ? (size_t)lp[L->n] : 0; S->p = lp; S->i = li; S->x = lx;
    S->sorted = 1; S->packed = 1; S->xtype = L->xtype;
    return S;
}

int cholmod_l_free_factor(cholmod_factor** L, cholmod_common* cm) {
    (void)cm;
    if (L && *L) {
        b200s_chol_free((b200s_chol*)(*L)->b200s);
        free((*L)->super);       /* pi and px alias it */
        free((*L)->x);
        free(*L);
        *L = NULL;
    }
    return 1;
}

/* ---------------------------------------------------------------------------------------------------------------
 * KLU
 * ------------------------------------------------------------------------------------------------------------- */
static int klu_status_of(b200s_status st) {
    switch (st) {
        case B200S_OK: return KLU_OK;
        case B200S_SINGULAR: return KLU_SINGULAR;
        case B200S_OUT_OF_MEMORY: return KLU_OUT_OF_MEMORY;
        case B200S_TOO_LARGE: return KLU_TOO_LARGE;
        case B200S_NO_DEVICE:
        case B200S_CUDA_ERROR:
            fprintf(stderr, "kvxopt.klu (B200): %s: %s\n", b200s_strerror(st), b200s_last_error());
            return KLU_INVALID;
        default: return KLU_INVALID;
    }
}

int klu_l_defaults(klu_l_common* cm) {
    if (!cm) return 0;
    cm->tol = 0.001; cm->memgrow = 1.2; cm->initmem_amd = 1.2; cm->initmem = 10; cm->maxwork = 0;
    cm->btf = 1; cm->ordering = 0; cm->scale = 2; cm->halt_if_singular = 1; cm->status = KLU_OK;
    return 1;
}

/* klu.c:142,264 -> b200s_klu_analyze */
klu_l_symbolic* klu_l_analyze(int64_t n, int64_t* Ap, int64_t* Ai, klu_l_common* cm) {
    if (!cm) return NULL;
    b200s_klu_sym* S = NULL;
    b200s_status st = b200s_klu_analyze(n, Ap, Ai, &S);
    cm->status = klu_status_of(st);
    if (st != B200S_OK) return NULL;
    klu_l_symbolic* Y = (klu_l_symbolic*)calloc(1, sizeof *Y);
    if (!Y) { b200s_klu_free_symbolic(S); cm->status = KLU_OUT_OF_MEMORY; return NULL; }
    Y->n = n; Y->nz = n > 0 ? Ap[n] : 0; Y->b200s = S;
    Y->nblocks = 0;          /* known after the first factorization (the block list travels with the numeric object) */
    return Y;
}

/* klu.c:160,337 -> b200s_klu_factor; the members get_numeric / get_det read directly are filled from b200s_klu_extract */
klu_l_numeric* klu_l_factor(int64_t* Ap, int64_t* Ai, double* Ax, klu_l_symbolic* Y, klu_l_common* cm) {
    if (!cm) return NULL;
    if (!Y) { cm->status = KLU_INVALID; return NULL; }
    b200s_klu_num* N = NULL;
    b200s_status st = b200s_klu_factor((b200s_klu_sym*)Y->b200s, Ap, Ai, Ax, &N);
    cm->status = klu_status_of(st);
    if (st != B200S_OK) return NULL;
    b200s_klu_info_t inf;
    b200s_klu_info(N, &inf);
    const int64_t n = inf.n;
    klu_l_numeric* F = (klu_l_numeric*)calloc(1, sizeof *F);
    int64_t *Up = NULL, *Ui = NULL, *Q = NULL, *R = NULL;
    double* Ux = NULL;
    if (F) {
        F->n = n; F->nblocks = inf.nblocks; F->lnz = inf.nnz_L; F->unz = inf.nnz_U; F->nzoff = inf.nnz_F; F->b200s = N;
        F->Pnum = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        F->Rs = (double*)malloc((size_t)(n + 1) * sizeof(double));
        F->Udiag = malloc((size_t)(n + 1) * sizeof(double));
        Up = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        Ui = (int64_t*)malloc((size_t)(inf.nnz_U + 1) * sizeof(int64_t));
        Ux = (double*)malloc((size_t)(inf.nnz_U + 1) * sizeof(double));
        Q = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        R = (int64_t*)malloc((size_t)(inf.nblocks + 2) * sizeof(int64_t));
    }
    if (!F || !F->Pnum || !F->Rs || !F->Udiag || !Up || !Ui || !Ux || !Q || !R) {
        if (F) { free(F->Pnum); free(F->Rs); free(F->Udiag); free(F); }
        free(Up); free(Ui); free(Ux); free(Q); free(R);
        b200s_klu_free_numeric(N);
        cm->status = KLU_OUT_OF_MEMORY;
        return NULL;
    }
    st = b200s_klu_extract(N, NULL, NULL, NULL, Up, Ui, Ux, NULL, NULL, NULL, F->Pnum, Q, F->Rs, R);
    if (st == B200S_OK) {
        /* U(k,k) is the last entry of column k (rows ascending) */
        for (int64_t k = 0; k < n; k++) ((double*)F->Udiag)[k] = Ux[Up[k + 1] - 1];
        /* klu.c:766 reads Symbolic->Q, :475 Symbolic->nblocks: the ordering is fixed at analyze time and identical for
         * every numeric object of this symbolic object, so it is recorded there once */
        if (!Y->Q) { Y->Q = Q; Q = NULL; Y->R = R; R = NULL; Y->nblocks = inf.nblocks; Y->maxblock = inf.max_block; }
    }
    free(Up); free(Ui); free(Ux); free(Q); free(R);
    if (st != B200S_OK) {
        klu_l_numeric* f = F;
        klu_l_free_numeric(&f, cm);
        cm->status = klu_status_of(st);
        return NULL;
    }
    return F;
}

static int klu_solve_common(klu_l_numeric* F, int trans, int64_t ldim, int64_t nrhs, double* B, klu_l_common* cm) {
    if (!cm) return 0;
    if (!F) { cm->status = KLU_INVALID; return 0; }
    b200s_status st = b200s_klu_solve((b200s_klu_num*)F->b200s, trans, B, nrhs, ldim);
    cm->status = klu_status_of(st);
    return st == B200S_OK;
}
/* klu.c:189,651 -> b200s_klu_solve(trans = 0);  klu.c:191,655 -> b200s_klu_solve(trans = 1) */
int klu_l_solve(klu_l_symbolic* Y, klu_l_numeric* F, int64_t ldim, int64_t nrhs, double* B, klu_l_common* cm) {
    (void)Y;
    return klu_solve_common(F, 0, ldim, nrhs, B, cm);
}
int klu_l_tsolve(klu_l_symbolic* Y, klu_l_numeric* F, int64_t ldim, int64_t nrhs, double* B, klu_l_common* cm) {
    (void)Y;
    return klu_solve_common(F, 1, ldim, nrhs, B, cm);
}

/* klu.c:461 -> b200s_klu_extract (Rs are the scale factors themselves; the wrapper inverts them, klu.c:516-522) */
int klu_l_extract(klu_l_numeric* F, klu_l_symbolic* Y, int64_t* Lp, int64_t* Li, double* Lx, int64_t* Up, int64_t* Ui, double* Ux,
                  int64_t* Fp, int64_t* Fi, double* Fx, int64_t* P, int64_t* Q, double* Rs, int64_t* R, klu_l_common* cm) {
    (void)Y;
    if (!cm) return 0;
    if (!F) { cm->status = KLU_INVALID; return 0; }
    b200s_status st = b200s_klu_extract((b200s_klu_num*)F->b200s, Lp, Li, Lx, Up, Ui, Ux, Fp, Fi, Fx, P, Q, Rs, R);
    cm->status = klu_status_of(st);
    return st == B200S_OK;
}

int klu_l_free_symbolic(klu_l_symbolic** Y, klu_l_common* cm) {
    (void)cm;
    if (Y && *Y) { b200s_klu_free_symbolic((b200s_klu_sym*)(*Y)->b200s); free((*Y)->Q); free((*Y)->R); free(*Y); *Y = NULL; }
    return 1;
}
int klu_l_free_numeric(klu_l_numeric** F, klu_l_common* cm) {
    (void)cm;
    if (F && *F) { b200s_klu_free_numeric((b200s_klu_num*)(*F)->b200s); free((*F)->Pnum); free((*F)->Rs); free((*F)->Udiag); free(*F); *F = NULL; }
    return 1;
}

/* ---- complex KLU (klu.c:161-162,348-355,468-479,661-668,754-813): b200s_klu_factor_z / _solve_z / _extract_z ------------- */
/* klu.c:352 -> b200s_klu_factor_z; Udiag holds (re, im) pairs (klu.c:760 reads it as double complex) */
klu_l_numeric* klu_zl_factor(int64_t* Ap, int64_t* Ai, double* Ax, klu_l_symbolic* Y, klu_l_common* cm) {
    if (!cm) return NULL;
    if (!Y) { cm->status = KLU_INVALID; return NULL; }
    b200s_klu_num* N = NULL;
    b200s_status st = b200s_klu_factor_z((b200s_klu_sym*)Y->b200s, Ap, Ai, Ax, &N);
    cm->status = klu_status_of(st);
    if (st != B200S_OK) return NULL;
    b200s_klu_info_t inf;
    b200s_klu_info(N, &inf);
    const int64_t n = inf.n;
    klu_l_numeric* F = (klu_l_numeric*)calloc(1, sizeof *F);
    int64_t *Up = NULL, *Ui = NULL, *Q = NULL, *R = NULL;
    double* Ux = NULL;
    if (F) {
        F->n = n; F->nblocks = inf.nblocks; F->lnz = inf.nnz_L; F->unz = inf.nnz_U; F->nzoff = inf.nnz_F; F->b200s = N;
        F->Pnum = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        F->Rs = (double*)malloc((size_t)(n + 1) * sizeof(double));
        F->Udiag = malloc((size_t)(n + 1) * 2 * sizeof(double));
        Up = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        Ui = (int64_t*)malloc((size_t)(inf.nnz_U + 1) * sizeof(int64_t));
        Ux = (double*)malloc((size_t)(inf.nnz_U + 1) * 2 * sizeof(double));
        Q = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
        R = (int64_t*)malloc((size_t)(inf.nblocks + 2) * sizeof(int64_t));
    }
    if (!F || !F->Pnum || !F->Rs || !F->Udiag || !Up || !Ui || !Ux || !Q || !R) {
        if (F) { free(F->Pnum); free(F->Rs); free(F->Udiag); free(F); }
        free(Up); free(Ui); free(Ux); free(Q); free(R);
        b200s_klu_free_numeric(N);
        cm->status = KLU_OUT_OF_MEMORY;
        return NULL;
    }
    st = b200s_klu_extract_z(N, NULL, NULL, NULL, Up, Ui, Ux, NULL, NULL, NULL, F->Pnum, Q, F->Rs, R);
    if (st == B200S_OK) {
        for (int64_t k = 0; k < n; k++) {
            ((double*)F->Udiag)[2 * k] = Ux[2 * (Up[k + 1] - 1)];
            ((double*)F->Udiag)[2 * k + 1] = Ux[2 * (Up[k + 1] - 1) + 1];
        }
        if (!Y->Q) { Y->Q = Q; Q = NULL; Y->R = R; R = NULL; Y->nblocks = inf.nblocks; Y->maxblock = inf.max_block; }
    }
    free(Up); free(Ui); free(Ux); free(Q); free(R);
    if (st != B200S_OK) {
        klu_l_numeric* f = F;
        klu_l_free_numeric(&f, cm);
        cm->status = klu_status_of(st);
        return NULL;
    }
    return F;
}
/* klu.c:661 -> b200s_klu_solve_z(trans = 0); klu.c:665 -> trans = 1 (A^T) or 2 (A^H, conj_solve) */
int klu_zl_solve(klu_l_symbolic* Y, klu_l_numeric* F, int64_t ldim, int64_t nrhs, double* B, klu_l_common* cm) {
    (void)Y;
    if (!cm) return 0;
    if (!F) { cm->status = KLU_INVALID; return 0; }
    b200s_status st = b200s_klu_solve_z((b200s_klu_num*)F->b200s, 0, B, nrhs, ldim);
    cm->status = klu_status_of(st);
    return st == B200S_OK;
}
int klu_zl_tsolve(klu_l_symbolic* Y, klu_l_numeric* F, int64_t ldim, int64_t nrhs, double* B, int conj_solve, klu_l_common* cm) {
    (void)Y;
    if (!cm) return 0;
    if (!F) { cm->status = KLU_INVALID; return 0; }
    b200s_status st = b200s_klu_solve_z((b200s_klu_num*)F->b200s, conj_solve ? 2 : 1, B, nrhs, ldim);
    cm->status = klu_status_of(st);
    return st == B200S_OK;
}
int klu_zl_free_numeric(klu_l_numeric** F, klu_l_common* cm) { return klu_l_free_numeric(F, cm); }
/* klu.c:480: SuiteSparse's split real / imaginary arrays, filled from the interleaved values of b200s_klu_extract_z */
int klu_zl_extract(klu_l_numeric* F, klu_l_symbolic* Y, int64_t* Lp, int64_t* Li, double* Lx, double* Lz, int64_t* Up, int64_t* Ui,
                   double* Ux, double* Uz, int64_t* Fp, int64_t* Fi, double* Fx, double* Fz, int64_t* P, int64_t* Q, double* Rs,
                   int64_t* R, klu_l_common* cm) {
    (void)Y;
    if (!cm) return 0;
    if (!F) { cm->status = KLU_INVALID; return 0; }
    double* l = (double*)malloc((size_t)(2 * F->lnz + 2) * sizeof(double));
    double* u = (double*)malloc((size_t)(2 * F->unz + 2) * sizeof(double));
    double* f = (double*)malloc((size_t)(2 * F->nzoff + 2) * sizeof(double));
    if (!l || !u || !f) { free(l); free(u); free(f); cm->status = KLU_OUT_OF_MEMORY; return 0; }
    b200s_status st = b200s_klu_extract_z((b200s_klu_num*)F->b200s, Lp, Li, l, Up, Ui, u, Fp, Fi, f, P, Q, Rs, R);
    if (st == B200S_OK) {
        for (int64_t p = 0; p < F->lnz; p++) { if (Lx) Lx[p] = l[2 * p]; if (Lz) Lz[p] = l[2 * p + 1]; }
        for (int64_t p = 0; p < F->unz; p++) { if (Ux) Ux[p] = u[2 * p]; if (Uz) Uz[p] = u[2 * p + 1]; }
        for (int64_t p = 0; p < F->nzoff; p++) { if (Fx) Fx[p] = f[2 * p]; if (Fz) Fz[p] = f[2 * p + 1]; }
    }
    free(l); free(u); free(f);
    cm->status = klu_status_of(st);
    return st == B200S_OK;
}
