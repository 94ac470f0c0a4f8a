/*
 * cholmod.h -- the subset of the SuiteSparse CHOLMOD C API that kvxopt's wrapper src/C/cholmod.c uses, implemented
 * on top of libb200sparse.so (include/b200sparse.h) by kvxopt_b200/csrc/suitesparse_shim.c.
 *
 * Purpose: the reference's extension module source compiles UNMODIFIED against this header and links to the B200
 * engine in place of SuiteSparse (reference setup.py:389-394 links -lcholmod; here -lb200sparse), so `import
 * kvxopt.cholmod` is the reference's own C module with the numeric work on the GPU.  Only what cholmod.c touches is
 * declared; struct members it never reads are omitted.  Enumerators carry CHOLMOD's published values.
 *
 *   cholmod.c call site                       -> C ABI entry point
 *   cholmod_l_analyze_p   (:274,663,811)      -> b200s_chol_analyze
 *   cholmod_l_factorize   (:362,677,824)      -> b200s_chol_factorize (+ b200s_chol_diag for the members diag() reads)
 *   cholmod_l_solve       (:483,735)          -> b200s_chol_solve
 *   cholmod_l_spsolve     (:567,858)          -> b200s_chol_spsolve
 *   cholmod_l_factor_to_sparse (:969)         -> b200s_chol_get_L
 *   cholmod_l_free_factor (:213,...)          -> b200s_chol_free
 */
#ifndef B200S_SHIM_CHOLMOD_H
#define B200S_SHIM_CHOLMOD_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CHOLMOD_OK 0
#define CHOLMOD_NOT_INSTALLED (-1)
#define CHOLMOD_OUT_OF_MEMORY (-2)
#define CHOLMOD_TOO_LARGE (-3)
#define CHOLMOD_INVALID (-4)
#define CHOLMOD_GPU_PROBLEM (-5)
#define CHOLMOD_NOT_POSDEF 1
#define CHOLMOD_DSMALL 2

#define CHOLMOD_PATTERN 0
#define CHOLMOD_REAL 1
#define CHOLMOD_COMPLEX 2
#define CHOLMOD_ZOMPLEX 3

#define CHOLMOD_A 0
#define CHOLMOD_LDLt 1
#define CHOLMOD_LD 2
#define CHOLMOD_DLt 3
#define CHOLMOD_L 4
#define CHOLMOD_Lt 5
#define CHOLMOD_D 6
#define CHOLMOD_P 7
#define CHOLMOD_Pt 8

typedef struct cholmod_common_struct {
    int status;
    int print, supernodal, nmethods, postorder;   /* the members set_options (cholmod.c:87-129) writes */
    double dbound;
} cholmod_common;

typedef struct cholmod_sparse_struct {
    size_t nrow, ncol, nzmax;
    void *p, *i, *nz, *x, *z;
    int stype, itype, xtype, dtype, sorted, packed;
} cholmod_sparse;

typedef struct cholmod_dense_struct {
    size_t nrow, ncol, nzmax, d;
    void *x, *z;
    int xtype, dtype;
} cholmod_dense;

typedef struct cholmod_factor_struct {
    size_t n, minor;
    /* supernodal members read by diag() (cholmod.c:927-943).  The panels live in HBM; what is presented here is a
     * host view with one 1 x 1 "supernode" per column holding the diagonal of L (refreshed by every factorize). */
    size_t nsuper;
    void *super, *pi, *px, *x;
    int xtype, is_ll, is_super;
    void* b200s;                 /* b200s_chol* */
    int stype;
    int zfactor;                 /* analysed from a CHOLMOD_COMPLEX matrix: the b200s_chol_*_z entry points serve it */
} cholmod_factor;

int cholmod_l_start(cholmod_common*);
int cholmod_l_finish(cholmod_common*);
int cholmod_l_defaults(cholmod_common*);
cholmod_sparse* cholmod_l_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted, int packed, int stype, int xtype,
                                          cholmod_common*);
int cholmod_l_free_sparse(cholmod_sparse**, cholmod_common*);
cholmod_dense* cholmod_l_allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype, cholmod_common*);
int cholmod_l_free_dense(cholmod_dense**, cholmod_common*);
int cholmod_l_check_perm(void* Perm, size_t len, size_t n, cholmod_common*);
cholmod_factor* cholmod_l_analyze_p(cholmod_sparse* A, void* UserPerm, void* fset, size_t fsize, cholmod_common*);
int cholmod_l_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common*);
cholmod_dense* cholmod_l_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common*);
cholmod_sparse* cholmod_l_spsolve(int sys, cholmod_factor* L, cholmod_sparse* B, cholmod_common*);
cholmod_sparse* cholmod_l_factor_to_sparse(cholmod_factor* L, cholmod_common*);
int cholmod_l_free_factor(cholmod_factor**, cholmod_common*);

#ifdef __cplusplus
}
#endif
#endif
