/*
 * klu.h -- the subset of the SuiteSparse KLU C API that kvxopt's wrapper src/C/klu.c uses, implemented on top of
 * libb200sparse.so (include/b200sparse.h) by kvxopt_b200/csrc/suitesparse_shim.c, so that the reference's extension
 * module source compiles UNMODIFIED and links to the B200 engine in place of -lklu (reference setup.py:292-298).
 *
 *   klu.c call site                         -> C ABI entry point
 *   klu_l_analyze (:142,264)                -> b200s_klu_analyze
 *   klu_l_factor  (:160,337)                -> b200s_klu_factor (+ b200s_klu_extract for the members get_det reads)
 *   klu_l_solve / klu_l_tsolve (:189,651)   -> b200s_klu_solve
 *   klu_l_extract (:461)                    -> b200s_klu_extract
 *   klu_l_free_symbolic / _numeric          -> b200s_klu_free_symbolic / _numeric
 * The complex entry points (klu_zl_*) are served by b200s_klu_factor_z / _solve_z / _extract_z (complex factor of the host
 * pivot search for extract / determinant, solves on the device through the real embedding).
 */
#ifndef B200S_SHIM_KLU_H
#define B200S_SHIM_KLU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KLU_OK 0
#define KLU_SINGULAR 1
#define KLU_OUT_OF_MEMORY (-2)
#define KLU_INVALID (-3)
#define KLU_TOO_LARGE (-4)

typedef struct klu_l_common_struct {
    double tol, memgrow, initmem_amd, initmem, maxwork;
    int btf, ordering, scale, halt_if_singular;
    int status;
} klu_l_common;

typedef struct {
    int64_t n, nz, nblocks, maxblock;
    int64_t *Q, *R;              /* column permutation and block boundaries (read by get_det klu.c:766 and get_numeric) */
    void* b200s;                 /* b200s_klu_sym* */
} klu_l_symbolic;

typedef struct {
    int64_t n, nblocks, lnz, unz, nzoff;
    int64_t* Pnum;               /* final row permutation      (klu.c:765)  */
    double* Rs;                  /* row scale factors          (klu.c:767)  */
    void* Udiag;                 /* diagonal of U              (klu.c:747)  */
    void* b200s;                 /* b200s_klu_num* */
} klu_l_numeric;

int klu_l_defaults(klu_l_common*);
klu_l_symbolic* klu_l_analyze(int64_t n, int64_t* Ap, int64_t* Ai, klu_l_common*);
klu_l_numeric* klu_l_factor(int64_t* Ap, int64_t* Ai, double* Ax, klu_l_symbolic*, klu_l_common*);
int klu_l_solve(klu_l_symbolic*, klu_l_numeric*, int64_t ldim, int64_t nrhs, double* B, klu_l_common*);
int klu_l_tsolve(klu_l_symbolic*, klu_l_numeric*, int64_t ldim, int64_t nrhs, double* B, klu_l_common*);
int klu_l_free_symbolic(klu_l_symbolic**, klu_l_common*);
int klu_l_free_numeric(klu_l_numeric**, klu_l_common*);
int klu_l_extract(klu_l_numeric*, klu_l_symbolic*, int64_t* Lp, int64_t* Li, double* Lx, int64_t* Up, int64_t* Ui, double* Ux,
                  int64_t* Fp, int64_t* Fi, double* Fx, int64_t* P, int64_t* Q, double* Rs, int64_t* R, klu_l_common*);
/* complex variants: declared because klu.c calls them for 'z' matrices; they set Common->status = KLU_INVALID */
klu_l_numeric* klu_zl_factor(int64_t* Ap, int64_t* Ai, double* Ax, klu_l_symbolic*, klu_l_common*);
int klu_zl_solve(klu_l_symbolic*, klu_l_numeric*, int64_t ldim, int64_t nrhs, double* B, klu_l_common*);
int klu_zl_tsolve(klu_l_symbolic*, klu_l_numeric*, int64_t ldim, int64_t nrhs, double* B, int conj_solve, klu_l_common*);
int klu_zl_free_numeric(klu_l_numeric**, klu_l_common*);
int klu_zl_extract(klu_l_numeric*, klu_l_symbolic*, int64_t* Lp, int64_t* Li, double* Lx, double* Lz, int64_t* Up, int64_t* Ui,
                   double* Ux, double* Uz, int64_t* Fp, int64_t* Fi, double* Fx, double* Fz, int64_t* P, int64_t* Q, double* Rs,
                   int64_t* R, klu_l_common*);

#ifdef __cplusplus
}
#endif
#endif
