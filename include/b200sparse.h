/*
 * b200sparse.h -- C ABI of libb200sparse.so, the B200 (sm_100a) sparse direct-solve engine that
 * kvxopt's `cholmod` and `klu` extension modules call in place of SuiteSparse.
 *
 * Every entry point names the reference interface it replaces (paths under the kvxopt source tree,
 * file:line).  All pointers are HOST pointers unless the name ends in `_dev`; all index arrays are
 * 64-bit (kvxopt's int_t = Py_ssize_t, src/C/kvxopt.h:46) compressed-column storage exactly as held
 * by a kvxopt `spmatrix` (src/C/kvxopt.h:58-69): colptr[ncols+1], rowind[nnz] sorted per column,
 * values[nnz].  Dense blocks are column-major FP64 as held by a kvxopt `matrix` (kvxopt.h:48-56).
 *
 * The numeric work (factorization, triangular solves, batched refactorization) runs in hand-written
 * CUDA kernels; there is no CPU fallback: numeric calls return B200S_NO_DEVICE when no GPU is present.
 * The symbolic analysis runs on the host and its result is uploaded once per sparsity pattern.
 */
#ifndef B200SPARSE_H
#define B200SPARSE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t b200s_int;

/* Status codes mirror the ones the reference wrappers branch on
 * (src/C/cholmod.c:366-395, 679-723; src/C/klu.c:144-183, 370-375). */
typedef enum {
    B200S_OK            = 0,
    B200S_NOT_POSDEF    = 1,   /* CHOLMOD_NOT_POSDEF: `minor` holds the failing column        */
    B200S_SINGULAR      = 2,   /* KLU_SINGULAR / zero pivot                                   */
    B200S_OUT_OF_MEMORY = -2,  /* CHOLMOD_OUT_OF_MEMORY / KLU_OUT_OF_MEMORY -> MemoryError     */
    B200S_TOO_LARGE     = -3,  /* CHOLMOD_TOO_LARGE / KLU_TOO_LARGE                            */
    B200S_INVALID       = -4,  /* CHOLMOD_INVALID / KLU_INVALID -> ValueError                  */
    B200S_NO_DEVICE     = -5,  /* no CUDA device: numeric paths refuse to run (no CPU fallback)*/
    B200S_CUDA_ERROR    = -6   /* a CUDA runtime call failed; see b200s_last_error()           */
} b200s_status;

const char* b200s_strerror(b200s_status s);
const char* b200s_last_error(void);      /* text of the last CUDA/runtime failure on this thread */
const char* b200s_version(void);
int         b200s_device_count(void);    /* 0 when no GPU / driver                                */
b200s_status b200s_set_device(int dev);  /* device used by handles created afterwards             */

/* ------------------------------------------------------------------------------------------------
 * Sparse Cholesky  (replaces the cholmod_l_* calls of src/C/cholmod.c)
 * ---------------------------------------------------------------------------------------------- */

typedef struct b200s_chol b200s_chol;    /* opaque: symbolic plan + device-resident numeric factor */

/* Mirrors the `cholmod.options` keys read by set_options (src/C/cholmod.c:87-129) plus the relaxed
 * supernode parameters CHOLMOD keeps in its Common object. */
typedef struct {
    int    supernodal;   /* 2, 1: supernodal LL^T.  0: LDL^T without pivoting (CHOLMOD's simplicial mode, cholmod.c:60-64):
                            any symmetric matrix with nonzero pivots (quasi-definite KKT systems); the kernels compute
                            P A P' = Lt S Lt', S = diag(+-1), and answer sys 1..6 / getfactor with L = Lt diag(Lt)^-1,
                            D = S diag(Lt)^2; a zero pivot returns NOT_POSDEF with `minor`.  Other values: INVALID */
    int    nmethods;     /* 0: user perm if given else AMD; 1: user perm (must be given); 2: as 0   */
    int    postorder;    /* 1: etree postorder after the fill-reducing ordering (always applied to  */
                         /*    make supernodes contiguous; 0 is accepted and ignored)               */
    double dbound;       /* lower bound for diagonal entries of L (0 = off)                         */
    int    ordering;     /* 0: AMD (own implementation); 1: natural (identity) when no user perm    */
    int    nrelax[3];    /* relaxed amalgamation thresholds, default {4,16,48}                      */
    double zrelax[3];    /* default {0.8,0.1,0.05}                                                  */
    int    block;        /* dense block-column width used inside large fronts (0 = default 128)     */
    int    max_merge_cols; /* 0 = no limit.  Amalgamation does not merge a supernode into its parent when the result would
                              be wider than this.  The multi-GPU factorization sets it (kvxopt_b200/dist.py) so that a top-level
                              separator is not absorbed by its parent -- a free merge (same structure), but the merged front has
                              no Schur complement left to share between GPUs */
} b200s_chol_opts;

void b200s_chol_default_opts(b200s_chol_opts* o);

/* cholmod.symbolic (src/C/cholmod.c:244-291): pack(A,uplo) :132-181 + cholmod_l_analyze_p :269.
 * Only the `uplo` triangle of (colptr,rowind) is referenced.  perm may be NULL. Host only. */
b200s_status b200s_chol_analyze(b200s_int n, const b200s_int* colptr, const b200s_int* rowind,
                                char uplo, const b200s_int* perm, const b200s_chol_opts* opts,
                                b200s_chol** out);

/* cholmod.numeric (src/C/cholmod.c:322-398): cholmod_l_factorize :362.  (colptr, rowind, val) is the caller's matrix as
 * the reference rebuilds it on every call (pack :132-181).  Its pattern is compared with the one given to analyze (one
 * memcmp on the fast path); a different pattern is re-mapped through A's own indices: analysed entries A does not have
 * are zero (a subset pattern is accepted, as by CHOLMOD), an entry of the referenced triangle outside the analysed pattern
 * returns B200S_INVALID.  colptr = rowind = NULL: the caller guarantees the analysed pattern and order (values only).
 * On B200S_NOT_POSDEF *minor_out is the first non-positive pivot column (in permuted order, like
 * L->minor, cholmod.c:376-380). */
b200s_status b200s_chol_factorize(b200s_chol* F, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                  b200s_int* minor_out);
/* Same with val already resident in device memory (HBM); used by bench `value`. */
b200s_status b200s_chol_factorize_dev(b200s_chol* F, const double* val_dev, b200s_int* minor_out);

/* cholmod.solve (src/C/cholmod.c:429-499) -- all nrhs columns in one call instead of the
 * reference's per-column loop :481-493.  sys: 0 A x=b, 4 L x=b, 5 L^T x=b, 7 x=P b, 8 x=P^T b;
 * with LL^T D=I so 1 (LDL^T x=b) = 0-without-permutation, 2 (LD x=b) = 4, 3 (DL^T x=b) = 5, 6 = copy
 * (cholmod.c:437-439).  B is n x nrhs column-major with leading dimension ldB, overwritten. */
b200s_status b200s_chol_solve(b200s_chol* F, int sys, double* B, b200s_int nrhs, b200s_int ldB);
b200s_status b200s_chol_solve_dev(b200s_chol* F, int sys, double* B_dev, b200s_int nrhs, b200s_int ldB);

/* cholmod.spsolve (src/C/cholmod.c:524-587): sparse right-hand side, sparse result holding the
 * numerically nonzero entries.  Output arrays are allocated by the library; release with b200s_free. */
b200s_status b200s_chol_spsolve(b200s_chol* F, int sys, b200s_int nrows, b200s_int ncols,
                                const b200s_int* Bp, const b200s_int* Bi, const double* Bx,
                                b200s_int** Xp, b200s_int** Xi, double** Xx);

/* ---- complex Hermitian matrices ('z' spmatrix, src/C/cholmod.c:144,153,343-357: CHOLMOD_COMPLEX) -------------------------
 * Factored through the real symmetric embedding a + ib -> [[a, -b], [b, a]] of order 2n (rows/columns 2i, 2i+1 per complex
 * index i, pairs kept adjacent by the ordering), whose Cholesky factor is the embedding of the complex factor.  Values are
 * (re, im) pairs as a kvxopt 'z' matrix stores them.  A complex dense right-hand side IS its real embedding: solve it with
 * b200s_chol_solve(F, sys, (double*)B, nrhs, 2*ldB).  Same status codes; *minor is a complex column index. */
b200s_status b200s_chol_analyze_z(b200s_int n, const b200s_int* colptr, const b200s_int* rowind, char uplo, const b200s_int* perm,
                                  const b200s_chol_opts* opts, b200s_chol** out);
b200s_status b200s_chol_factorize_z(b200s_chol* F, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                    b200s_int* minor_out);
b200s_status b200s_chol_spsolve_z(b200s_chol* F, int sys, b200s_int nrows, b200s_int ncols, const b200s_int* Bp,
                                  const b200s_int* Bi, const double* Bx, b200s_int** Xp, b200s_int** Xi, double** Xx);
b200s_status b200s_chol_diag_z(b200s_chol* F, double* d_out);      /* n (re, im) pairs, im = 0 */
b200s_status b200s_chol_get_L_z(b200s_chol* F, b200s_int** Lp, b200s_int** Li, double** Lx);

/* cholmod.diag (src/C/cholmod.c:900-945): the n diagonal entries of L, in factor (permuted) order. */
b200s_status b200s_chol_diag(b200s_chol* F, double* d_out);

/* cholmod.getfactor (src/C/cholmod.c:948-985): L as CCS (cholmod_l_factor_to_sparse), zeros inside
 * relaxed supernodes are dropped.  Arrays are library-allocated; release with b200s_free. */
b200s_status b200s_chol_get_L(b200s_chol* F, b200s_int** Lp, b200s_int** Li, double** Lx);

typedef struct {
    b200s_int n, nsuper, nnz_L, nnz_A, nlevels, max_front_rows, max_front_cols;
    b200s_int factor_bytes, workspace_bytes;
    double    flops;          /* sum over supernodes of c^3/3 + c^2 r + c r^2 (r = rows below)      */
    double    flops_potrf, flops_trsm, flops_syrk;
    int       is_numeric;     /* 0 after analyze, 1 after a successful factorize                    */
    b200s_int minor;          /* n when the last factorization succeeded                            */
    /* device-event timings (ms) of the last factorize / solve call, 0 when not measured            */
    double    ms_h2d, ms_assemble, ms_factor, ms_total, ms_solve;
    double    ms_analyze;     /* host wall time of analyze                                          */
    double    ms_dense_update; /* device time inside the SYRK/GEMM update kernels of the last factorize */
    double    ms_potrf, ms_trsm, ms_extend;
    double    flops_update;   /* flops executed by the tiled DMMA update kernel (numerator of its roofline) */
    b200s_int zn;             /* complex order for factor objects made by b200s_chol_analyze_z (then n = 2 zn), else 0 */
} b200s_chol_info_t;
b200s_status b200s_chol_info(const b200s_chol* F, b200s_chol_info_t* info);
/* when on, factorize records per-kernel-class CUDA events (adds launch gaps; for profiling only) */
b200s_status b200s_chol_set_profiling(b200s_chol* F, int on);

/* L->Perm (fill-reducing + postorder): perm_out[k] = original index of row/column k of P A P^T. */
b200s_status b200s_chol_get_perm(const b200s_chol* F, b200s_int* perm_out);
/* supernode partition for tests / tools: super[nsuper+1] first columns, rowptr[nsuper+1], rows[] */
b200s_status b200s_chol_get_super(const b200s_chol* F, b200s_int* super_out, b200s_int* rowptr_out,
                                  b200s_int* rows_out /* may be NULL to query sizes only */);

/* ---- building blocks of the multi-GPU (subtree-to-subcube) factorization; no reference counterpart ----------
 * One process per GPU holds the same plan.  b200s_chol_set_owned marks the fronts this process factors (bytes,
 * nsuper of them; NULL = all).  The factorization is stepped level by level: begin (upload + assemble), one call
 * per level (kernels are enqueued on the handle's stream; non-owned fronts are skipped), end (synchronise, minor).
 * Between levels the caller moves update matrices of fronts whose parent lives on another GPU: they sit at
 * W_dev[uoff[s] .. uoff[s]+usize[s]) on every process (identical layout), panels at L_dev[loff[s] .. +lsize[s]).
 * kvxopt_b200/dist.py drives this with torch.distributed (NCCL send/recv over NVLink). */
b200s_status b200s_chol_set_owned(b200s_chol* F, const unsigned char* owned);
b200s_status b200s_chol_factor_begin(b200s_chol* F, const double* val, int val_on_device);
b200s_status b200s_chol_factor_level(b200s_chol* F, b200s_int level);
b200s_status b200s_chol_factor_end(b200s_chol* F, b200s_int* minor_out);
b200s_status b200s_chol_sync(b200s_chol* F);             /* wait for the handle's stream */
/* A level in two steps, for fronts whose Schur complement C = -L21 L21' is SHARED by several GPUs (the top separators of the
 * tree, which would otherwise run on one GPU while the others idle): phase 1 = extend-add, small fronts, panels and in-panel
 * updates of the owned fronts; then the caller copies the factored panel L_dev[loff[s] .. +lsize[s]) of every shared front
 * to its helpers; phase 2 = the Schur complements; phase 3 = both (= b200s_chol_factor_level).
 * b200s_chol_set_syrk_split: per front (arrays of nsuper) own[s] != 0: this process computes the column tiles
 * [tile_lo[s], tile_hi[s]) (tiles of 64 columns of the update matrix) -- in place in W_dev when base[s] == LLONG_MIN (the
 * front's owner: W already holds the children's contributions), else into scratch_dev[base[s] ..) (a helper: the buffer
 * must be zero; column tile tile_lo[s] first, leading dimension as in W, i.e. the even number >= nrows - ncols + (ncols & 1)).
 * own == NULL restores the default (every owned front complete and in place). */
b200s_status b200s_chol_factor_level_phase(b200s_chol* F, b200s_int level, int phase);
b200s_status b200s_chol_set_syrk_split(b200s_chol* F, const unsigned char* own, const int* tile_lo, const int* tile_hi,
                                       const long long* base, double* scratch_dev);
/* per front (arrays of nsuper, any may be NULL); offsets and sizes in doubles */
b200s_status b200s_chol_front_layout(const b200s_chol* F, b200s_int* parent, b200s_int* level, b200s_int* ncols,
                                     b200s_int* nrows, b200s_int* loff, b200s_int* lsize, b200s_int* uoff,
                                     b200s_int* usize);
b200s_status b200s_chol_device_buffers(b200s_chol* F, double** L_dev, double** W_dev);
/* Distributed triangular solves (one right-hand side; LL' factors): the panels stay on the GPUs that factored them.
 *   begin: X = P b (b_dev: n doubles on this device; every process passes the same b);
 *   level (backward = 0, levels 0 .. nlevels-1): the owned fronts of the level gather their pivots from X and their children's
 *     update vectors from the work-vector buffer T, solve, leave y in X[col0 ..] and their own update vector in
 *     T[rowptr[s] + ncols[s] .. rowptr[s] + nrows[s]) -- before the call the caller copies the update vectors of children owned
 *     by another process to the same place in its T (identical layout everywhere);
 *   level (backward = 1, levels nlevels-1 .. 0): the owned fronts gather x at their row lists from X and write x(columns) to
 *     X[col0[s] .. col0[s] + ncols[s]) -- before the call the caller copies the solution entries of ancestors owned elsewhere;
 *   end: x_dev = P' X (meaningful where X is complete: the caller collects the column ranges first).
 * b200s_chol_solve_buffers returns T and X (device pointers, valid until the factor object is freed or solves with more
 * right-hand-side columns grow the workspace); b200s_chol_front_layout2 the per-front offsets into them. */
b200s_status b200s_chol_solve_dist_begin(b200s_chol* F, const double* b_dev);
b200s_status b200s_chol_solve_dist_level(b200s_chol* F, int backward, b200s_int level);
b200s_status b200s_chol_solve_dist_end(b200s_chol* F, double* x_dev);
b200s_status b200s_chol_solve_buffers(b200s_chol* F, double** T_dev, double** X_dev);
b200s_status b200s_chol_front_layout2(const b200s_chol* F, b200s_int* rowptr, b200s_int* col0);
/* after the panels of all fronts have been gathered into L_dev: declare the factor numeric so that solves run */
b200s_status b200s_chol_set_numeric(b200s_chol* F, int numeric, b200s_int minor);

void b200s_chol_free(b200s_chol* F);     /* capsule destructor (src/C/cholmod.c:210-214) */
void b200s_free(void* p);

/* Geometric nested-dissection permutation of an nx*ny*nz grid numbered x-fastest (BASELINE config 4).
 * perm_out[k] = grid index eliminated k-th; pass it as `perm` to b200s_chol_analyze. */
b200s_status b200s_grid_nd_perm(b200s_int nx, b200s_int ny, b200s_int nz, b200s_int leaf,
                                b200s_int* perm_out);
/* Engine extension: which triangular sweeps of one-right-hand-side solves run as ONE persistent kernel per level over the
 * level's large fronts (bit 0: forward, bit 1: backward; -1: the default = both, or B200S_SOLVE_PERSIST) instead of two
 * launches per 128-column block step.  Both paths give bit-identical solutions; the launch-per-step path also serves several
 * right-hand sides and the ownership-masked distributed solves. */
b200s_status b200s_chol_set_solve_sweeps(b200s_chol* F, int mode);
/* Test hook (host only, no device needed): replays on the CPU the static work lists and the flag / counter waits of the
 * persistent solve sweeps (k_fwd_persist / k_bwd_persist: one kernel per level runs every 128-column block step of the level's
 * large fronts, one right-hand side) for one front of nr rows and nc pivot columns on nctas CTAs.  0: every (block, tile) is
 * applied exactly once and in the order of the launch-per-step kernels, every block is solved once and only after the rows it
 * reads, and no CTA is left waiting; > 0: the first rule broken; -1: invalid arguments. */
int b200s_persist_schedule_check(b200s_int nr, b200s_int nc, b200s_int nctas);
/* Test hook (host only): the child lists of the numeric phase -- per assembly item (k_extend_add: a column range of a front) and
 * per forward-gather chunk (k_fwd_gather: 512 rows of a front) the children that reach into it, which is what lets a front
 * with thousands of one-entry children (the 3 x 3 KKT matrices of kkt.ldl) be assembled without every CTA walking all of them
 * -- rebuilt for EVERY front of the analysed plan and compared with their definition.  B200S_OK or B200S_INVALID (b200s_last_error). */
b200s_status b200s_chol_child_lists_check(const b200s_chol* F);
/* AMD ordering of the symmetric pattern of the `uplo` triangle (src/C/amd.c `order`, host only). */
b200s_status b200s_amd_order(b200s_int n, const b200s_int* colptr, const b200s_int* rowind,
                             char uplo, b200s_int* perm_out);

/* ------------------------------------------------------------------------------------------------
 * KLU  (replaces the klu_l_* calls of src/C/klu.c)
 * ---------------------------------------------------------------------------------------------- */

typedef struct b200s_klu_sym b200s_klu_sym;   /* BTF + per-block ordering (klu_l_symbolic)          */
typedef struct b200s_klu_num b200s_klu_num;   /* L, U, F, pivots, row scaling (klu_l_numeric) plus   */
                                              /* the device-resident refactor plan                   */

/* klu.symbolic (src/C/klu.c:242-291): klu_l_analyze with klu_defaults (btf=1, AMD, scale=2). Host. */
b200s_status b200s_klu_analyze(b200s_int n, const b200s_int* colptr, const b200s_int* rowind,
                               b200s_klu_sym** out);
/* klu.numeric (src/C/klu.c:310-379): klu_l_factor -- threshold partial pivoting (tol 1e-3, diagonal
 * preferred), row scaling by max |row|.  The pivot search runs on the host once per pattern; the
 * resulting pattern + pivot sequence is what the batched device refactorization reuses. */
b200s_status b200s_klu_factor(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind,
                              const double* val, b200s_klu_num** out);

/* The pivot search of b200s_klu_factor alone (host only, no device upload): pattern of L, U, F, the pivot
 * order and the row scaling of ONE matrix.  It exists so that the host logic can be verified without a
 * GPU and so that tools can inspect the pivot sequence; solves on such an object return B200S_NO_DEVICE.
 * b200s_klu_extract_host returns the values this pivot search computed as a by-product. */
b200s_status b200s_klu_pivot_host(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind,
                                  const double* val, b200s_klu_num** out);
b200s_status b200s_klu_extract_host(const b200s_klu_num* N, double* Lx, double* Ux, double* Fx, double* Rs);

/* Batched numeric refactorization (klu_refactor semantics: same pattern, same pivot sequence, fresh
 * row scaling) of `batch` matrices whose values are vals[b*ldv + k], k indexing the CCS given to
 * b200s_klu_factor.  Factors stay resident on the device for b200s_klu_solve_batch.
 * status_per_matrix[b] = B200S_OK or B200S_SINGULAR (zero/NaN pivot).  May be NULL. */
b200s_status b200s_klu_refactor_batch(b200s_klu_num* N, const double* vals, b200s_int batch,
                                      b200s_int ldv, int* status_per_matrix);
b200s_status b200s_klu_refactor_batch_dev(b200s_klu_num* N, const double* vals_dev, b200s_int batch,
                                          b200s_int ldv, int* status_per_matrix);
/* Pipelined form of b200s_klu_refactor_batch for callers that stream many batches: _begin enqueues the upload of
 * `vals` (host memory, ideally pinned; it must stay valid until the matching _end) and the refactorization kernels
 * and returns at once; _end waits for the OLDEST batch in flight and returns its per-matrix status.  Up to two
 * batches may be in flight: the upload of batch i+1 then overlaps the kernels of batch i.  The device factors
 * always hold the most recently begun batch (solve after the matching _end and before the next _begin). */
b200s_status b200s_klu_refactor_batch_begin(b200s_klu_num* N, const double* vals, b200s_int batch, b200s_int ldv);
b200s_status b200s_klu_refactor_batch_end(b200s_klu_num* N, int* status_per_matrix);
/* klu.solve (src/C/klu.c:593-690) for every matrix of the last refactored batch:
 * B is batch blocks of n x nrhs column-major (block stride ldB*nrhs), overwritten.  trans 0 = 'N', 1 = 'T'. */
b200s_status b200s_klu_solve_batch(b200s_klu_num* N, int trans, double* B, b200s_int nrhs,
                                   b200s_int ldB, b200s_int batch);
b200s_status b200s_klu_solve_batch_dev(b200s_klu_num* N, int trans, double* B_dev, b200s_int nrhs,
                                       b200s_int ldB, b200s_int batch);
/* klu.solve for the single matrix given to b200s_klu_factor (a batch of one on the device). */
b200s_status b200s_klu_solve(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB);

typedef struct {
    b200s_int n, nblocks, nnz_A, nnz_L, nnz_U, nnz_F, nlevels, max_block;
    double    flops;             /* multiply-adds x2 of one refactorization                        */
    b200s_int bytes_per_refactor;/* 8*(nnz_A + nnz_L + nnz_U + nnz_F + 2n): compulsory traffic      */
    double    ms_h2d, ms_refactor, ms_solve;  /* device-event times of the last batch call          */
    double    ms_kernel;         /* of which: the sparse refactorization kernel (k_klu_refactor_wave) */
    double    ms_dense;          /* ... and the dense trailing block (pack + k_klu_dense_lu + unpack) */
    b200s_int launches;          /* kernels launched by the last refactor_batch call                 */
} b200s_klu_info_t;
b200s_status b200s_klu_info(const b200s_klu_num* N, b200s_klu_info_t* info);

/* klu.get_numeric (src/C/klu.c:392-566): klu_l_extract of the host factor.  Sizes come from
 * b200s_klu_info; arrays are caller-allocated: Lp[n+1],Li[nnz_L],Lx; Up[n+1],Ui[nnz_U],Ux;
 * Fp[n+1],Fi[nnz_F],Fx; P[n],Q[n],Rs[n] (the scale factors themselves, NOT inverted), R[nblocks+1].
 * Any pointer may be NULL. */
b200s_status b200s_klu_extract(const b200s_klu_num* N,
                               b200s_int* Lp, b200s_int* Li, double* Lx,
                               b200s_int* Up, b200s_int* Ui, double* Ux,
                               b200s_int* Fp, b200s_int* Fi, double* Fx,
                               b200s_int* P, b200s_int* Q, double* Rs, b200s_int* R);
/* download the factor of matrix b of the last batch in the same layout (values only; pattern/P/Q shared) */
b200s_status b200s_klu_extract_batch(b200s_klu_num* N, b200s_int b, double* Lx, double* Ux,
                                     double* Fx, double* Rs);

/* Read-only view of the static refactorization plan (pattern + pivot order frozen by b200s_klu_factor)
 * that the batched kernels replay; pointers stay valid while N lives.  For tests and tools.
 * Value slot v of matrix b lives at LU[v*batch_padded + b]; column k owns slots [cbeg[k], cbeg[k+1]):
 * U above the diagonal (ascending rows), U(k,k) at udiag_slot[k], L below the diagonal from lslot0[k];
 * F entries of column k start at fslot0[k].  slot_src[v] indexes the caller's value array (-1 = fill-in),
 * slot_row[v] is the pivotal row whose scale factor divides it.  Column k applies, for each u in
 * [upd_ptr[k], upd_ptr[k+1]): LU[dest[upd_dest[u]+t]] -= LU[upd_lslot[u]+t] * LU[upd_uslot[u]], t < upd_cnt[u]. */
typedef struct {
    b200s_int n, nlevels, nslots, lu_slots, nnz_A, nupd, ndest;
    const int64_t *cbeg, *rowptr, *upd_ptr, *upd_dest;
    const int32_t *udiag_slot, *slot_src, *slot_row, *rowent, *level_ptr, *level_cols, *upd_uslot, *upd_lslot,
                  *upd_cnt, *dest, *lslot0, *fslot0;
    /* wave schedule of the fast kernel (klu_gpu.cu): statistics */
    b200s_int nwaves, nwaves_with_deps, nbatches, nsegments, staged_rows;
    b200s_int npieces, npiece_users;     /* staged source blocks (<= 4 columns of one L supernode) x row pieces; (piece, user column) pairs */
    b200s_int wave_ok;                   /* 0: pattern outside the wave kernel's budget (the level-schedule kernel factors it) */
    b200s_int nearly, nearly_levels;     /* columns factored level by level before the waves (k_klu_early) and their levels */
} b200s_klu_plan_view_t;
b200s_status b200s_klu_plan_view(const b200s_klu_num* N, b200s_klu_plan_view_t* view);
/* Replays the wave-schedule tables of the plan on the HOST for one matrix with the values `val` and returns the factor in the
 * layout of b200s_klu_extract_host.  Verification of the host-built plan in CPU tests; not a factorization path. */
b200s_status b200s_klu_plan_emulate_host(const b200s_klu_num* N, const double* val, double* Lx, double* Ux, double* Fx, double* Rs);

/* Replays on the HOST the operation tape of the one-matrix solve kernel (klu_solve / klu_tsolve, klu.c:593-690, written down
 * once per pattern as column operations in execution order) with the values of the pivot search, in place on B (n x nrhs,
 * leading dimension ldB); trans: 0 = 'N', 1 = 'T'.  Verification of the host-built tape in CPU tests; not a solve path. */
b200s_status b200s_klu_solve_tape_host(const b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB);

/* ---- complex matrices ('z'; klu_zl_*: src/C/klu.c:161-162,348-355,468-479,661-668,754-813) ------------------------------------
 * b200s_klu_analyze is type-independent (klu.c:266).  val: nnz (re, im) pairs.  The threshold-pivoting factorization runs on
 * the host in complex arithmetic and is what b200s_klu_extract_z returns (get_numeric, get_det); the solves run on the device
 * through the real embedding a + ib -> [[a, -b], [b, a]] of order 2n, whose factor object b200s_klu_embedded exposes for the
 * batched entry points.  b200s_klu_solve_z: trans 0 = A x = b, 1 = A^T x = b, 2 = A^H x = b; B holds (re, im) pairs, ldB counts
 * complex numbers.  b200s_klu_info reports the counts of the complex factor. */
b200s_status b200s_klu_factor_z(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                b200s_klu_num** out);
b200s_status b200s_klu_solve_z(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB);
b200s_status b200s_klu_extract_z(const b200s_klu_num* N, b200s_int* Lp, b200s_int* Li, double* Lx, b200s_int* Up, b200s_int* Ui,
                                 double* Ux, b200s_int* Fp, b200s_int* Fi, double* Fx, b200s_int* P, b200s_int* Q, double* Rs,
                                 b200s_int* R);
b200s_klu_num* b200s_klu_embedded(b200s_klu_num* N);

/* the pre-ordering held by a symbolic object (klu_symbolic's P, Q, R: BTF + per-block AMD, before any numeric pivoting):
 * P[n], Q[n], R[nblocks+1]; any may be NULL; *nblocks may be NULL.  For tests and tools. */
b200s_status b200s_klu_symbolic_perm(const b200s_klu_sym* S, b200s_int* P, b200s_int* Q, b200s_int* R, b200s_int* nblocks);
void b200s_klu_free_symbolic(b200s_klu_sym* S);   /* src/C/klu.c:51-61 */
void b200s_klu_free_numeric(b200s_klu_num* N);    /* src/C/klu.c:63-72 */

/* ---- device-side reduced KKT solver for LP/QP cones (the GPU counterpart of misc.kkt_chol2, -----------------------
 * reference src/python/misc.py:1352-1567; plugged in through the reference's kktsolver callable API,
 * src/python/coneprog.py:323-345).  G: ml x n, A: p x n (p may be 0, then Ap/Ai/Ax may be NULL), both CCS with sorted
 * row indices; Hp/Hi: pattern of the n x n matrix H whose lower triangle enters S (NULL: no H, an LP).
 * b200s_kkt_factor(di, Hx): S = H + G' diag(di)^2 G (+ A'A in singular mode) is assembled on the device in a fixed
 * pattern and factored there; K = A S^-1 A' as well.  NOT_POSDEF with *minor = failing column (of S, or of K).
 * b200s_kkt_solve(x, y, z): misc.py:1489-1565 in place on host vectors (one upload, one download):
 *   on entry bx, by, bz; on exit ux, uy, W*uz. */
typedef struct b200s_kkt b200s_kkt;
b200s_status b200s_kkt_create(b200s_int n, b200s_int ml, b200s_int p, const b200s_int* Gp, const b200s_int* Gi,
                              const double* Gx, const b200s_int* Ap, const b200s_int* Ai, const double* Ax,
                              const b200s_int* Hp, const b200s_int* Hi, b200s_kkt** out);
b200s_status b200s_kkt_set_singular(b200s_kkt* K, int on);    /* misc.py:1427-1447: S += A'A from now on */
b200s_status b200s_kkt_factor(b200s_kkt* K, const double* di, const double* Hx, b200s_int* minor);
b200s_status b200s_kkt_solve(b200s_kkt* K, double* x, double* y, double* z);
typedef struct {
    b200s_int n, ml, p, nnz_S, nterms, nnz_L, singular_mode;
    double flops, ms_assemble, ms_factor, ms_solve;
    const b200s_int *Sp, *Si;       /* lower-triangle pattern of S (valid while K lives) */
} b200s_kkt_info_t;
b200s_status b200s_kkt_info(const b200s_kkt* K, b200s_kkt_info_t* info);
/* evaluates the assembly term lists on the host (verification of the host-built plan in CPU tests; not a solve path) */
b200s_status b200s_kkt_plan_check_host(const b200s_kkt* K, const double* di, const double* Hx, double* Sx);
void b200s_kkt_free(b200s_kkt* K);

/* ---- value assembler: y = M w on the device, M a fixed sparse matrix in CSR ------------------------------------------------------
 * For the KKT plug-ins that assemble their matrix outside the library (kvxopt_b200.kkt.ldl2: the dense syrk / gemm assembly of
 * misc.kkt_ldl2, misc.py:1160-1178, as a sparse linear map from w = [W['di']^2; values of H; values of A] to the stored entries
 * of K).  create uploads M once per pattern; apply uploads w (ncols doubles, host), evaluates y (nrows) on the device -- one
 * thread per row, terms in list order -- and returns the device pointer (valid until the next apply or free), ready for
 * b200s_chol_factorize_dev; get copies the last y to the host (tests). */
typedef struct b200s_spmv b200s_spmv;
b200s_status b200s_spmv_create(b200s_int nrows, b200s_int ncols, const b200s_int* rowptr, const b200s_int* colind, const double* val,
                               b200s_spmv** out);
b200s_status b200s_spmv_apply(b200s_spmv* M, const double* w_host, double** y_dev_out);
b200s_status b200s_spmv_get(b200s_spmv* M, double* y_host);
void b200s_spmv_free(b200s_spmv* M);

/* ---- dense reduced KKT solver: the GPU counterpart of misc.kkt_chol, the 'chol' kktsolver ----------------------------
 * (reference src/python/misc.py:1213-1349).  G ml x n and A p x n are dense column-major.  create: QR of A' on the host
 * once per problem (misc.py:1246-1251), kept as compact WY.  factor(di, H): K = [Q1 Q2]'(H + G' diag(di)^2 G)[Q1 Q2] and the
 * dense Cholesky factorization of its (2,2) block of order n-p on the device (misc.py:1258-1282); H is dense n x n
 * column-major (lower triangle referenced) or NULL.  solve: misc.py:1284-1345 in place on host vectors. */
typedef struct b200s_kktd b200s_kktd;
b200s_status b200s_kktd_create(b200s_int n, b200s_int ml, b200s_int p, const double* G, const double* A, b200s_kktd** out);
b200s_status b200s_kktd_factor(b200s_kktd* K, const double* di, const double* H, b200s_int* minor);
b200s_status b200s_kktd_solve(b200s_kktd* K, double* x, double* y, double* z);
typedef struct {
    b200s_int n, ml, p, launches;
    double flops, ms_factor, ms_solve;
} b200s_kktd_info_t;
b200s_status b200s_kktd_info(const b200s_kktd* K, b200s_kktd_info_t* info);
/* the host QR of A' (tests): V n x p unit lower trapezoid, T p x p and R p x p upper triangular; Q = I - V T V' */
b200s_status b200s_kktd_get_qr(const b200s_kktd* K, double* V, double* T, double* R);
void b200s_kktd_free(b200s_kktd* K);

#ifdef __cplusplus
}
#endif
#endif /* B200SPARSE_H */
