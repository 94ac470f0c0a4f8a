// Host side of the KLU replacement: BTF pre-ordering, per-block AMD, the pivoting Gilbert-Peierls
// factorization that fixes the pattern and pivot sequence, and the static refactorization plan
// that the batched CUDA kernels execute.
#pragma once
#include "host.hpp"

namespace b200s {

struct KluSymbolic {
    i32 n = 0;
    i64 nnz = 0;
    i32 nblocks = 0, maxblock = 0, structural_rank = 0;
    std::vector<i32> P, Q;       // row / column permutation to block upper triangular form (+ per-block AMD)
    std::vector<i32> R;          // block boundaries, nblocks+1
    std::vector<i64> Ap;         // copy of the pattern analysed (n+1)
    std::vector<i32> Ai;
};

struct KluNumeric {
    i32 n = 0;
    // all row/column indices below are in FINAL permuted numbering (row k = pivotal row k, column k = Q[k])
    std::vector<i64> Lp, Up, Fp; // n+1 each
    std::vector<i32> Li, Ui, Fi;
    std::vector<double> Lx, Ux, Fx;   // L has its unit diagonal stored first in each column; U's diagonal is last
    std::vector<i32> Pnum;       // Pnum[k] = original row of pivotal row k
    std::vector<double> Rs;      // Rs[k] = scale factor of pivotal row k  (row k of  R \ (P A Q))
    double flops = 0;            // 2 * multiply-adds of one (re)factorization
    i32 singular_col = -1;
};

// klu_l_analyze with klu_defaults: btf = 1, ordering = AMD.  Throws std::invalid_argument on bad input.
void klu_analyze(i64 n, const i64* Ap, const i64* Ai, KluSymbolic& S);
// klu_l_factor with klu_defaults: scale = 2 (max |row|), tol = 1e-3.  Returns 0, or 2 (singular).
int klu_factor(const KluSymbolic& S, const double* Ax, KluNumeric& N);

// ---- static refactorization plan (pattern + pivot order fixed) ------------------------------------
// Value slots: slot v of matrix b lives at LU[v * batch + b].  Column k of the permuted matrix owns the
// contiguous slots [cbeg[k], cbeg[k+1]): U entries above the diagonal (ascending row), the diagonal,
// then the L entries below it (ascending row).  F entries follow all columns.
struct KluPlan {
    i32 n = 0, nlevels = 0;
    i64 nslots = 0, nnzA = 0, lu_slots = 0;
    std::vector<i64> cbeg;            // n+1
    std::vector<i32> udiag_slot;      // slot of U(k,k)
    std::vector<i32> slot_src;        // per slot: index into the caller's value array, or -1 (fill-in)
    std::vector<i32> slot_row;        // per slot: pivotal row (for scaling by Rs)
    // row scaling: entries of A grouped by pivotal row
    std::vector<i64> rowptr;          // n+1
    std::vector<i32> rowent;          // indices into the caller's value array
    // level schedule of columns; per column a list of "updates": for U entry (j,k): source slot of u_jk,
    // first L slot of column j, count, and offset into dest[] (destination slots inside column k)
    std::vector<i32> level_ptr, level_cols;
    std::vector<i64> upd_ptr;         // n+1: updates of column k are upd[upd_ptr[k] .. upd_ptr[k+1])
    std::vector<i32> upd_uslot, upd_lslot, upd_cnt;
    std::vector<i64> upd_dest;        // offset into dest[]
    std::vector<i32> dest;
    // solve schedule uses Lp/Li/Up/Ui/Fp/Fi of the numeric object with slots: L(i,k) at lslot, etc.
    std::vector<i32> lslot0;          // per column: first L slot (below diagonal)
    std::vector<i32> fslot0;          // per column: first F slot
};
void klu_build_plan(const KluSymbolic& S, const KluNumeric& N, KluPlan& plan);

}  // namespace b200s
