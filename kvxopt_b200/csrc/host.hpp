// Host-side (CPU, integer) part of the engine: orderings, elimination tree, supernodes, the
// "pattern plan" that is uploaded once per sparsity pattern.  No CUDA in this header.
#pragma once
#include <cstdint>
#include <vector>
#include <string>

namespace b200s {

typedef int64_t i64;
typedef int32_t i32;

// ---- ordering.cpp -------------------------------------------------------------------------------
// Symmetric pattern (both triangles, no diagonal) in CSR/CSC form.
struct SymPattern {
    i32 n = 0;
    std::vector<i64> ptr;   // n+1
    std::vector<i32> idx;   // neighbours, unsorted
};
// build A+A' pattern from one triangle of a CCS matrix (64-bit host arrays as kvxopt stores them)
SymPattern sym_pattern_from_triangle(i64 n, const i64* colptr, const i64* rowind, char uplo);
// approximate minimum degree ordering (quotient graph, element absorption, supervariables,
// mass elimination).  Returns perm: perm[k] = node eliminated k-th.
std::vector<i32> amd_order(const SymPattern& G);
// geometric nested dissection on an nx*ny*nz grid (x fastest)
std::vector<i32> grid_nd(i64 nx, i64 ny, i64 nz, i64 leaf);

// ---- symbolic.cpp -------------------------------------------------------------------------------
// tiling constants shared by the plan statistics and the CUDA schedule (chol_gpu.cu)
constexpr int PLAN_NB = 128;        // block-column width inside large fronts
constexpr int PLAN_SMALL_NR = 128;  // fronts with nr <= this are factored by one CTA in shared memory

struct CholOpts {
    int supernodal = 2, nmethods = 0, postorder = 1, ordering = 0;
    double dbound = 0.0;
    int nrelax[3] = {4, 16, 48};
    double zrelax[3] = {0.8, 0.1, 0.05};
    int block = 128;
    int max_merge_cols = 0;   // relaxed amalgamation never builds a supernode wider than this (0 = no limit); see b200sparse.h
};

// One front (= supernode) of the multifrontal plan.
struct Front {
    i32 col0;      // first column (permuted numbering)
    i32 nc;        // number of pivot columns
    i32 nr;        // number of rows of the front (>= nc); rows[rowptr..rowptr+nr) sorted, first nc = pivots
    i32 ld;        // leading dimension of the panel in L storage (nr rounded up to even)
    i64 rowptr;    // offset into rows[]
    i64 loff;      // offset (in doubles) of the nr x nc panel in L storage
    i32 parent;    // parent front or -1
    i32 level;     // 0 = leaf; parent.level > child.level
    i64 uoff;      // offset (in doubles) of the update matrix in the workspace.  Its storage origin is front row
                   // r0 = nc - (nc&1) (even), size mu = nr - r0, ld = mu rounded up to even: update entry (i,j),
                   // 0 <= j <= i < nr-nc, lives at uoff + (i+uo) + (j+uo)*ldu with uo = nc&1.
    i64 reloff;    // offset into rel[]: rel[reloff + i] = position of update row i in the parent's row list
};

struct CholPlan {
    i32 n = 0;
    i64 nnzA = 0;                 // entries of the caller's CCS (all of them, both triangles)
    std::vector<i32> perm, iperm; // perm[k] = original index of permuted row k
    std::vector<i32> parent;      // etree (permuted numbering)
    std::vector<i32> colcount;    // column counts of L (permuted, before relaxation)
    std::vector<Front> fronts;    // postordered
    std::vector<i32> rows;        // concatenated row lists
    std::vector<i32> rel;         // concatenated relative indices (child update row -> parent row position)
    std::vector<i32> sn_of_col;   // front owning each permuted column
    std::vector<i64> amap;        // for every entry k of the caller's CCS: offset in L storage or -1 (ignored)
    std::vector<i32> child_ptr, child_idx;  // children of each front, ascending
    std::vector<i32> level_ptr, level_fronts; // fronts grouped by level
    i64 lsize = 0;                // doubles in L storage
    i64 wsize = 0;                // doubles in the update workspace (after lifetime planning)
    i64 nnzL = 0;                 // structural nonzeros of L (trapezoids, incl. relaxed zeros)
    i32 nlevels = 0, max_nr = 0, max_nc = 0;
    double flops = 0, flops_potrf = 0, flops_trsm = 0, flops_syrk = 0;
    double flops_update = 0;      // flops executed by the tiled DMMA update kernel (fronts with nr > PLAN_SMALL_NR)
    double ms_analyze = 0;
};

// throws std::invalid_argument on bad input (bad perm, unsorted/duplicate rows, out-of-range index)
void chol_analyze(i64 n, const i64* colptr, const i64* rowind, char uplo, const i64* user_perm,
                  const CholOpts& opts, CholPlan& plan);

}  // namespace b200s
