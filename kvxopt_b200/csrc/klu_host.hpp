// Host side of the KLU replacement: BTF pre-ordering, per-block AMD, the pivoting Gilbert-Peierls
// factorization that fixes the pattern and pivot sequence, and the static refactorization plan
// that the batched CUDA kernels execute.
#pragma once
#include "host.hpp"
#include <complex>
#include <cstdint>

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

// T = double (klu_l_*) or std::complex<double> (klu_zl_*: src/C/klu.c:161-162,348-355)
template <class T>
struct KluNumericT {
    i32 n = 0;
    // all row/column indices below are in FINAL permuted numbering (row k = pivotal row k, column k = Q[k])
    std::vector<i64> Lp, Up, Fp; // n+1 each
    std::vector<i32> Li, Ui, Fi;
    std::vector<T> Lx, Ux, Fx;        // L has its unit diagonal stored first in each column; U's diagonal is last
    std::vector<i32> Pnum;       // Pnum[k] = original row of pivotal row k
    std::vector<double> Rs;      // Rs[k] = scale factor of pivotal row k  (row k of  R \ (P A Q))
    double flops = 0;            // 2 * multiply-adds of one (re)factorization
    i32 singular_col = -1;
};
using KluNumeric = KluNumericT<double>;
using KluNumericZ = KluNumericT<std::complex<double>>;

// klu_l_analyze with klu_defaults: btf = 1, ordering = AMD.  Throws std::invalid_argument on bad input.
void klu_analyze(i64 n, const i64* Ap, const i64* Ai, KluSymbolic& S);
// klu_l_factor with klu_defaults: scale = 2 (max |row|), tol = 1e-3.  Returns 0, or 2 (singular).
int klu_factor(const KluSymbolic& S, const double* Ax, KluNumeric& N);
// the same for complex values (|z| = hypot as KLU's ABS, row scale = max |z| of the row); host only: the complex factor serves
// get_numeric / get_det, the solves run on the device through the real embedding (klu_capi.cu)
int klu_factor_z(const KluSymbolic& S, const std::complex<double>* Ax, KluNumericZ& N);

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
    // wave schedule (the fast kernel): columns 0..n-1 in order, grouped into waves of consecutive columns, one
    // warp per column, the column's slots held in shared memory (rows x 32 matrices).  A wave is bounded by the
    // warp count and by the shared-memory rows.  Updates of column k split at upd_split[k]: sources in earlier
    // waves (applied independently by every warp) / sources inside the wave (applied in pivot order).
    // EARLY columns: the columns in the first, wide levels of the dependency graph (row and column dependencies: level[k] >
    // level[j] whenever U(j,k) != 0 or L(k,j) != 0) -- thousands of tiny, mutually independent columns per level.  They are
    // factored level by level by k_klu_early (one warp per (column, group of 32 matrices), the whole GPU busy, HBM-bound)
    // before the wave kernel runs; the waves below cover the remaining columns only (ne_cols, in pivot order).  Because the
    // early set is closed under both kinds of dependency, an early source never needs a value produced by a later column, so
    // every column's update list is ordered [early sources ascending, the others ascending].
    std::vector<unsigned char> early;        // n
    std::vector<i32> elevel_ptr;             // nelevels+1 -> ecols
    std::vector<i32> ecols;                  // early columns grouped by level, each level sorted by column length
    std::vector<i32> ne_cols;                // the other columns in pivot order; waves are ranges of POSITIONS in this list
    std::vector<i32> ne_pos;                 // n: position of a column in ne_cols (-1 for early columns)
    std::vector<i32> wave_rows;              // nwaves: shared-memory rows of the wave
    std::vector<i32> wrun_ptr;               // nwaves+1 -> runs of consecutive columns = contiguous slots: one bulk store each
    std::vector<i32> wrun_slot, wrun_row, wrun_cnt;     // first slot, first shared-memory row, rows
    std::vector<i32> wave_col0;       // nwaves+1 (positions in ne_cols)
    std::vector<i32> wave_hasdep;     // nwaves: 1 when some column of the wave depends on another column of it
    std::vector<i32> col_roff;        // n: first shared-memory row of the column inside its wave
    std::vector<i64> upd_split;       // n
    std::vector<i32> upd_src;         // per update: source column j
    // per wave: the union of the earlier-wave source columns its columns need, cut into segments (source column,
    // row range) and packed into batches of at most KLU_CHUNK_ROWS L entries; the CTA stages each batch in shared
    // memory once (cp.async, KLU_STAGES deep) and every column of the wave consumes it from there
    std::vector<i64> wbatch_ptr;      // nwaves+1 -> batches
    std::vector<i64> bseg_ptr;        // nbatches+1 -> segments
    std::vector<i32> seg_src, seg_off, seg_cnt, seg_row;   // source column, first L row, rows, first row in the stage
    std::vector<i32> batch_rowslot;   // nbatches * KLU_CHUNK_ROWS: global slot staged into each row (-1 = unused)
    // staged PIECES: the unit of update.  A source block = up to KLU_SN_MAX consecutive columns j0..j0+g-1 of one supernode of
    // L (nested patterns); a user column of the wave takes a suffix s0..g-1 of it (fill closure).  The block's common rows
    // below it (R of them) are cut into pieces of `nrows` rows that fit a batch: stage rows [r0 + a*nrows, +nrows) hold
    // L(rows i0.., j0+a); the first piece also stages the strictly lower triangle L(j0+a, j0+b) at tri0 (column b's g-1-b
    // entries consecutively).  seg_* above are the TMA copy descriptors that realise this layout.
    std::vector<i64> bpiece_ptr = {0};       // nbatches+1 -> pieces
    std::vector<i32> pc_j0, pc_g, pc_i0, pc_nrows, pc_r0, pc_tri0;
    std::vector<i64> pc_user_ptr;            // npieces+1 -> users
    std::vector<i32> pc_user_col, pc_user_s0;  // column of the wave (0-based), first source of the block it uses
    std::vector<i64> pc_user_upd;            // its update index of source j0+s0
    // per batch: KLU_WAVE_WARPS records of KLU_REC_U32 words -- {nseg, then per matched segment (first stage row) |
    // (rows << 8) | (row of u_jk << 16)}, the destination row (uint16) of every staged row -- followed by the row -> slot
    // table (KLU_CHUNK_ROWS words) of the batch issued while this one is consumed
    std::vector<uint32_t> bentry;
    std::vector<i32> wave_rowsrc;     // nwaves * KLU_WAVE_ROWS: value-array index gathered into each shared-memory row
                                      // (-1 = fill-in slot, zero)
    // per wave: blob with the updates whose source column is inside the wave (16-byte units):
    // ints [2q],[2q+1] = first update / count of warp q; updates {source warp, row of u_jk, rows, dest offset};
    // then the destination rows as uint16
    std::vector<i64> wblob_ptr;       // nwaves+1, in 16-byte units
    std::vector<uint32_t> wblob;
    i32 max_col_len = 0;
    bool wave_ok = true;              // false: pattern outside the wave kernel's budget (level-schedule kernel instead)
    bool have_refactor = false;       // the update lists and kernel schedules below the slot layout have been built
    // dense trailing block ("spine"): the last spine_nd columns are nearly dense after fill (76 % of the work of a
    // power-flow Jacobian).  The wave kernel applies to them only the updates from columns < spine0; the block itself
    // is then factored per matrix by a dense tensor-core LU (k_klu_dense_lu).  spine_nd = 0: disabled.
    i32 spine0 = 0, spine_nd = 0;
    std::vector<i64> upd_end;         // n: end of the updates the wave kernel applies to column k
    std::vector<i32> dense_meta;            // per block column KLU_DENSE_META ints: index of its first block entry, row bitmap
    std::vector<i32> dense_slot;            // value slot of every block entry (column by column, rows ascending)
    // solve schedule uses Lp/Li/Up/Ui/Fp/Fi of the numeric object with slots: L(i,k) at lslot, etc.
    std::vector<i32> lslot0;          // per column: first L slot (below diagonal)
    std::vector<i32> fslot0;          // per column: first F slot
};
constexpr int KLU_EARLY_MINW = 128;   // a level is "wide" (early) when it has at least this many columns: its launch is then memory-bound, not latency-bound ...
constexpr int KLU_EARLY_MAXLEN = 16;  // ... and a column is early only up to this length (shared-memory scratch rows of one warp of k_klu_early)
constexpr int KLU_WAVE_WARPS = 16;   // columns (warps) per wave
constexpr int KLU_WAVE_ROWS = 528;   // shared-memory rows for the columns of a wave (x 32 matrices x 8 B = 132 KiB)
constexpr int KLU_META_INT4 = 72;    // per batch: header {nseg} + up to 64 segment descriptors, padded to 1152 B
constexpr int KLU_CHUNK_ROWS = 128;   // L entries staged per batch (32 KiB per stage)
constexpr int KLU_STAGES = 2;
constexpr uint32_t KLU_SKIP = 0xffffffffu;
constexpr int KLU_MAXSEG = 15;        // matched segments per (batch, column); the host closes a batch before it overflows
constexpr int KLU_SN_MAX = 4;         // columns per source block (register-blocked update: one read-modify-write of a destination row per block)
constexpr int KLU_REC_HDR = 32;       // {npieces, 15 x (w0, w1)} padded
constexpr int KLU_REC_U32 = KLU_REC_HDR + KLU_CHUNK_ROWS / 2;   // per (batch, column): header + 128 uint16 destination rows
constexpr int KLU_BLOB_BYTES = 8192;  // cap of the in-wave update blob
constexpr long long KLU_WAVE_MAX_STAGED = 8ll << 20;   // staged rows over all waves (x ~52 B of tables): 8 M rows ~ 440 MB
// largest dense trailing block: 112 x 116 doubles of shared memory = two CTAs of k_klu_dense_lu per SM (sweep on the B200 with
// 4096 x ACTIVSg2000, wave kernel + dense block per batch: 160 columns, one CTA per SM 5.6 + 1.7 ms; 112: 6.1 + 1.05 ms;
// 96: 6.3 + 0.84 ms; 80: 6.7 + 0.64 ms)
constexpr int KLU_DENSE_MAX = 112, KLU_DENSE_META = 1 + (KLU_DENSE_MAX + 31) / 32;
// refactor_tables = false: only the slot layout (what loading host values, extraction and the solves need) and the level
// count; the update lists and the schedules of the refactorization kernels are added by a second call with true (the slot
// layout is a pure function of (S, N): both calls agree).
void klu_build_plan(const KluSymbolic& S, const KluNumeric& N, KluPlan& plan, bool refactor_tables = true);
// Host interpreter of the WAVE schedule (the tables k_klu_refactor_wave executes: gather, staged pieces, in-wave updates,
// dense trailing block) for one matrix.  Verification of the host-built plan in CPU tests; not a factorization path of the
// product.  Ax: the caller's values; out: LU[nslots] (U above diagonal, pivot, L below, then F), Rs[n].  Returns 0 or ST_SINGULAR.
int klu_plan_emulate(const KluSymbolic& S, const KluPlan& plan, const double* Ax, double* LU, double* Rs);

}  // namespace b200s
