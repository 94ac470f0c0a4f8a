// KLU replacement, host part (reference call sites: src/C/klu.c:136,264 klu_l_analyze; :142,337
// klu_l_factor).  SuiteSparse KLU is not in the reference tree; the published algorithm is restated
// (Davis & Palamadai Natarajan, "Algorithm 907: KLU", ACM TOMS 2010): permutation to block upper
// triangular form (maximum transversal + strongly connected components), AMD on A+A' inside every
// block, then per block a left-looking Gilbert-Peierls LU with threshold partial pivoting that
// prefers the diagonal (tol = 1e-3) on the row-scaled matrix (scale = 2: max |row|).
// The pivoting factorization runs here once per pattern; its pattern and pivot sequence are frozen
// into a KluPlan that the batched CUDA refactorization (klu_gpu.cu) replays for every matrix.
#include "klu_host.hpp"
#include "gpu.hpp"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace b200s {

namespace {

// maximum transversal: match[i] = column matched to row i (or -1)
std::vector<i32> max_transversal(i32 n, const i64* Ap, const i32* Ai, i32& nmatch) {
    std::vector<i32> match(n, -1), cheap(n), visited(n, -1), jstack, pstack_col;
    std::vector<i64> pos(n);
    for (i32 j = 0; j < n; j++) cheap[j] = 0;
    nmatch = 0;
    std::vector<i32> colstack(n), rowstack(n);
    std::vector<i64> posstack(n);
    for (i32 j0 = 0; j0 < n; j0++) {
        // iterative DFS for an augmenting path starting at column j0
        i32 head = 0;
        colstack[0] = j0;
        posstack[0] = -1;          // -1: cheap phase not done yet
        bool found = false;
        while (head >= 0) {
            const i32 j = colstack[head];
            if (posstack[head] == -1) {
                // cheap assignment: first unmatched row of column j
                bool got = false;
                for (i64& p = pos[j] = Ap[j] + cheap[j]; p < Ap[j + 1]; p++) {
                    i32 i = Ai[p];
                    if (match[i] == -1) { cheap[j] = (i32)(p - Ap[j]) + 1; rowstack[head] = i; got = true; break; }
                }
                if (got) { found = true; break; }
                cheap[j] = (i32)(Ap[j + 1] - Ap[j]);
                visited[j] = j0;
                posstack[head] = Ap[j];
            }
            bool descended = false;
            for (i64 p = posstack[head]; p < Ap[j + 1]; p++) {
                i32 i = Ai[p];
                i32 jn = match[i];
                if (jn >= 0 && visited[jn] != j0) {
                    posstack[head] = p + 1;
                    rowstack[head] = i;
                    head++;
                    colstack[head] = jn;
                    posstack[head] = -1;
                    descended = true;
                    break;
                }
            }
            if (!descended) head--;
        }
        if (found) {
            for (i32 h = head; h >= 0; h--) match[rowstack[h]] = colstack[h];
            nmatch++;
        }
    }
    return match;
}

}  // namespace

void klu_analyze(i64 n64, const i64* Ap, const i64* Ai64, KluSymbolic& S) {
    if (n64 < 0 || n64 > 0x7fffffff - 16) throw std::invalid_argument("matrix order out of range");
    const i32 n = (i32)n64;
    S = KluSymbolic();
    S.n = n;
    if (n == 0) { S.R.assign(1, 0); S.Ap.assign(1, 0); return; }
    if (Ap[0] != 0) throw std::invalid_argument("colptr[0] must be 0");
    for (i32 j = 0; j < n; j++) {
        if (Ap[j + 1] < Ap[j]) throw std::invalid_argument("colptr must be nondecreasing");
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            if (Ai64[p] < 0 || Ai64[p] >= n) throw std::invalid_argument("row index out of range");
            if (p > Ap[j] && Ai64[p] <= Ai64[p - 1]) throw std::invalid_argument("row indices must be strictly increasing within a column");
        }
    }
    S.nnz = Ap[n];
    if (S.nnz > 0x7fffffff - 16) throw std::invalid_argument("too many entries");
    S.Ap.assign(Ap, Ap + n + 1);
    S.Ai.resize(S.nnz);
    for (i64 p = 0; p < S.nnz; p++) S.Ai[p] = (i32)Ai64[p];
    const i32* Ai = S.Ai.data();

    const bool tdbg = getenv("B200S_DEBUG") != nullptr;
    auto tlast = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!tdbg) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200s klu analyze] %-26s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tlast).count());
        tlast = now;
    };
    lap("checks + copy");
    i32 nmatch = 0;
    std::vector<i32> match = max_transversal(n, Ap, Ai, nmatch);
    lap("maximum transversal");
    S.structural_rank = nmatch;
    if (nmatch < n) {
        // structurally singular: complete the matching arbitrarily so that the permutations stay valid;
        // the numeric factorization will report the zero pivot.
        std::vector<char> used(n, 0);
        for (i32 i = 0; i < n; i++) if (match[i] >= 0) used[match[i]] = 1;
        i32 j = 0;
        for (i32 i = 0; i < n; i++)
            if (match[i] < 0) { while (used[j]) j++; match[i] = j; used[j] = 1; }
    }
    // C(:,i) = A(:,match[i]) has a zero-free diagonal.  Tarjan SCC on the graph i -> r for C(r,i) != 0.
    std::vector<i32> index(n, -1), low(n), onstack(n, 0), stack, blockof(n, -1), order;
    std::vector<i32> cs(n);
    std::vector<i64> ps(n);
    order.reserve(n);
    std::vector<i32> bstart;
    i32 counter = 0, nb = 0;
    for (i32 root = 0; root < n; root++) {
        if (index[root] != -1) continue;
        i32 head = 0;
        cs[0] = root; ps[0] = Ap[match[root]];
        index[root] = low[root] = counter++;
        stack.push_back(root); onstack[root] = 1;
        while (head >= 0) {
            const i32 v = cs[head];
            const i32 col = match[v];
            bool descended = false;
            for (i64& p = ps[head]; p < Ap[col + 1]; p++) {
                const i32 w = Ai[p];
                if (index[w] == -1) {
                    p++;
                    head++;
                    cs[head] = w; ps[head] = Ap[match[w]];
                    index[w] = low[w] = counter++;
                    stack.push_back(w); onstack[w] = 1;
                    descended = true;
                    break;
                } else if (onstack[w]) low[v] = std::min(low[v], index[w]);
            }
            if (descended) continue;
            if (low[v] == index[v]) {
                bstart.push_back((i32)order.size());
                while (true) {
                    i32 w = stack.back(); stack.pop_back(); onstack[w] = 0;
                    blockof[w] = nb;
                    order.push_back(w);
                    if (w == v) break;
                }
                nb++;
            }
            head--;
            if (head >= 0) low[cs[head]] = std::min(low[cs[head]], low[v]);
        }
    }
    bstart.push_back(n);
    S.nblocks = nb;
    S.R.assign(bstart.begin(), bstart.end());
    lap("strongly connected comp.");
    // per-block AMD on B + B' (B = diagonal block of C), applied symmetrically
    std::vector<i32> local(n, -1);
    for (i32 b = 0; b < nb; b++) {
        const i32 k0 = S.R[b], nk = S.R[b + 1] - k0;
        S.maxblock = std::max(S.maxblock, nk);
        if (nk <= 2) continue;
        for (i32 t = 0; t < nk; t++) local[order[k0 + t]] = t;
        SymPattern G;
        G.n = nk;
        // pattern of B + B' without the diagonal, as CSR built in two counting passes (duplicates removed per row after a sort
        // of the short rows; a vector per row cost more than AMD itself on ACTIVSg2000)
        std::vector<i64> rp((size_t)nk + 1, 0);
        for (i32 t = 0; t < nk; t++) {
            const i32 v = order[k0 + t], col = match[v];
            for (i64 p = Ap[col]; p < Ap[col + 1]; p++) {
                const i32 w = Ai[p];
                if (w == v || blockof[w] != b) continue;
                rp[t + 1]++; rp[local[w] + 1]++;
            }
        }
        for (i32 t = 0; t < nk; t++) rp[t + 1] += rp[t];
        std::vector<i32> flat((size_t)rp[nk]);
        {
            std::vector<i64> cur(rp.begin(), rp.end() - 1);
            for (i32 t = 0; t < nk; t++) {
                const i32 v = order[k0 + t], col = match[v];
                for (i64 p = Ap[col]; p < Ap[col + 1]; p++) {
                    const i32 w = Ai[p];
                    if (w == v || blockof[w] != b) continue;
                    flat[cur[t]++] = local[w];
                    flat[cur[local[w]]++] = t;
                }
            }
        }
        G.ptr.assign(nk + 1, 0);
        G.idx.reserve(flat.size());
        for (i32 t = 0; t < nk; t++) {
            i32* r0 = flat.data() + rp[t];
            i32* r1 = flat.data() + rp[t + 1];
            std::sort(r0, r1);
            r1 = std::unique(r0, r1);
            G.idx.insert(G.idx.end(), r0, r1);
            G.ptr[t + 1] = (i64)G.idx.size();
        }
        if (nk > n / 2) lap("block graph");
        std::vector<i32> pl = amd_order(G);
        if (nk > n / 2) lap("AMD of the block");
        std::vector<i32> neworder(nk);
        for (i32 t = 0; t < nk; t++) neworder[t] = order[k0 + pl[t]];
        std::copy(neworder.begin(), neworder.end(), order.begin() + k0);
    }
    S.P = order;
    S.Q.resize(n);
    for (i32 k = 0; k < n; k++) S.Q[k] = match[order[k]];
}

template <class T>
static int klu_factor_t(const KluSymbolic& S, const T* Ax, KluNumericT<T>& N) {
    const i32 n = S.n;
    N = KluNumericT<T>();
    N.n = n;
    N.Lp.assign(n + 1, 0); N.Up.assign(n + 1, 0); N.Fp.assign(n + 1, 0);
    N.Pnum.assign(n, -1); N.Rs.assign(n, 1.0);
    if (n == 0) return ST_OK;
    const i64* Ap = S.Ap.data();
    const i32* Ai = S.Ai.data();
    const double tol = 1e-3;
    // row scaling by max |row| (scale = 2); a zero row is left unscaled and will produce a zero pivot
    std::vector<double> rs(n, 0.0);
    for (i32 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) rs[Ai[p]] = std::max(rs[Ai[p]], (double)std::abs(Ax[p]));
    for (i32 i = 0; i < n; i++) if (!(rs[i] > 0.0)) rs[i] = 1.0;
    std::vector<i32> pinvP(n);                    // row (original) -> position after the symbolic permutation
    for (i32 k = 0; k < n; k++) pinvP[S.P[k]] = k;
    std::vector<i32> pivpos(n, -1);               // symbolic row position -> final pivotal position
    std::vector<i32> rowat(n, -1);                // final pivotal position -> symbolic row position
    std::vector<T> x(n, T(0));
    std::vector<i32> mark(n, -1), reach, dstack, lcol_of(n, -1);
    std::vector<i64> pstack;
    reach.reserve(n); dstack.reserve(n); pstack.reserve(n);
    // L is built with symbolic row positions and remapped to pivotal positions at the end
    std::vector<i32> Li_tmp; std::vector<T> Lx_tmp;
    std::vector<i32> xi;
    // Symmetric pruning (Eisenstat-Liu, as in KLU's kernel): once some later column k has U(j,k) != 0 and its pivot row in
    // L(:,j), the rows of L(:,j) that were not pivotal then are all in L(:,k), so the depth-first search only needs the
    // pivotal head [Lp[j]+1, lpend[j]) of column j.  The reach (as a set) is unchanged; the search drops from ~3x the
    // cost of the numeric updates to a fraction of it on structurally symmetric patterns (power-flow Jacobians).
    std::vector<i64> lpend(n, 0);
    std::vector<char> pruned(n, 0);
    double fl = 0;
    int status = ST_OK;
    for (i32 b = 0; b < S.nblocks; b++) {
        const i32 k0 = S.R[b], k1 = S.R[b + 1];
        for (i32 k = k0; k < k1; k++) {
            const i32 col = S.Q[k];
            // scatter the scaled column: rows above the block go to F, rows inside the block into x
            xi.clear();
            reach.clear();
            for (i64 p = Ap[col]; p < Ap[col + 1]; p++) {
                const i32 r = pinvP[Ai[p]];
                const T v = Ax[p] / rs[Ai[p]];
                if (r < k0) { N.Fi.push_back(r); N.Fx.push_back(v); }        // r is remapped to pivotal later
                else if (r >= k1) throw std::logic_error("klu: entry below the block diagonal");
                else { x[r] = v; if (mark[r] != k) { mark[r] = k; xi.push_back(r); } }
            }
            // reach of the column pattern in the graph of L (DFS over already-pivotal rows)
            for (size_t q = 0, q_end = xi.size(); q < q_end; q++) {
                const i32 r0 = xi[q];
                if (pivpos[r0] < 0 || lcol_of[r0] == k) continue;
                dstack.clear(); pstack.clear();
                dstack.push_back(r0); pstack.push_back(N.Lp[pivpos[r0]] + 1);
                lcol_of[r0] = k;
                while (!dstack.empty()) {
                    const i32 r = dstack.back();
                    const i32 jc = pivpos[r];
                    bool desc = false;
                    i64& pp = pstack.back();
                    const i64 pend = lpend[jc];
                    for (; pp < pend; pp++) {
                        const i32 rr = Li_tmp[pp];
                        if (mark[rr] != k) { mark[rr] = k; xi.push_back(rr); x[rr] = T(0); }
                        if (pivpos[rr] >= 0 && lcol_of[rr] != k) {
                            lcol_of[rr] = k;
                            pp++;
                            dstack.push_back(rr); pstack.push_back(N.Lp[pivpos[rr]] + 1);
                            desc = true;
                            break;
                        }
                    }
                    if (!desc) { reach.push_back(r); dstack.pop_back(); pstack.pop_back(); }
                }
            }
            // numeric sparse triangular solve in topological order (reverse of the DFS finish order)
            for (i32 t = (i32)reach.size() - 1; t >= 0; t--) {
                const i32 r = reach[t], jc = pivpos[r];
                const T xj = x[r];
                for (i64 pp = N.Lp[jc] + 1; pp < N.Lp[jc + 1]; pp++) x[Li_tmp[pp]] -= Lx_tmp[pp] * xj;
                fl += 2.0 * (double)(N.Lp[jc + 1] - N.Lp[jc] - 1);
            }
            // pivot search among the non-pivotal rows; prefer the diagonal (symbolic row position k)
            double amax = -1.0; i32 prow = -1;
            for (i32 r : xi)
                if (pivpos[r] < 0) { double a = (double)std::abs(x[r]); if (a > amax) { amax = a; prow = r; } }
            if (pivpos[k] < 0 && mark[k] == k && (double)std::abs(x[k]) >= tol * amax && x[k] != T(0)) prow = k;
            if (prow < 0 || !(amax > 0.0) || x[prow] == T(0)) {
                // numerically (or structurally) singular column: KLU with halt_if_singular stops here
                N.singular_col = k;
                status = ST_SINGULAR;
                for (i32 r : xi) x[r] = T(0);
                return status;
            }
            const T piv = x[prow];
            pivpos[prow] = k;
            rowat[k] = prow;
            // U(:,k): pivotal rows (stored with pivotal positions), diagonal last
            for (i32 r : xi)
                if (pivpos[r] >= 0 && r != prow) { N.Ui.push_back(pivpos[r]); N.Ux.push_back(x[r]); }
            N.Ui.push_back(k); N.Ux.push_back(piv);
            N.Up[k + 1] = (i64)N.Ui.size();
            // L(:,k): unit diagonal first, then non-pivotal rows divided by the pivot
            Li_tmp.push_back(prow); Lx_tmp.push_back(T(1));
            for (i32 r : xi)
                if (pivpos[r] < 0) { Li_tmp.push_back(r); Lx_tmp.push_back(x[r] / piv); }
            N.Lp[k + 1] = (i64)Li_tmp.size();
            lpend[k] = N.Lp[k + 1];
            N.Fp[k + 1] = (i64)N.Fi.size();
            // prune the columns j with U(j,k) != 0 whose L(:,j) holds this column's pivot row
            for (i32 r : xi) {
                const i32 j = pivpos[r];
                if (j < 0 || r == prow || pruned[j]) continue;
                const i64 b0 = N.Lp[j] + 1, b1 = N.Lp[j + 1];
                bool has = false;
                for (i64 pp = b0; pp < b1; pp++) if (Li_tmp[pp] == prow) { has = true; break; }
                if (!has) continue;
                i64 head = b0, tail = b1;
                while (head < tail) {
                    if (pivpos[Li_tmp[head]] >= 0) head++;
                    else { tail--; std::swap(Li_tmp[head], Li_tmp[tail]); std::swap(Lx_tmp[head], Lx_tmp[tail]); }
                }
                lpend[j] = head;
                pruned[j] = 1;
            }
            for (i32 r : xi) x[r] = T(0);
        }
    }
    // remap rows to pivotal positions, sort columns by row
    N.Li.resize(Li_tmp.size()); N.Lx = Lx_tmp;
    for (size_t p = 0; p < Li_tmp.size(); p++) N.Li[p] = pivpos[Li_tmp[p]];
    for (size_t p = 0; p < N.Fi.size(); p++) N.Fi[p] = pivpos[N.Fi[p]];
    // rows ascending inside every column by a double transposition (bucket by row, then read the rows in order):
    // O(nnz + n) instead of one comparison sort per column
    auto sort_cols = [&](std::vector<i64>& Cp, std::vector<i32>& Ci, std::vector<T>& Cx) {
        const i64 nz = Cp[n];
        if (nz == 0) return;
        std::vector<i64> rp(n + 1, 0), cpos(Cp.begin(), Cp.end() - 1);
        for (i64 p = 0; p < nz; p++) rp[Ci[p] + 1]++;
        for (i32 i = 0; i < n; i++) rp[i + 1] += rp[i];
        std::vector<i32> rc(nz); std::vector<T> rx(nz);
        for (i32 k = 0; k < n; k++)
            for (i64 p = Cp[k]; p < Cp[k + 1]; p++) { const i64 q = rp[Ci[p]]++; rc[q] = k; rx[q] = Cx[p]; }
        // rp[i] is now the end of row i
        i64 q = 0;
        for (i32 i = 0; i < n; i++)
            for (; q < rp[i]; q++) { const i64 d = cpos[rc[q]]++; Ci[d] = i; Cx[d] = rx[q]; }
    };
    sort_cols(N.Lp, N.Li, N.Lx);
    sort_cols(N.Up, N.Ui, N.Ux);
    sort_cols(N.Fp, N.Fi, N.Fx);
    for (i32 k = 0; k < n; k++) { N.Pnum[k] = S.P[rowat[k]]; N.Rs[k] = rs[N.Pnum[k]]; }
    N.flops = fl;
    return status;
}

int klu_factor(const KluSymbolic& S, const double* Ax, KluNumeric& N) { return klu_factor_t<double>(S, Ax, N); }
int klu_factor_z(const KluSymbolic& S, const std::complex<double>* Ax, KluNumericZ& N) { return klu_factor_t<std::complex<double>>(S, Ax, N); }

void klu_build_plan(const KluSymbolic& S, const KluNumeric& N, KluPlan& P, bool refactor_tables) {
    const i32 n = S.n;
    const bool tdbg = getenv("B200S_DEBUG") != nullptr;
    auto tlast = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!tdbg) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200s klu plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tlast).count());
        tlast = now;
    };
    P = KluPlan();
    P.n = n;
    P.nnzA = S.nnz;
    P.cbeg.assign(n + 1, 0);
    P.udiag_slot.assign(n, 0);
    P.lslot0.assign(n, 0);
    P.fslot0.assign(n, 0);
    // slots: per column [U above diag..., U diag, L below diag...]
    i64 s = 0;
    std::vector<i64> uslot0(n);
    for (i32 k = 0; k < n; k++) {
        P.cbeg[k] = s;
        uslot0[k] = s;
        s += N.Up[k + 1] - N.Up[k];
        P.udiag_slot[k] = (i32)(s - 1);
        P.lslot0[k] = (i32)s;
        s += N.Lp[k + 1] - N.Lp[k] - 1;
    }
    P.cbeg[n] = s;
    P.lu_slots = s;
    for (i32 k = 0; k < n; k++) { P.fslot0[k] = (i32)s; s += N.Fp[k + 1] - N.Fp[k]; }
    P.nslots = s;
    if (s > 0x7fffffff - 16) throw std::invalid_argument("klu plan too large");
    P.slot_src.assign(s, -1);
    P.slot_row.assign(s, 0);
    // row position (pivotal) -> slot inside the current column
    std::vector<i32> slot_of_row(n, -1), pinvnum(n);
    for (i32 k = 0; k < n; k++) pinvnum[N.Pnum[k]] = k;
    P.have_refactor = refactor_tables;
    if (refactor_tables) {
        // one destination per multiply-add of the factorization, one update per off-diagonal U entry
        const size_t nupd = (size_t)(N.Up[n] - n);
        P.dest.reserve((size_t)(N.flops / 2) + 16);
        P.upd_src.reserve(nupd); P.upd_uslot.reserve(nupd); P.upd_lslot.reserve(nupd); P.upd_cnt.reserve(nupd); P.upd_dest.reserve(nupd);
        P.upd_ptr.reserve((size_t)n + 1);
    }
    for (i32 k = 0; k < n; k++) {
        for (i64 p = N.Up[k]; p < N.Up[k + 1]; p++) { i32 sl = (i32)(uslot0[k] + (p - N.Up[k])); P.slot_row[sl] = N.Ui[p]; slot_of_row[N.Ui[p]] = sl; }
        for (i64 p = N.Lp[k] + 1; p < N.Lp[k + 1]; p++) { i32 sl = (i32)(P.lslot0[k] + (p - N.Lp[k] - 1)); P.slot_row[sl] = N.Li[p]; slot_of_row[N.Li[p]] = sl; }
        // F slots are looked up separately (rows above the block)
        const i32 col = S.Q[k];
        for (i64 p = S.Ap[col]; p < S.Ap[col + 1]; p++) {
            const i32 r = pinvnum[S.Ai[p]];
            // F entry?
            const i32* fb = N.Fi.data() + N.Fp[k];
            const i32* fe = N.Fi.data() + N.Fp[k + 1];
            const i32* it = std::lower_bound(fb, fe, r);
            if (it != fe && *it == r) {
                i32 sl = P.fslot0[k] + (i32)(it - fb);
                P.slot_src[sl] = (i32)p; P.slot_row[sl] = r;
            } else {
                i32 sl = slot_of_row[r];
                if (sl < P.cbeg[k] || sl >= P.cbeg[k + 1] || P.slot_row[sl] != r) throw std::logic_error("klu plan: entry outside the LU pattern");
                P.slot_src[sl] = (i32)p;
            }
        }
        // updates of column k: for each U entry (j,k), j < k ascending, L(:,j) below the diagonal
        P.upd_ptr.push_back((i64)P.upd_uslot.size());
        if (refactor_tables)
        for (i64 p = N.Up[k]; p < N.Up[k + 1] - 1; p++) {
            const i32 j = N.Ui[p];
            const i64 cnt = N.Lp[j + 1] - N.Lp[j] - 1;
            if (cnt == 0) continue;
            P.upd_src.push_back(j);
            P.upd_uslot.push_back((i32)(uslot0[k] + (p - N.Up[k])));
            P.upd_lslot.push_back(P.lslot0[j]);
            P.upd_cnt.push_back((i32)cnt);
            P.upd_dest.push_back((i64)P.dest.size());
            for (i64 q = N.Lp[j] + 1; q < N.Lp[j + 1]; q++) {
                const i32 sl = slot_of_row[N.Li[q]];
                if (sl < P.cbeg[k] || sl >= P.cbeg[k + 1] || P.slot_row[sl] != N.Li[q]) throw std::logic_error("klu plan: fill entry missing");
                P.dest.push_back(sl);
            }
        }
    }
    P.upd_ptr.push_back((i64)P.upd_uslot.size());
    // F entries' scaling rows were set above; entries of F with no source cannot exist
    // row lists for the scaling pass
    P.rowptr.assign(n + 1, 0);
    for (i64 p = 0; p < S.nnz; p++) P.rowptr[pinvnum[S.Ai[p]] + 1]++;
    for (i32 i = 0; i < n; i++) P.rowptr[i + 1] += P.rowptr[i];
    P.rowent.resize(S.nnz);
    {
        std::vector<i64> pos(P.rowptr.begin(), P.rowptr.end() - 1);
        for (i64 p = 0; p < S.nnz; p++) P.rowent[pos[pinvnum[S.Ai[p]]]++] = (i32)p;
    }
    lap("slots + update lists");
    {
        // number of levels of the column dependency graph (column k depends on every column j with U(j,k) != 0)
        std::vector<i32> lvq(n, 0);
        P.nlevels = 0;
        for (i32 k = 0; k < n; k++) {
            i32 l = 0;
            for (i64 p = N.Up[k]; p < N.Up[k + 1] - 1; p++) l = std::max(l, lvq[N.Ui[p]] + 1);
            lvq[k] = l;
            P.nlevels = std::max(P.nlevels, l + 1);
        }
    }
    if (!refactor_tables) return;
    // ---- dense trailing block: the largest nd <= KLU_DENSE_MAX (multiple of 16) inside the last BTF block whose
    // L+U pattern restricted to the last nd rows/columns is at least 30 % dense
    P.upd_end.assign(n, 0);
    for (i32 k = 0; k < n; k++) P.upd_end[k] = P.upd_ptr[k + 1];
    P.spine0 = n; P.spine_nd = 0;
    {
        const i32 lastblk = S.nblocks > 0 ? S.R[S.nblocks] - S.R[S.nblocks - 1] : 0;
        const i32 dmax = getenv("B200S_KLU_DENSE_MAX") ? std::min<i32>(KLU_DENSE_MAX, atoi(getenv("B200S_KLU_DENSE_MAX"))) : KLU_DENSE_MAX;
        for (i32 nd = std::min<i32>(dmax, (lastblk / 16) * 16); nd >= 48; nd -= 16) {
            const i32 s0 = n - nd;
            i64 cnt = 0;
            for (i32 k = s0; k < n; k++)
                for (i64 sl = P.cbeg[k]; sl < P.cbeg[k + 1]; sl++) if (P.slot_row[sl] >= s0) cnt++;
            if ((double)cnt >= 0.30 * nd * nd) { P.spine0 = s0; P.spine_nd = nd; break; }
        }
        if (P.spine_nd > 0) {
            const i32 s0 = P.spine0;
            P.dense_meta.assign((size_t)P.spine_nd * KLU_DENSE_META, 0);
            for (i32 k = s0; k < n; k++) {
                i64 u = P.upd_ptr[k];
                while (u < P.upd_ptr[k + 1] && P.upd_src[u] < s0) u++;
                P.upd_end[k] = u;
                // slots of a column ascend by row: the block entries are its tail
                i32* m = P.dense_meta.data() + (size_t)(k - s0) * KLU_DENSE_META;
                i64 sl = P.cbeg[k];
                while (sl < P.cbeg[k + 1] && P.slot_row[sl] < s0) sl++;
                m[0] = (i32)P.dense_slot.size();
                const i64 first = sl;
                for (; sl < P.cbeg[k + 1]; sl++) {
                    const i32 r = P.slot_row[sl] - s0;
                    P.dense_slot.push_back((i32)sl);
                    if (r < 0 || (sl > first && P.slot_row[sl] <= P.slot_row[sl - 1])) throw std::logic_error("klu plan: column slots not ascending by row");
                    m[1 + (r >> 5)] |= (i32)(1u << (r & 31));
                }
            }
        }
    }
    lap("dense block");
    // ---- early columns: the wide first levels of the dependency graph (see KluPlan::early)
    P.early.assign(n, 0);
    P.elevel_ptr.assign(1, 0);
    P.ecols.clear();
    {
        const char* ev = getenv("B200S_KLU_EARLY");
        const bool enabled = !(ev && atoi(ev) == 0);
        // level over row AND column dependencies: columns ascending, level[k] is final once the U part of column k has been
        // looked at (the L rows of earlier columns were pushed before)
        std::vector<i32> lv(n, 0);
        for (i32 k = 0; k < n; k++) {
            i32 l = lv[k];
            for (i64 q = N.Up[k]; q < N.Up[k + 1] - 1; q++) l = std::max(l, lv[N.Ui[q]] + 1);
            lv[k] = l;
            for (i64 q = N.Lp[k] + 1; q < N.Lp[k + 1]; q++) lv[N.Li[q]] = std::max(lv[N.Li[q]], l + 1);
        }
        i32 nl = 0;
        for (i32 k = 0; k < n; k++) nl = std::max(nl, lv[k] + 1);
        std::vector<i32> cnt(nl, 0);
        for (i32 k = 0; k < n; k++) {
            if (k >= P.spine0 || P.cbeg[k + 1] - P.cbeg[k] > KLU_EARLY_MAXLEN) continue;       // long columns and the dense block's are never early
            cnt[lv[k]]++;
        }
        i32 lmax = -1, total = 0;
        const i32 minw = getenv("B200S_KLU_EARLY_MINW") ? atoi(getenv("B200S_KLU_EARLY_MINW")) : KLU_EARLY_MINW;
        const i32 maxlev = getenv("B200S_KLU_EARLY_LEVELS") ? atoi(getenv("B200S_KLU_EARLY_LEVELS")) : 0x7fffffff;
        while (enabled && lmax + 1 < nl && lmax + 1 < maxlev && cnt[lmax + 1] >= minw) { lmax++; total += cnt[lmax]; }
        if (total < 8 * minw) lmax = -1;          // not worth the extra launches
        if (tdbg) fprintf(stderr, "[b200s klu plan] dependency levels %d, early levels %d with %d of %d columns\n", nl, lmax + 1, lmax >= 0 ? total : 0, n);
        if (lmax >= 0) {
            // early = in a wide level, short enough for the warp's scratch, and depending (by column and by row) on early
            // columns only -- the set stays closed under both kinds of dependency
            std::vector<unsigned char> rowdep_late(n, 0);       // row k receives an L entry from a late column
            for (i32 k = 0; k < n; k++) {
                bool e = lv[k] <= lmax && k < P.spine0 && P.cbeg[k + 1] - P.cbeg[k] <= KLU_EARLY_MAXLEN && !rowdep_late[k];
                for (i64 q = N.Up[k]; e && q < N.Up[k + 1] - 1; q++) e = P.early[N.Ui[q]] != 0;
                P.early[k] = e;
                if (!e) for (i64 q = N.Lp[k] + 1; q < N.Lp[k + 1]; q++) rowdep_late[N.Li[q]] = 1;
            }
            P.elevel_ptr.assign(lmax + 2, 0);
            for (i32 k = 0; k < n; k++) if (P.early[k]) P.elevel_ptr[lv[k] + 1]++;
            for (i32 l = 0; l <= lmax; l++) P.elevel_ptr[l + 1] += P.elevel_ptr[l];
            P.ecols.resize(P.elevel_ptr[lmax + 1]);
            std::vector<i32> pos(P.elevel_ptr.begin(), P.elevel_ptr.end() - 1);
            for (i32 k = 0; k < n; k++) if (P.early[k]) P.ecols[pos[lv[k]]++] = k;
            for (i32 l = 0; l <= lmax; l++) {
                i32* b = P.ecols.data() + P.elevel_ptr[l];
                i32* e = P.ecols.data() + P.elevel_ptr[l + 1];
                std::stable_sort(b, e, [&](i32 a, i32 c) { return P.cbeg[a + 1] - P.cbeg[a] < P.cbeg[c + 1] - P.cbeg[c]; });
            }
            // update lists: early sources first (both parts keep their ascending order)
            std::vector<i32> t_src, t_us, t_ls, t_cnt; std::vector<i64> t_dst;
            for (i32 k = 0; k < n; k++) {
                const i64 u0 = P.upd_ptr[k], u1 = P.upd_ptr[k + 1];
                bool mixed = false, seen_late = false;
                for (i64 u = u0; u < u1; u++) { if (!P.early[P.upd_src[u]]) seen_late = true; else if (seen_late) mixed = true; }
                if (P.early[k] && seen_late) throw std::logic_error("klu plan: early column with a late source");
                if (!mixed) continue;
                t_src.clear(); t_us.clear(); t_ls.clear(); t_cnt.clear(); t_dst.clear();
                for (int pass = 0; pass < 2; pass++)
                    for (i64 u = u0; u < u1; u++)
                        if ((P.early[P.upd_src[u]] != 0) == (pass == 0)) {
                            t_src.push_back(P.upd_src[u]); t_us.push_back(P.upd_uslot[u]); t_ls.push_back(P.upd_lslot[u]);
                            t_cnt.push_back(P.upd_cnt[u]); t_dst.push_back(P.upd_dest[u]);
                        }
                for (i64 u = u0; u < u1; u++) {
                    P.upd_src[u] = t_src[u - u0]; P.upd_uslot[u] = t_us[u - u0]; P.upd_lslot[u] = t_ls[u - u0];
                    P.upd_cnt[u] = t_cnt[u - u0]; P.upd_dest[u] = t_dst[u - u0];
                }
            }
            // the dense block's columns: the updates the wave kernel applies end where the sources inside the block begin
            for (i32 k = P.spine0; k < n; k++) {
                i64 u = P.upd_ptr[k];
                while (u < P.upd_ptr[k + 1] && P.upd_src[u] < P.spine0) u++;
                for (i64 v = u; v < P.upd_ptr[k + 1]; v++) if (P.upd_src[v] < P.spine0) throw std::logic_error("klu plan: dense-block sources are not a suffix");
                P.upd_end[k] = u;
            }
        }
    }
    P.ne_cols.clear();
    P.ne_pos.assign(n, -1);
    for (i32 k = 0; k < n; k++) if (!P.early[k]) { P.ne_pos[k] = (i32)P.ne_cols.size(); P.ne_cols.push_back(k); }
    const i32 nne = (i32)P.ne_cols.size();
    // a source is "inside the wave that starts at position p0" when it is a late column at a position >= p0
    auto in_wave = [&](i32 j, i32 p0) { return !P.early[j] && P.ne_pos[j] >= p0; };
    lap("early columns");
    // wave schedule
    P.col_roff.assign(n, 0);
    P.upd_split.assign(n, 0);
    P.wave_col0.clear();
    P.wave_rows.clear();
    P.wrun_ptr.assign(1, 0);
    P.wrun_slot.clear(); P.wrun_row.clear(); P.wrun_cnt.clear();
    P.max_col_len = 0;
    for (i32 k = 0; k < n; k++) P.max_col_len = std::max<i32>(P.max_col_len, (i32)(P.cbeg[k + 1] - P.cbeg[k]));
    {
        // in-wave blob size if the wave [k0, k1) were closed: header + updates + dest lists
        // (waves are ranges [p0, p1) of positions in ne_cols)
        auto blob_bytes = [&](i32 p0, i32 p1) {
            i64 nu = 0, nd = 0;
            for (i32 q = p0; q < p1; q++) {
                const i32 c = P.ne_cols[q];
                for (i64 u = P.upd_ptr[c]; u < P.upd_end[c]; u++)
                    if (in_wave(P.upd_src[u], p0)) { nu++; nd += P.upd_cnt[u]; }
            }
            return (i64)(2 * KLU_WAVE_WARPS) * 4 + nu * 16 + ((nd * 2 + 15) / 16) * 16;
        };
        i32 pk = 0;
        while (pk < nne) {
            P.wave_col0.push_back(pk);
            const i32 p0 = pk;
            i32 rows = 0, cnt = 0;
            while (pk < nne && cnt < KLU_WAVE_WARPS) {
                const i32 k = P.ne_cols[pk];
                const i32 len = (i32)(P.cbeg[k + 1] - P.cbeg[k]);
                if (cnt > 0 && rows + len > KLU_WAVE_ROWS) break;
                if (cnt > 0 && blob_bytes(p0, pk + 1) > KLU_BLOB_BYTES) break;
                P.col_roff[k] = rows;
                // runs of consecutive columns own consecutive slots and consecutive shared-memory rows
                if (cnt > 0 && P.ne_cols[pk - 1] == k - 1) P.wrun_cnt.back() += len;
                else { P.wrun_slot.push_back((i32)P.cbeg[k]); P.wrun_row.push_back(rows); P.wrun_cnt.push_back(len); }
                rows += len;
                cnt++;
                pk++;
            }
            P.wave_rows.push_back(rows);
            P.wrun_ptr.push_back((i32)P.wrun_slot.size());
        }
        P.wave_col0.push_back(nne);
        // Budget: every wave re-streams the earlier L columns its columns need, with ~52 bytes of records per staged row.
        // With heavy fill (columns longer than the shared-memory wave, or narrow waves over a dense factor) the tables
        // grow like the flop count; such patterns go to the level-schedule kernel instead and no wave tables are built.
        {
            i64 staged = 0;
            std::vector<i32> mark(n, -1);
            for (i32 w = 0; w + 1 < (i32)P.wave_col0.size() && staged <= KLU_WAVE_MAX_STAGED; w++) {
                const i32 p0 = P.wave_col0[w];
                for (i32 q = p0; q < P.wave_col0[w + 1]; q++) {
                    const i32 c = P.ne_cols[q];
                    for (i64 u = P.upd_ptr[c]; u < P.upd_end[c] && !in_wave(P.upd_src[u], p0); u++)
                        if (mark[P.upd_src[u]] != w) { mark[P.upd_src[u]] = w; staged += P.upd_cnt[u]; }
                }
            }
            P.wave_ok = P.max_col_len <= KLU_WAVE_ROWS && staged <= KLU_WAVE_MAX_STAGED;
            if (!P.wave_ok) {
                // level-schedule kernel for everything (the early-first order of the update lists is a valid order for it too)
                P.wave_col0.assign(1, 0);
                P.wave_rows.clear(); P.wrun_ptr.assign(1, 0); P.wrun_slot.clear(); P.wrun_row.clear(); P.wrun_cnt.clear();
                P.spine0 = n; P.spine_nd = 0;
                P.dense_meta.clear(); P.dense_slot.clear();
                for (i32 k = 0; k < n; k++) P.upd_end[k] = P.upd_ptr[k + 1];
                std::fill(P.early.begin(), P.early.end(), 0);
                P.elevel_ptr.assign(1, 0); P.ecols.clear();
            }
        }
        const i32 nw = (i32)P.wave_col0.size() - 1;
        P.wave_hasdep.assign(nw, 0);
        P.wbatch_ptr.assign(1, 0);
        P.bseg_ptr.assign(1, 0);
        // nested[j]: Lpattern(j) = {j+1} u Lpattern(j+1) -- columns j and j+1 belong to one supernode of L
        std::vector<unsigned char> nested(n, 0);
        for (i32 j = 0; j + 1 < n; j++) {
            const i64 a0 = N.Lp[j] + 1, a1 = N.Lp[j + 1], b0 = N.Lp[j + 1] + 1, b1 = N.Lp[j + 2];
            if (a1 - a0 != b1 - b0 + 1 || a1 == a0 || N.Li[a0] != j + 1) continue;
            bool same = true;
            for (i64 q = 0; q < b1 - b0 && same; q++) same = N.Li[a0 + 1 + q] == N.Li[b0 + q];
            nested[j] = same;
        }
        const i32 snmax = getenv("B200S_KLU_SN_MAX") ? std::max(1, std::min<i32>(KLU_SN_MAX, atoi(getenv("B200S_KLU_SN_MAX")))) : KLU_SN_MAX;
        std::vector<i32> srcs, matched, s0of;
        std::vector<i64> ucur;
        for (i32 w = 0; w < nw; w++) {
            const i32 k0 = P.wave_col0[w], wc = P.wave_col0[w + 1] - k0;        // positions in ne_cols
            const i32* wcol = P.ne_cols.data() + k0;                              // the wave's columns
            srcs.clear();
            for (i32 q = 0; q < wc; q++) {
                const i32 c = wcol[q];
                i64 u = P.upd_ptr[c];
                while (u < P.upd_end[c] && !in_wave(P.upd_src[u], k0)) { srcs.push_back(P.upd_src[u]); u++; }
                P.upd_split[c] = u;
                if (u < P.upd_end[c]) P.wave_hasdep[w] = 1;
            }
            // the order of every column's update list: early sources first, ascending inside both classes
            auto src_less = [&](i32 a, i32 b) { return P.early[a] != P.early[b] ? P.early[a] > P.early[b] : a < b; };
            std::sort(srcs.begin(), srcs.end(), src_less);
            srcs.erase(std::unique(srcs.begin(), srcs.end()), srcs.end());
            i32 fill = KLU_CHUNK_ROWS;      // row units used in the open batch (full => start a new one)
            matched.assign(wc, 0);                                      // pieces of the open batch used by each column
            ucur.resize(wc);
            s0of.resize(wc);
            for (i32 q = 0; q < wc; q++) ucur[q] = P.upd_ptr[wcol[q]];
            auto open_batch = [&]() {
                std::fill(matched.begin(), matched.end(), 0);
                if (!P.seg_src.empty() && P.bseg_ptr.back() != (i64)P.seg_src.size()) {
                    P.bseg_ptr.push_back((i64)P.seg_src.size());
                    P.bpiece_ptr.push_back((i64)P.pc_j0.size());
                }
                P.batch_rowslot.insert(P.batch_rowslot.end(), KLU_CHUNK_ROWS, -1);
                fill = 0;
            };
            auto add_copy = [&](i32 j, i32 off, i32 cnt, i32 row) {
                P.seg_src.push_back(j); P.seg_off.push_back(off); P.seg_cnt.push_back(cnt); P.seg_row.push_back(row);
                i32* rs = P.batch_rowslot.data() + P.batch_rowslot.size() - KLU_CHUNK_ROWS;
                for (i32 r = 0; r < cnt; r++) rs[row + r] = P.lslot0[j] + off + r;
            };
            size_t si = 0;
            while (si < srcs.size()) {
                // ---- source block: up to snmax consecutive columns of one L supernode, each user column taking a suffix of it
                const i32 j0 = srcs[si];
                i32 g = 1;
                while (g < snmax && si + g < srcs.size() && srcs[si + g] == j0 + g && nested[j0 + g - 1] && P.early[j0 + g] == P.early[j0]) g++;
                // users and their first source inside the block; the block is cut where a user would skip a column
                // (cannot happen for a structurally closed pattern, checked all the same)
                for (;;) {
                    bool ok = true;
                    for (i32 q = 0; q < wc && ok; q++) {
                        const i32 c = wcol[q];
                        i64 u = ucur[q];
                        s0of[q] = -1;
                        if (u >= P.upd_split[c] || P.early[P.upd_src[u]] != P.early[j0] || P.upd_src[u] < j0 || P.upd_src[u] >= j0 + g) continue;
                        s0of[q] = P.upd_src[u] - j0;
                        for (i32 a = s0of[q]; a < g && ok; a++, u++) ok = u < P.upd_split[c] && P.upd_src[u] == j0 + a;
                    }
                    if (ok || g == 1) break;
                    g--;
                }
                si += g;
                const i32 R = (i32)(N.Lp[j0 + g] - N.Lp[j0 + g - 1] - 1);       // rows below the block, common to its columns
                const i32 tri = g * (g - 1) / 2;
                i32 i0 = 0;
                bool first = true;
                while (first || i0 < R) {
                    bool overflow = false;
                    for (i32 q = 0; q < wc; q++) if (s0of[q] >= 0 && matched[q] >= KLU_MAXSEG) overflow = true;
                    const i32 need0 = first ? tri : 0;
                    i32 avail = KLU_CHUNK_ROWS - fill - need0;
                    i32 nrows = avail > 0 ? std::min(R - i0, avail / g) : 0;
                    // a piece that would be cut far below a full batch's worth goes to a fresh batch
                    if (overflow || fill == KLU_CHUNK_ROWS || avail < 0 || (nrows < R - i0 && nrows < std::min(R - i0, 16)) || (R == 0 && avail < 0)) {
                        open_batch();
                        avail = KLU_CHUNK_ROWS - need0;
                        nrows = std::min(R - i0, avail / g);
                    }
                    const i32 tri0 = first && tri > 0 ? fill : -1;
                    if (first && tri > 0) {
                        i32 t = fill;
                        for (i32 b = 0; b + 1 < g; b++) { add_copy(j0 + b, 0, g - 1 - b, t); t += g - 1 - b; }
                        fill += tri;
                    }
                    const i32 r0 = fill;
                    for (i32 a = 0; a < g && nrows > 0; a++) add_copy(j0 + a, (g - 1 - a) + i0, nrows, r0 + a * nrows);
                    fill += g * nrows;
                    P.pc_j0.push_back(j0); P.pc_g.push_back(g); P.pc_i0.push_back(i0); P.pc_nrows.push_back(nrows);
                    P.pc_r0.push_back(r0); P.pc_tri0.push_back(tri0);
                    P.pc_user_ptr.push_back((i64)P.pc_user_col.size());
                    for (i32 q = 0; q < wc; q++)
                        if (s0of[q] >= 0) {
                            matched[q]++;
                            P.pc_user_col.push_back(q); P.pc_user_s0.push_back(s0of[q]); P.pc_user_upd.push_back(ucur[q]);
                        }
                    i0 += nrows;
                    first = false;
                    if (R == 0) break;
                }
                for (i32 q = 0; q < wc; q++) if (s0of[q] >= 0) ucur[q] += g - s0of[q];
            }
            for (i32 q = 0; q < wc; q++) if (ucur[q] != P.upd_split[wcol[q]]) throw std::logic_error("klu plan: staged blocks do not cover a column");
            if (P.bseg_ptr.back() != (i64)P.seg_src.size()) { P.bseg_ptr.push_back((i64)P.seg_src.size()); P.bpiece_ptr.push_back((i64)P.pc_j0.size()); }
            P.wbatch_ptr.push_back((i64)P.bseg_ptr.size() - 1);
        }
        P.pc_user_ptr.push_back((i64)P.pc_user_col.size());
        if (getenv("B200S_KLU_DUMP_WAVES")) {
            // per wave: first column, columns, batches, pieces, (piece, user) pairs, sum over batches of the largest per-column
            // piece count (what the lock-step batch ring serialises), largest per-column total, multiply-adds
            for (i32 w = 0; w < nw; w++) {
                const i32 k0 = P.wave_col0[w], wc = P.wave_col0[w + 1] - k0;
                i64 npc = 0, npu = 0, lock = 0; double fma = 0;
                std::vector<i64> tot(wc, 0);
                for (i64 bi = P.wbatch_ptr[w]; bi < P.wbatch_ptr[w + 1]; bi++) {
                    std::vector<i64> m(wc, 0);
                    for (i64 pc = P.bpiece_ptr[bi]; pc < P.bpiece_ptr[bi + 1]; pc++) {
                        npc++;
                        for (i64 uq = P.pc_user_ptr[pc]; uq < P.pc_user_ptr[pc + 1]; uq++) {
                            npu++; m[P.pc_user_col[uq]]++; tot[P.pc_user_col[uq]]++;
                            fma += (double)(P.pc_g[pc] - P.pc_user_s0[uq]) * P.pc_nrows[pc];
                        }
                    }
                    lock += *std::max_element(m.begin(), m.end());
                }
                fprintf(stderr, "wave %d col0 %d wc %d rows %d batches %lld pieces %lld users %lld lockstep %lld maxcol %lld fma %.0f\n", w, P.ne_cols[k0], wc,
                        (int)P.wave_rows[w], (long long)(P.wbatch_ptr[w + 1] - P.wbatch_ptr[w]), (long long)npc, (long long)npu,
                        (long long)lock, (long long)*std::max_element(tot.begin(), tot.end()), fma);
                (void)k0;
            }
        }
        if (tdbg) {
            // multiply-adds by the number of source columns a (piece, user) pair applies at once
            double fm[KLU_SN_MAX + 1] = {0};
            i64 np_[KLU_SN_MAX + 1] = {0};
            for (size_t pc = 0; pc < P.pc_j0.size(); pc++)
                for (i64 uq = P.pc_user_ptr[pc]; uq < P.pc_user_ptr[pc + 1]; uq++) {
                    const i32 gu = P.pc_g[pc] - P.pc_user_s0[uq];
                    fm[gu] += (double)gu * P.pc_nrows[pc]; np_[gu]++;
                }
            for (i32 q = 1; q <= KLU_SN_MAX; q++) fprintf(stderr, "[b200s klu plan] (piece, user) pairs applying %d source columns: %lld, multiply-adds %.0f\n", q, (long long)np_[q], fm[q]);
        }
    }
    lap("waves + batches");
    // ---- staged tables of the wave kernel
    {
        const i32 nw = (i32)P.wave_col0.size() - 1;
        const size_t nbatch = P.bseg_ptr.size() - 1;
        const size_t BST = (size_t)KLU_WAVE_WARPS * KLU_REC_U32 + KLU_CHUNK_ROWS;
        P.bentry.assign(nbatch * BST, 0u);
        for (i32 w = 0; w < nw; w++)
            for (i64 bi = P.wbatch_ptr[w]; bi < P.wbatch_ptr[w + 1]; bi++) {
                const i64 nx = bi + KLU_STAGES - 1;
                for (i32 r = 0; r < KLU_CHUNK_ROWS; r++)
                    P.bentry[bi * BST + (size_t)KLU_WAVE_WARPS * KLU_REC_U32 + r] =
                        (nx < P.wbatch_ptr[w + 1]) ? (uint32_t)P.batch_rowslot[nx * KLU_CHUNK_ROWS + r] : 0xffffffffu;
            }
        P.wave_rowsrc.assign((size_t)nw * KLU_WAVE_ROWS, -1);
        P.wblob_ptr.assign(1, 0);
        for (i32 w = 0; w < nw; w++) {
            const i32 k0 = P.wave_col0[w], k1 = P.wave_col0[w + 1];          // positions in ne_cols
            for (i32 pq = k0; pq < k1; pq++) {
                const i32 c = P.ne_cols[pq];
                const i64 cb = P.cbeg[c];
                const i32 len = (i32)(P.cbeg[c + 1] - cb);
                if (P.col_roff[c] + len > KLU_WAVE_ROWS) continue;      // oversized single column: fallback kernel
                for (i32 sl = 0; sl < len; sl++) P.wave_rowsrc[(size_t)w * KLU_WAVE_ROWS + P.col_roff[c] + sl] = P.slot_src[cb + sl];
            }
            // records of the staged pieces, per (batch, column of the wave): {npieces, then two words per piece:
            //   w0 = first stage row of the rectangle | rows << 8 | (row of u for the first source used, in the column) << 16
            //   w1 = block width g | first source used s0 << 4 | first stage row of the triangle << 8 | has-triangle << 16 |
            //        continuation-of-the-previous-piece's-block << 17},
            // and for every rectangle row (stage row of the block's first column) the destination row in the column
            for (i64 bi = P.wbatch_ptr[w]; bi < P.wbatch_ptr[w + 1]; bi++)
                for (i64 pc = P.bpiece_ptr[bi]; pc < P.bpiece_ptr[bi + 1]; pc++)
                    for (i64 uq = P.pc_user_ptr[pc]; uq < P.pc_user_ptr[pc + 1]; uq++) {
                        const i32 q = P.pc_user_col[uq], c = P.ne_cols[k0 + q], s0 = P.pc_user_s0[uq], g = P.pc_g[pc];
                        const i64 cb = P.cbeg[c], u0 = P.pc_user_upd[uq], ulast = u0 + (g - 1 - s0);
                        uint32_t* rec = P.bentry.data() + (size_t)bi * BST + (size_t)q * KLU_REC_U32;
                        if (rec[0] >= (uint32_t)KLU_MAXSEG) throw std::logic_error("klu plan: too many pieces in a batch record");
                        const uint32_t k = rec[0]++;
                        const i32 uloc = (i32)(P.upd_uslot[u0] - cb);
                        for (i32 t = 1; t < g - s0; t++)
                            if (P.upd_uslot[u0 + t] != P.upd_uslot[u0] + t) throw std::logic_error("klu plan: U rows of a source block are not consecutive");
                        rec[1 + 2 * k] = (uint32_t)P.pc_r0[pc] | ((uint32_t)P.pc_nrows[pc] << 8) | ((uint32_t)uloc << 16);
                        rec[2 + 2 * k] = (uint32_t)g | ((uint32_t)s0 << 4) | ((uint32_t)std::max(P.pc_tri0[pc], 0) << 8) |
                                         ((uint32_t)(P.pc_tri0[pc] >= 0) << 16) | ((uint32_t)(P.pc_i0[pc] > 0) << 17);
                        uint16_t* dd = reinterpret_cast<uint16_t*>(rec + KLU_REC_HDR);
                        if (P.upd_src[ulast] != P.pc_j0[pc] + g - 1 || P.upd_cnt[ulast] < P.pc_i0[pc] + P.pc_nrows[pc]) throw std::logic_error("klu plan: piece / update mismatch");
                        for (i32 r = 0; r < P.pc_nrows[pc]; r++)
                            dd[P.pc_r0[pc] + r] = (uint16_t)(P.dest[P.upd_dest[ulast] + P.pc_i0[pc] + r] - cb);
                    }
            // in-wave blob
            std::vector<uint32_t> hdr(2 * KLU_WAVE_WARPS, 0), upd;
            std::vector<uint16_t> dst;
            for (i32 pq = k0; pq < k1; pq++) {
                const i32 c = P.ne_cols[pq];
                const i64 cb = P.cbeg[c];
                hdr[2 * (pq - k0)] = (uint32_t)(upd.size() / 4);
                for (i64 u = P.upd_split[c]; u < P.upd_end[c]; u++) {
                    {
                        const i32 j = P.upd_src[u];
                        const uint32_t srcrow0 = (uint32_t)(P.col_roff[j] + (P.lslot0[j] - P.cbeg[j]));   // first L row of the source in xs
                        if (P.ne_pos[j] < k0 || P.ne_pos[j] >= pq) throw std::logic_error("klu plan: in-wave source outside the wave");
                        upd.push_back((uint32_t)(P.ne_pos[j] - k0) | (srcrow0 << 8));
                    }
                    upd.push_back((uint32_t)(P.upd_uslot[u] - cb));
                    upd.push_back((uint32_t)P.upd_cnt[u]);
                    upd.push_back((uint32_t)dst.size());
                    for (i32 r = 0; r < P.upd_cnt[u]; r++) dst.push_back((uint16_t)(P.dest[P.upd_dest[u] + r] - cb));
                }
                hdr[2 * (pq - k0) + 1] = (uint32_t)(upd.size() / 4) - hdr[2 * (pq - k0)];
            }
            while (dst.size() % 8) dst.push_back(0);
            const size_t before = P.wblob.size();
            P.wblob.insert(P.wblob.end(), hdr.begin(), hdr.end());
            P.wblob.insert(P.wblob.end(), upd.begin(), upd.end());
            for (size_t q = 0; q < dst.size(); q += 2) P.wblob.push_back((uint32_t)dst[q] | ((uint32_t)dst[q + 1] << 16));
            if ((P.wblob.size() - before) * 4 > (size_t)KLU_BLOB_BYTES && k1 - k0 > 1) throw std::logic_error("klu plan: in-wave blob too large");
            P.wblob_ptr.push_back((i64)(P.wblob.size() / 4));
        }
    }
    lap("staged tables + blobs");
    // level schedule: column k depends on every column j with U(j,k) != 0
    std::vector<i32> level(n, 0);
    P.nlevels = 0;
    for (i32 k = 0; k < n; k++) {
        i32 lv = 0;
        for (i64 p = N.Up[k]; p < N.Up[k + 1] - 1; p++) lv = std::max(lv, level[N.Ui[p]] + 1);
        level[k] = lv;
        P.nlevels = std::max(P.nlevels, lv + 1);
    }
    P.level_ptr.assign(P.nlevels + 1, 0);
    for (i32 k = 0; k < n; k++) P.level_ptr[level[k] + 1]++;
    for (i32 l = 0; l < P.nlevels; l++) P.level_ptr[l + 1] += P.level_ptr[l];
    P.level_cols.resize(n);
    {
        std::vector<i32> pos(P.level_ptr.begin(), P.level_ptr.end() - 1);
        for (i32 k = 0; k < n; k++) P.level_cols[pos[level[k]]++] = k;
    }
}

// Host interpreter of the wave schedule: executes, for ONE matrix, exactly the tables k_klu_refactor_wave replays (row
// scaling, gather through wave_rowsrc, TMA copy descriptors seg_*, piece records in bentry, in-wave blobs) followed by the
// dense trailing block as a plain unpivoted elimination.  Test hook for the host-built plan; never a product path.
int klu_plan_emulate(const KluSymbolic& S, const KluPlan& P, const double* Ax, double* LU, double* Rs) {
    const i32 n = P.n;
    int status = 0;
    std::vector<double> As((size_t)std::max<i64>(P.nnzA, 1));
    for (i32 i = 0; i < n; i++) {
        double m = 0.0;
        for (i64 p = P.rowptr[i]; p < P.rowptr[i + 1]; p++) m = std::max(m, std::fabs(Ax[P.rowent[p]]));
        Rs[i] = m > 0.0 ? m : 1.0;
        for (i64 p = P.rowptr[i]; p < P.rowptr[i + 1]; p++) As[P.rowent[p]] = Ax[P.rowent[p]] / Rs[i];
    }
    for (i64 v = 0; v < P.nslots; v++) LU[v] = 0.0;
    for (i64 v = P.lu_slots; v < P.nslots; v++) LU[v] = P.slot_src[v] >= 0 ? As[P.slot_src[v]] : 0.0;
    if (!P.wave_ok) return ST_INVALID;
    std::vector<double> xs(KLU_WAVE_ROWS, 0.0), stage(KLU_CHUNK_ROWS, 0.0);
    const size_t BST = (size_t)KLU_WAVE_WARPS * KLU_REC_U32 + KLU_CHUNK_ROWS;
    const i32 nw = (i32)P.wave_col0.size() - 1;
    (void)S;
    // early columns, level by level, as k_klu_early does: scaled input values, the updates in list order, pivot, scaling
    for (size_t e = 0; e < P.ecols.size(); e++) {
        const i32 k = P.ecols[e];
        const i64 cb = P.cbeg[k], ce = P.cbeg[k + 1];
        for (i64 sl = cb; sl < ce; sl++) LU[sl] = P.slot_src[sl] >= 0 ? As[P.slot_src[sl]] : 0.0;
        for (i64 u = P.upd_ptr[k]; u < P.upd_ptr[k + 1]; u++) {
            if (!P.early[P.upd_src[u]]) return ST_INVALID;
            const double ujk = LU[P.upd_uslot[u]];
            for (i32 t = 0; t < P.upd_cnt[u]; t++) LU[P.dest[P.upd_dest[u] + t]] = std::fma(-LU[P.upd_lslot[u] + t], ujk, LU[P.dest[P.upd_dest[u] + t]]);
        }
        const double piv = LU[P.udiag_slot[k]];
        if (!(std::fabs(piv) > 0.0)) status = ST_SINGULAR;
        const double rpiv = 1.0 / piv;
        for (i64 sl = P.lslot0[k]; sl < ce; sl++) LU[sl] *= rpiv;
    }
    for (i32 w = 0; w < nw; w++) {
        const i32 k0 = P.wave_col0[w], k1 = P.wave_col0[w + 1], wc = k1 - k0;       // positions in ne_cols
        const i32 wrows = P.wave_rows[w];
        for (i32 r = 0; r < wrows; r++) {
            const i32 src = P.wave_rowsrc[(size_t)w * KLU_WAVE_ROWS + r];
            xs[r] = src >= 0 ? As[src] : 0.0;
        }
        for (i64 bi = P.wbatch_ptr[w]; bi < P.wbatch_ptr[w + 1]; bi++) {
            std::fill(stage.begin(), stage.end(), std::nan(""));
            for (i64 sg = P.bseg_ptr[bi]; sg < P.bseg_ptr[bi + 1]; sg++)
                for (i32 r = 0; r < P.seg_cnt[sg]; r++) stage[P.seg_row[sg] + r] = LU[P.lslot0[P.seg_src[sg]] + P.seg_off[sg] + r];
            for (i32 q = 0; q < wc; q++) {
                const uint32_t* rec = P.bentry.data() + (size_t)bi * BST + (size_t)q * KLU_REC_U32;
                const uint16_t* dd = reinterpret_cast<const uint16_t*>(rec + KLU_REC_HDR);
                double* x = xs.data() + P.col_roff[P.ne_cols[k0 + q]];
                for (uint32_t k = 0; k < rec[0]; k++) {
                    const uint32_t w0 = rec[1 + 2 * k], w1 = rec[2 + 2 * k];
                    const i32 r0 = w0 & 0xff, nrows = (w0 >> 8) & 0xff, uloc = w0 >> 16;
                    const i32 g = w1 & 0xf, s0 = (w1 >> 4) & 0xf, tri0 = (w1 >> 8) & 0xff, hastri = (w1 >> 16) & 1;
                    const i32 gu = g - s0;
                    double u[KLU_SN_MAX];
                    for (i32 a = 0; a < gu; a++) u[a] = x[uloc + a];
                    if (hastri)
                        for (i32 a = 1; a < gu; a++) {
                            for (i32 b = 0; b < a; b++) {
                                const i32 B = s0 + b, A = s0 + a;
                                u[a] = std::fma(-stage[tri0 + B * (g - 1) - B * (B - 1) / 2 + (A - B - 1)], u[b], u[a]);
                            }
                            x[uloc + a] = u[a];
                        }
                    for (i32 i = 0; i < nrows; i++) {
                        double acc = x[dd[r0 + i]];
                        for (i32 a = 0; a < gu; a++) acc = std::fma(-stage[r0 + (s0 + a) * nrows + i], u[a], acc);
                        x[dd[r0 + i]] = acc;
                    }
                }
            }
        }
        const uint32_t* bl = P.wblob.data() + P.wblob_ptr[w] * 4;
        const i32 nupd_wave = (i32)(bl[2 * (wc - 1)] + bl[2 * (wc - 1) + 1]);
        const uint32_t* updl = bl + 2 * KLU_WAVE_WARPS;
        const uint16_t* bdst = reinterpret_cast<const uint16_t*>(updl + 4 * nupd_wave);
        for (i32 q = 0; q < wc; q++) {
            const i32 c = P.ne_cols[k0 + q];
            double* x = xs.data() + P.col_roff[c];
            for (uint32_t ui = bl[2 * q]; ui < bl[2 * q] + bl[2 * q + 1]; ui++) {
                const uint32_t w0 = updl[4 * ui];
                const double uj = x[updl[4 * ui + 1]];
                const i32 cnt = (i32)updl[4 * ui + 2];
                const uint16_t* d = bdst + updl[4 * ui + 3];
                const double* lsrc = xs.data() + (w0 >> 8);
                for (i32 t = 0; t < cnt; t++) x[d[t]] = std::fma(-lsrc[t], uj, x[d[t]]);
            }
            if (c < P.spine0) {
                const i64 cb = P.cbeg[c];
                const double piv = x[P.udiag_slot[c] - cb];
                if (!(std::fabs(piv) > 0.0)) status = ST_SINGULAR;
                const double rpiv = 1.0 / piv;
                for (i64 sl = P.lslot0[c] - cb; sl < P.cbeg[c + 1] - cb; sl++) x[sl] *= rpiv;
            }
        }
        for (i32 rn = P.wrun_ptr[w]; rn < P.wrun_ptr[w + 1]; rn++)          // the wave's bulk stores
            for (i32 r = 0; r < P.wrun_cnt[rn]; r++) LU[P.wrun_slot[rn] + r] = xs[P.wrun_row[rn] + r];
    }
    for (i32 k = P.spine0; k < n; k++) {          // dense trailing block: the updates the wave kernel left out, then the pivot
        for (i64 u = P.upd_end[k]; u < P.upd_ptr[k + 1]; u++) {
            const double ujk = LU[P.upd_uslot[u]];
            for (i32 t = 0; t < P.upd_cnt[u]; t++) LU[P.dest[P.upd_dest[u] + t]] = std::fma(-LU[P.upd_lslot[u] + t], ujk, LU[P.dest[P.upd_dest[u] + t]]);
        }
        const double piv = LU[P.udiag_slot[k]];
        if (!(std::fabs(piv) > 0.0)) status = ST_SINGULAR;
        for (i64 sl = P.lslot0[k]; sl < P.cbeg[k + 1]; sl++) LU[sl] /= piv;
    }
    return status;
}

}  // namespace b200s
