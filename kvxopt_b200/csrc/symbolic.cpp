// Symbolic analysis for the multifrontal supernodal Cholesky plan (host, integer only).
//
// Replaces what the reference obtains from cholmod_l_analyze_p (src/C/cholmod.c:269 in `symbolic`,
// :663 in `linsolve`, :811 in `splinsolve`) after `pack` has selected one triangle
// (src/C/cholmod.c:132-181): fill-reducing ordering, elimination tree, postorder, column counts,
// fundamental + relaxed supernodes and each supernode's row structure.  On top of that it lays
// out what the GPU numeric phase needs: panel offsets, child->parent relative indices, level
// sets, update-matrix lifetimes and the scatter map from the caller's CCS values to panel slots.
//
// Algorithms: Liu's elimination tree with path compression; the Gilbert-Ng-Peyton skeleton-graph
// column counts; CHOLMOD-style relaxed amalgamation thresholds (nrelax/zrelax).  All code is
// written for this engine.
#include "host.hpp"
#include <algorithm>
#include <chrono>
#include <map>
#include <set>
#include <stdexcept>

namespace b200s {

namespace {

// Lower-triangular pattern of P A P' in two forms: by column (rows > j) and by row (cols < i).
struct PermPattern {
    std::vector<i64> lp, up;   // n+1 each
    std::vector<i32> li, ui;   // li: rows below the diagonal of each column; ui: columns left of the diagonal of each row
};

// ui (row lists) is all the elimination tree needs, li (column lists) all that the column counts and the fronts need
template <class Sel>
void build_perm_pattern(i64 n, const i64* colptr, const i64* rowind, Sel use, const std::vector<i32>& iperm,
                        PermPattern& P, bool with_li = true, bool with_ui = true) {
    P.lp.assign(n + 1, 0);
    P.up.assign(n + 1, 0);
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (i == j || !use(i, j)) continue;
            i32 r = iperm[i], c = iperm[j];
            if (r < c) std::swap(r, c);
            if (with_li) P.lp[c + 1]++;
            if (with_ui) P.up[r + 1]++;
        }
    for (i64 j = 0; j < n; j++) { P.lp[j + 1] += P.lp[j]; P.up[j + 1] += P.up[j]; }
    P.li.resize(P.lp[n]);
    P.ui.resize(P.up[n]);
    std::vector<i64> pl(P.lp.begin(), P.lp.end() - 1), pu(P.up.begin(), P.up.end() - 1);
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (i == j || !use(i, j)) continue;
            i32 r = iperm[i], c = iperm[j];
            if (r < c) std::swap(r, c);
            if (with_li) P.li[pl[c]++] = r;
            if (with_ui) P.ui[pu[r]++] = c;
        }
}

// Liu's algorithm: parent[] of the elimination tree from the row structure of the lower triangle.
std::vector<i32> etree(i32 n, const PermPattern& P) {
    std::vector<i32> parent(n, -1), anc(n, -1);
    for (i32 k = 0; k < n; k++)
        for (i64 p = P.up[k]; p < P.up[k + 1]; p++) {
            i32 i = P.ui[p];
            while (i != -1 && i < k) {
                i32 nx = anc[i];
                anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
        }
    return parent;
}

// Depth-first postorder of a forest; children visited in increasing index order.
std::vector<i32> postorder(i32 n, const std::vector<i32>& parent) {
    std::vector<i32> head(n, -1), next(n, -1), post;
    post.reserve(n);
    for (i32 j = n - 1; j >= 0; j--)
        if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    std::vector<i32> stack;
    for (i32 r = 0; r < n; r++) {
        if (parent[r] >= 0) continue;
        stack.push_back(r);
        while (!stack.empty()) {
            i32 v = stack.back();
            i32 c = head[v];
            if (c >= 0) { head[v] = next[c]; stack.push_back(c); }
            else { post.push_back(v); stack.pop_back(); }
        }
    }
    return post;
}

// Column counts of L for a postordered matrix (post = identity), skeleton-graph method.
std::vector<i32> column_counts(i32 n, const PermPattern& P, const std::vector<i32>& parent) {
    std::vector<i32> cc(n, 0), first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
    for (i32 k = 0; k < n; k++) {
        anc[k] = k;
        i32 j = k;
        cc[j] = (first[j] == -1) ? 1 : 0;                 // leaves of the etree start at 1
        for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
    }
    for (i32 j = 0; j < n; j++) {
        if (parent[j] != -1) cc[parent[j]]--;
        for (i64 p = P.lp[j]; p < P.lp[j + 1]; p++) {
            const i32 i = P.li[p];                        // A(i,j) != 0, i > j
            if (first[j] <= maxfirst[i]) continue;        // j is not a leaf of the row subtree of i
            maxfirst[i] = first[j];
            const i32 jprev = prevleaf[i];
            prevleaf[i] = j;
            cc[j]++;
            if (jprev != -1) {                            // subsequent leaf: remove the overlap at the LCA
                i32 q = jprev;
                while (q != anc[q]) q = anc[q];
                for (i32 s = jprev; s != q;) { i32 sp = anc[s]; anc[s] = q; s = sp; }
                cc[q]--;
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    for (i32 j = 0; j < n; j++)
        if (parent[j] != -1) cc[parent[j]] += cc[j];
    return cc;
}

// first-fit interval allocator with coalescing, used to lay update matrices out by lifetime
struct Arena {
    // free blocks indexed by offset (to merge neighbours) and by size (best fit in O(log n); the first-fit scan this
    // replaces was quadratic: 425 ms of the 940 ms analysis of a 200k-column pattern with 91k fronts)
    std::map<i64, i64> free_;                  // offset -> size
    std::set<std::pair<i64, i64>> by_size_;    // (size, offset)
    i64 top = 0;
    void drop(std::map<i64, i64>::iterator it) { by_size_.erase({it->second, it->first}); free_.erase(it); }
    void add(i64 off, i64 sz) { free_[off] = sz; by_size_.insert({sz, off}); }
    i64 alloc(i64 sz) {
        auto bs = by_size_.lower_bound({sz, (i64)-1});
        if (bs != by_size_.end()) {
            const i64 off = bs->second, rem = bs->first - sz;
            drop(free_.find(off));
            if (rem > 0) add(off + sz, rem);
            return off;
        }
        // extend the top, absorbing a free block that touches it
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second == top) {
                const i64 off = last->first;
                drop(last);
                top = off + sz;
                return off;
            }
        }
        const i64 off = top;
        top += sz;
        return off;
    }
    void release(i64 off, i64 sz) {
        auto nx = free_.lower_bound(off);
        if (nx != free_.end() && off + sz == nx->first) { sz += nx->second; drop(nx); }
        auto it = free_.lower_bound(off);
        if (it != free_.begin()) {
            auto pv = std::prev(it);
            if (pv->first + pv->second == off) { off = pv->first; sz += pv->second; drop(pv); }
        }
        add(off, sz);
    }
};

}  // namespace

void chol_analyze(i64 n64, const i64* colptr, const i64* rowind, char uplo, const i64* user_perm,
                  const CholOpts& opts, CholPlan& plan) {
    auto t0 = std::chrono::steady_clock::now();
    const bool tdbg = getenv("B200S_DEBUG") != nullptr;
    auto tlast = t0;
    auto lap = [&](const char* what) {
        if (!tdbg) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200s chol analyze] %-26s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tlast).count());
        tlast = now;
    };
    if (n64 < 0 || n64 > 0x7fffffff - 16) throw std::invalid_argument("matrix order out of range");
    if (uplo == 'l') uplo = 'L';
    if (uplo == 'u') uplo = 'U';
    if (uplo != 'L' && uplo != 'U') throw std::invalid_argument("uplo must be 'L' or 'U'");
    const i32 n = (i32)n64;
    plan = CholPlan();
    plan.n = n;
    if (n > 0 && colptr[0] != 0) throw std::invalid_argument("colptr[0] must be 0");
    for (i32 j = 0; j < n; j++) {
        if (colptr[j + 1] < colptr[j]) throw std::invalid_argument("colptr must be nondecreasing");
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            if (rowind[k] < 0 || rowind[k] >= n) throw std::invalid_argument("row index out of range");
            if (k > colptr[j] && rowind[k] <= rowind[k - 1])
                throw std::invalid_argument("row indices must be strictly increasing within a column");
        }
    }
    plan.nnzA = n > 0 ? colptr[n] : 0;
    if (plan.nnzA > 0x7fffffff - 16) throw std::invalid_argument("too many entries");
    const bool lower = uplo == 'L';
    auto use = [lower](i64 i, i64 j) { return lower ? i >= j : i <= j; };

    // ---- fill-reducing permutation
    std::vector<i32> perm(n);
    if (user_perm && opts.nmethods != 2) {
        std::vector<char> seen(n, 0);
        for (i32 k = 0; k < n; k++) {
            i64 v = user_perm[k];
            if (v < 0 || v >= n || seen[v]) throw std::invalid_argument("p is not a valid permutation");
            seen[v] = 1;
            perm[k] = (i32)v;
        }
    } else if (opts.nmethods == 1 && !user_perm) {
        throw std::invalid_argument("nmethods=1 requires a user permutation");
    } else if (opts.ordering == 1 && !user_perm) {
        for (i32 k = 0; k < n; k++) perm[k] = k;
    } else {
        // nmethods 0/2 with a user permutation: CHOLMOD would compare it with AMD and keep the one
        // with less fill; the comparison is done below on nnz(L).
        perm = amd_order(sym_pattern_from_triangle(n, colptr, rowind, uplo));
    }
    lap("ordering");

    auto analyze_perm = [&](std::vector<i32>& pm, std::vector<i32>& parent, PermPattern& PP, std::vector<i32>& cc) {
        std::vector<i32> ip(n);
        for (i32 k = 0; k < n; k++) ip[pm[k]] = k;
        PermPattern P0;
        build_perm_pattern(n, colptr, rowind, use, ip, P0, false);
        std::vector<i32> par0 = etree(n, P0);
        std::vector<i32> post = postorder(n, par0);
        std::vector<i32> pm2(n);
        for (i32 k = 0; k < n; k++) pm2[k] = pm[post[k]];
        pm.swap(pm2);
        for (i32 k = 0; k < n; k++) ip[pm[k]] = k;
        build_perm_pattern(n, colptr, rowind, use, ip, PP, true, false);
        // the elimination tree of the postordered matrix is the first tree relabelled (no second run of Liu's algorithm)
        {
            std::vector<i32> ipost(n);
            for (i32 k = 0; k < n; k++) ipost[post[k]] = k;
            parent.assign(n, -1);
            for (i32 k = 0; k < n; k++) { const i32 q = par0[post[k]]; parent[k] = q < 0 ? -1 : ipost[q]; }
        }
        for (i32 j = 0; j < n; j++)
            if (parent[j] != -1 && parent[j] <= j) throw std::logic_error("etree is not postordered");
        cc = column_counts(n, PP, parent);
    };

    PermPattern PP;
    std::vector<i32> parent, cc;
    analyze_perm(perm, parent, PP, cc);
    if (user_perm && opts.nmethods == 2) {
        // compare the user's ordering with AMD on predicted nnz(L) (cholmod.options['nmethods'] = 2)
        std::vector<i32> pu(n);
        std::vector<char> seen(n, 0);
        for (i32 k = 0; k < n; k++) {
            i64 v = user_perm[k];
            if (v < 0 || v >= n || seen[v]) throw std::invalid_argument("p is not a valid permutation");
            seen[v] = 1;
            pu[k] = (i32)v;
        }
        PermPattern PPu;
        std::vector<i32> parentu, ccu;
        analyze_perm(pu, parentu, PPu, ccu);
        double la = 0, lu = 0;
        for (i32 j = 0; j < n; j++) { la += cc[j]; lu += ccu[j]; }
        if (lu <= la) { perm.swap(pu); parent.swap(parentu); cc.swap(ccu); std::swap(PP, PPu); }
    }
    plan.perm = perm;
    plan.iperm.resize(n);
    for (i32 k = 0; k < n; k++) plan.iperm[perm[k]] = k;
    plan.parent = parent;
    plan.colcount = cc;
    lap("etree/postorder/colcounts");

    // ---- fundamental supernodes
    std::vector<i32> nchild(n, 0);
    for (i32 j = 0; j < n; j++) if (parent[j] != -1) nchild[parent[j]]++;
    std::vector<i32> fstart;                   // first column of each fundamental supernode
    for (i32 j = 0; j < n; j++) {
        bool join = j > 0 && parent[j - 1] == j && cc[j - 1] == cc[j] + 1 && nchild[j] == 1;
        if (!join) fstart.push_back(j);
    }
    const i32 nf = (i32)fstart.size();
    fstart.push_back(n);
    // ---- relaxed amalgamation: merge a supernode into its parent when it is the parent's last child
    // (contiguous columns) and the explicit zeros stay below the zrelax thresholds.
    std::vector<i32> sfirst(fstart.begin(), fstart.end() - 1), slast(nf);   // column ranges, mutable
    std::vector<double> szeros(nf, 0.0);
    std::vector<char> alive(nf, 1);
    std::vector<i32> f_of_col(n);
    for (i32 s = 0; s < nf; s++) { slast[s] = fstart[s + 1] - 1; for (i32 j = fstart[s]; j < fstart[s + 1]; j++) f_of_col[j] = s; }
    std::vector<double> hgt(nf);
    for (i32 s = 0; s < nf; s++) hgt[s] = cc[sfirst[s]];
    for (i32 s = 0; s < nf; s++) {
        if (!alive[s]) continue;
        i32 pj = parent[slast[s]];
        if (pj == -1) continue;
        i32 p = f_of_col[pj];
        if (sfirst[p] != slast[s] + 1) continue;
        const double ns = slast[s] - sfirst[s] + 1, np = slast[p] - sfirst[p] + 1;
        const double ntot = ns + np;
        const double hnew = ns + hgt[p];                            // rows of the merged trapezoid
        const double newzeros = ns * (hnew - hgt[s]);               // every column of s grows by (hnew - hgt[s])
        const double zeros = szeros[s] + szeros[p] + newzeros;
        const double lnz = ntot * hnew - ntot * (ntot - 1) / 2;     // entries of the merged trapezoid
        bool merge;
        if (ntot <= opts.nrelax[0]) merge = true;
        else {
            double z = zeros / lnz;
            if (ntot <= opts.nrelax[1]) merge = z < opts.zrelax[0];
            else if (ntot <= opts.nrelax[2]) merge = z < opts.zrelax[1];
            else merge = z < opts.zrelax[2];
        }
        if (newzeros == 0) merge = true;                            // free merge (same structure)
        if (opts.max_merge_cols > 0 && ntot > opts.max_merge_cols) merge = false;      // (multi-GPU: keep the top separators apart)
        if (!merge) continue;
        alive[s] = 0;
        sfirst[p] = sfirst[s];
        szeros[p] = zeros;
        hgt[p] = hnew;
        for (i32 j = sfirst[s]; j <= slast[s]; j++) f_of_col[j] = p;
    }
    // ---- final supernode list
    std::vector<Front>& F = plan.fronts;
    plan.sn_of_col.assign(n, -1);
    for (i32 s = 0; s < nf; s++) {
        if (!alive[s]) continue;
        Front f{};
        f.col0 = sfirst[s];
        f.nc = slast[s] - sfirst[s] + 1;
        f.parent = -1;
        F.push_back(f);
    }
    std::sort(F.begin(), F.end(), [](const Front& a, const Front& b) { return a.col0 < b.col0; });
    const i32 ns = (i32)F.size();
    for (i32 s = 0; s < ns; s++)
        for (i32 j = F[s].col0; j < F[s].col0 + F[s].nc; j++) plan.sn_of_col[j] = s;

    lap("supernode partition");
    // ---- row structure of every front: pivots, then A's rows below, then the children's update rows
    std::vector<i32> mark(n, -1);
    std::vector<std::vector<i32>> kids(ns);
    plan.rows.clear();
    std::vector<i32> tmp;
    std::vector<size_t> runs;
    for (i32 s = 0; s < ns; s++) {
        Front& f = F[s];
        const i32 c1 = f.col0 + f.nc;
        tmp.clear();
        for (i32 j = f.col0; j < c1; j++)
            for (i64 p = PP.lp[j]; p < PP.lp[j + 1]; p++) {
                i32 i = PP.li[p];
                if (i >= c1 && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
            }
        // A's rows are sorted once; every child's update rows arrive sorted (a filtered copy of its sorted row list), so
        // the list is a sequence of sorted runs that is merged pairwise instead of sorted from scratch
        std::sort(tmp.begin(), tmp.end());
        runs.clear();
        runs.push_back(0);
        if (!tmp.empty()) runs.push_back(tmp.size());
        for (i32 c : kids[s]) {
            const Front& g = F[c];
            for (i32 k = g.nc; k < g.nr; k++) {
                i32 i = plan.rows[g.rowptr + k];
                if (i >= c1 && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
            }
            if (tmp.size() > runs.back()) runs.push_back(tmp.size());
        }
        while (runs.size() > 2) {
            size_t w = 1;
            for (size_t r = 0; r + 2 < runs.size(); r += 2) {
                std::inplace_merge(tmp.begin() + runs[r], tmp.begin() + runs[r + 1], tmp.begin() + runs[r + 2]);
                runs[w++] = runs[r + 2];
            }
            if ((runs.size() - 1) % 2 == 1) runs[w++] = runs.back();      // odd run out: carried to the next round
            runs.resize(w);
        }
        f.rowptr = (i64)plan.rows.size();
        f.nr = f.nc + (i32)tmp.size();
        for (i32 j = f.col0; j < c1; j++) plan.rows.push_back(j);
        plan.rows.insert(plan.rows.end(), tmp.begin(), tmp.end());
        if (!tmp.empty()) {
            f.parent = plan.sn_of_col[tmp[0]];
            kids[f.parent].push_back(s);
        }
    }
    lap("front row structure");
    // ---- relative indices, levels, storage offsets, flop counts
    plan.child_ptr.assign(ns + 1, 0);
    for (i32 s = 0; s < ns; s++) plan.child_ptr[s + 1] = plan.child_ptr[s] + (i32)kids[s].size();
    plan.child_idx.resize(plan.child_ptr[ns]);
    for (i32 s = 0; s < ns; s++) std::copy(kids[s].begin(), kids[s].end(), plan.child_idx.begin() + plan.child_ptr[s]);
    std::vector<i32> posmap(n, -1);
    i64 relsz = 0;
    for (i32 s = 0; s < ns; s++) { F[s].reloff = relsz; relsz += F[s].nr - F[s].nc; }
    plan.rel.resize(relsz);
    for (i32 p = 0; p < ns; p++) {
        if (kids[p].empty()) continue;
        const Front& fp = F[p];
        for (i32 k = 0; k < fp.nr; k++) posmap[plan.rows[fp.rowptr + k]] = k;
        for (i32 c : kids[p]) {
            const Front& g = F[c];
            for (i32 k = g.nc; k < g.nr; k++) {
                i32 pos = posmap[plan.rows[g.rowptr + k]];
                if (pos < 0) throw std::logic_error("child update row missing from parent front");
                plan.rel[g.reloff + (k - g.nc)] = pos;
            }
        }
        for (i32 k = 0; k < fp.nr; k++) posmap[plan.rows[fp.rowptr + k]] = -1;
    }
    lap("relative indices");
    i64 loff = 0;
    plan.nnzL = 0;
    plan.max_nr = plan.max_nc = 0;
    for (i32 s = 0; s < ns; s++) {
        Front& f = F[s];
        f.level = 0;
        f.ld = (f.nr + 1) & ~1;
        f.loff = loff;
        loff += (i64)f.ld * f.nc;
        loff = (loff + 15) & ~(i64)15;                       // 128-byte aligned panels
        plan.nnzL += (i64)f.nc * f.nr - (i64)f.nc * (f.nc - 1) / 2;
        plan.max_nr = std::max(plan.max_nr, f.nr);
        plan.max_nc = std::max(plan.max_nc, f.nc);
        const double c = f.nc, r = f.nr - f.nc;
        plan.flops_potrf += c * c * c / 3.0;
        plan.flops_trsm += c * c * r;
        plan.flops_syrk += c * r * r;
        if (f.nr > PLAN_SMALL_NR) {
            double panel = 0;
            for (i32 k0 = 0; k0 < f.nc; k0 += PLAN_NB) {
                const double w = std::min(PLAN_NB, f.nc - k0), below = f.nr - (k0 + w);
                panel += w * w * w / 3.0 + w * w * below;
            }
            plan.flops_update += (c * c * c / 3.0 + c * c * r + c * r * r) - panel;
        }
    }
    plan.flops = plan.flops_potrf + plan.flops_trsm + plan.flops_syrk;
    plan.lsize = loff;
    for (i32 s = 0; s < ns; s++)
        if (F[s].parent >= 0) F[F[s].parent].level = std::max(F[F[s].parent].level, F[s].level + 1);
    plan.nlevels = 0;
    for (i32 s = 0; s < ns; s++) plan.nlevels = std::max(plan.nlevels, F[s].level + 1);
    plan.level_ptr.assign(plan.nlevels + 1, 0);
    for (i32 s = 0; s < ns; s++) plan.level_ptr[F[s].level + 1]++;
    for (i32 l = 0; l < plan.nlevels; l++) plan.level_ptr[l + 1] += plan.level_ptr[l];
    plan.level_fronts.resize(ns);
    {
        std::vector<i32> pos(plan.level_ptr.begin(), plan.level_ptr.end() - 1);
        for (i32 s = 0; s < ns; s++) plan.level_fronts[pos[F[s].level]++] = s;
    }
    lap("levels + offsets");
    // ---- update-matrix lifetimes: produced at level(s), consumed while level(parent) runs
    {
        Arena arena;
        std::vector<std::vector<i32>> release_at(plan.nlevels + 1);
        for (i32 l = 0; l < plan.nlevels; l++) {
            for (i32 s : release_at[l]) {
                i64 mu = F[s].nr - F[s].nc + (F[s].nc & 1), ldu = (mu + 1) & ~(i64)1;
                arena.release(F[s].uoff, (ldu * mu + 15) & ~(i64)15);
            }
            for (i32 q = plan.level_ptr[l]; q < plan.level_ptr[l + 1]; q++) {
                Front& f = F[plan.level_fronts[q]];
                i64 m = f.nr - f.nc;
                f.uoff = 0;
                if (m == 0) continue;
                // storage origin is the even row nc - (nc&1) of the front so that 16-byte tile accesses stay aligned
                i64 mu = m + (f.nc & 1), ldu = (mu + 1) & ~(i64)1;
                f.uoff = arena.alloc((ldu * mu + 15) & ~(i64)15);
                release_at[F[f.parent].level + 1].push_back(plan.level_fronts[q]);
            }
        }
        plan.wsize = arena.top;
    }
    lap("workspace arena");
    // ---- scatter map: caller's CCS entry k -> slot in L storage (or -1 when outside the chosen triangle)
    plan.amap.assign(plan.nnzA, -1);
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (!use(i, j)) continue;
            i32 r = plan.iperm[i], c = plan.iperm[j];
            if (r < c) std::swap(r, c);
            const Front& f = F[plan.sn_of_col[c]];
            i32 lr;
            if (r < f.col0 + f.nc) lr = r - f.col0;
            else {
                const i32* b = plan.rows.data() + f.rowptr + f.nc;
                const i32* e = plan.rows.data() + f.rowptr + f.nr;
                const i32* it = std::lower_bound(b, e, r);
                if (it == e || *it != r) throw std::logic_error("matrix entry outside the symbolic structure");
                lr = f.nc + (i32)(it - b);
            }
            plan.amap[k] = f.loff + (i64)(c - f.col0) * f.ld + lr;
        }
    lap("scatter map");
    plan.ms_analyze = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace b200s
