// Fill-reducing orderings computed on the host: an approximate-minimum-degree ordering written for
// this engine (the reference gets its ordering from SuiteSparse AMD through cholmod_l_analyze_p,
// src/C/cholmod.c:269,663 and src/C/amd.c; SuiteSparse is not part of the reference tree) and a
// geometric nested dissection for structured grids (BASELINE config 4).
//
// AMD here follows the published algorithm (Amestoy, Davis, Duff, "An approximate minimum degree
// ordering algorithm", SIMAX 1996): quotient graph with element absorption, approximate external
// degrees |A_i \ i| + |L_p \ i| + sum_e |L_e \ L_p|, aggressive absorption, mass elimination and
// supervariable detection by hashing.  Data layout and code are this project's own.
#include "host.hpp"
#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace b200s {

SymPattern sym_pattern_from_triangle(i64 n, const i64* colptr, const i64* rowind, char uplo) {
    SymPattern G;
    G.n = (i32)n;
    G.ptr.assign(n + 1, 0);
    const bool lower = (uplo == 'L' || uplo == 'l');
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (i == j) continue;
            if (lower ? (i < j) : (i > j)) continue;
            G.ptr[i + 1]++;
            G.ptr[j + 1]++;
        }
    for (i64 j = 0; j < n; j++) G.ptr[j + 1] += G.ptr[j];
    G.idx.resize(G.ptr[n]);
    std::vector<i64> pos(G.ptr.begin(), G.ptr.end() - 1);
    for (i64 j = 0; j < n; j++)
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            i64 i = rowind[k];
            if (i == j) continue;
            if (lower ? (i < j) : (i > j)) continue;
            G.idx[pos[i]++] = (i32)j;
            G.idx[pos[j]++] = (i32)i;
        }
    return G;
}

namespace {

enum : unsigned char { ST_VAR = 0, ST_ELEM = 1, ST_DEAD = 2, ST_GONE = 3, ST_DENSE = 4 };

struct Amd {
    i32 n;
    std::vector<i32> iw;           // per variable: [elements | variables]
    std::vector<i64> pe;
    std::vector<i32> len, elen, nv, degree;
    std::vector<unsigned char> st;
    std::vector<std::vector<i32>> evars;
    std::vector<i64> esize;
    std::vector<i32> head, nxt, prv;     // degree buckets
    std::vector<i64> w;
    std::vector<i32> mark;
    std::vector<i32> mnext, mtail;       // supervariable member chains
    std::vector<i32> hhead, hnext;
    std::vector<i64> hval;

    void list_insert(i32 i) {
        i32 d = degree[i];
        prv[i] = -1;
        nxt[i] = head[d];
        if (head[d] >= 0) prv[head[d]] = i;
        head[d] = i;
    }
    void list_remove(i32 i) {
        i32 d = degree[i];
        if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[d] = nxt[i];
        if (nxt[i] >= 0) prv[nxt[i]] = prv[i];
    }
    void chain_append(i32 into, i32 from) {   // members(into) += members(from)
        mnext[mtail[into]] = from;
        mtail[into] = mtail[from];
    }
};

}  // namespace

std::vector<i32> amd_order(const SymPattern& G) {
    Amd a;
    const i32 n = a.n = G.n;
    std::vector<i32> order;
    order.reserve(n);
    if (n == 0) return order;
    a.iw = G.idx;
    a.pe.assign(G.ptr.begin(), G.ptr.end() - 1);
    a.len.resize(n); a.elen.assign(n, 0); a.nv.assign(n, 1); a.degree.resize(n);
    a.st.assign(n, ST_VAR);
    a.evars.resize(n);
    a.esize.assign(n, 0);
    a.head.assign(n + 1, -1); a.nxt.assign(n, -1); a.prv.assign(n, -1);
    a.w.assign(n, 0);
    a.mark.assign(n, 0);
    a.mnext.assign(n, -1); a.mtail.resize(n);
    a.hhead.assign(n, -1); a.hnext.assign(n, -1); a.hval.assign(n, 0);
    for (i32 i = 0; i < n; i++) { a.len[i] = (i32)(G.ptr[i + 1] - G.ptr[i]); a.mtail[i] = i; }

    // dense nodes go last (they would only slow the quotient-graph updates down)
    const double dense_thr = std::max(16.0, 10.0 * std::sqrt((double)n));
    std::vector<i32> dense_nodes;
    i32 nleft = n;
    for (i32 i = 0; i < n; i++)
        if (a.len[i] > dense_thr) { a.st[i] = ST_DENSE; dense_nodes.push_back(i); nleft--; }
    for (i32 i = 0; i < n; i++) {
        if (a.st[i] != ST_VAR) continue;
        i32 d = 0;
        for (i64 k = a.pe[i]; k < a.pe[i] + a.len[i]; k++) if (a.st[a.iw[k]] == ST_VAR) d++;
        a.degree[i] = d;
        a.list_insert(i);
    }

    i32 mindeg = 0, stamp = 0;
    i64 wflg = 1;
    std::vector<i32> Lp, tmpE, tmpV, bucket_keys;
    auto emit = [&](i32 i) { for (i32 m = i; m >= 0; m = a.mnext[m]) order.push_back(m); };

    while (nleft > 0) {
        while (mindeg <= n && a.head[mindeg] < 0) mindeg++;
        if (mindeg > n) throw std::logic_error("amd: degree lists exhausted");
        const i32 p = a.head[mindeg];
        a.list_remove(p);
        const i32 tag = ++stamp;
        a.mark[p] = tag;
        Lp.clear();
        i64 nvpiv = a.nv[p];
        // ---- form the new element L_p = (A_p U union of L_e, e in E_p) \ p
        for (i64 k = a.pe[p] + a.elen[p]; k < a.pe[p] + a.len[p]; k++) {
            i32 v = a.iw[k];
            if (a.st[v] == ST_VAR && a.mark[v] != tag) { a.mark[v] = tag; Lp.push_back(v); a.list_remove(v); }
        }
        for (i64 k = a.pe[p]; k < a.pe[p] + a.elen[p]; k++) {
            i32 e = a.iw[k];
            if (a.st[e] != ST_ELEM) continue;
            for (i32 v : a.evars[e])
                if (a.st[v] == ST_VAR && a.mark[v] != tag) { a.mark[v] = tag; Lp.push_back(v); a.list_remove(v); }
            a.st[e] = ST_DEAD;
            std::vector<i32>().swap(a.evars[e]);
        }
        a.st[p] = ST_ELEM;
        i64 degLp = 0;
        for (i32 v : Lp) degLp += a.nv[v];
        // ---- pass 1: w[e] - wflg = |L_e \ L_p| for every element adjacent to a variable of L_p
        for (i32 i : Lp) {
            const i32 nvi = a.nv[i];
            for (i64 k = a.pe[i]; k < a.pe[i] + a.elen[i]; k++) {
                i32 e = a.iw[k];
                if (a.st[e] != ST_ELEM) continue;
                if (a.w[e] >= wflg) a.w[e] -= nvi; else a.w[e] = wflg + a.esize[e] - nvi;
            }
        }
        // ---- pass 2: prune lists, approximate degrees, mass elimination
        size_t keep = 0;
        bucket_keys.clear();
        for (size_t q = 0; q < Lp.size(); q++) {
            const i32 i = Lp[q];
            i64 deg = 0, hash = 0;
            tmpE.clear(); tmpV.clear();
            for (i64 k = a.pe[i]; k < a.pe[i] + a.elen[i]; k++) {
                i32 e = a.iw[k];
                if (a.st[e] != ST_ELEM) continue;
                i64 we = a.w[e] - wflg;
                if (we > 0) { deg += we; tmpE.push_back(e); hash += e; }
                else { a.st[e] = ST_DEAD; std::vector<i32>().swap(a.evars[e]); }   // aggressive absorption
            }
            for (i64 k = a.pe[i] + a.elen[i]; k < a.pe[i] + a.len[i]; k++) {
                i32 v = a.iw[k];
                if (a.st[v] == ST_VAR && a.mark[v] != tag) { deg += a.nv[v]; tmpV.push_back(v); hash += v; }
            }
            if (tmpE.empty() && tmpV.empty()) {            // indistinguishable from p: eliminate now
                a.st[i] = ST_GONE;
                a.chain_append(p, i);
                nvpiv += a.nv[i];
                degLp -= a.nv[i];
                a.nv[i] = 0;
                a.len[i] = a.elen[i] = 0;
                continue;
            }
            i64 pos = a.pe[i];
            a.iw[pos++] = p;
            for (i32 e : tmpE) a.iw[pos++] = e;
            a.elen[i] = (i32)(pos - a.pe[i]);
            for (i32 v : tmpV) a.iw[pos++] = v;
            a.len[i] = (i32)(pos - a.pe[i]);
            a.degree[i] = (i32)std::min<i64>(a.degree[i], deg);
            a.hval[i] = hash;
            Lp[keep++] = i;
        }
        Lp.resize(keep);
        // ---- pass 3: supervariable detection (same hash -> compare lists exactly)
        for (i32 i : Lp) {
            i32 b = (i32)(a.hval[i] % n);
            if (a.hhead[b] < 0) bucket_keys.push_back(b);
            a.hnext[i] = a.hhead[b];
            a.hhead[b] = i;
        }
        for (i32 b : bucket_keys) {
            for (i32 i = a.hhead[b]; i >= 0; i = a.hnext[i]) {
                if (a.st[i] != ST_VAR || a.hnext[i] < 0) continue;
                const i32 tg = ++stamp;
                for (i64 k = a.pe[i] + 1; k < a.pe[i] + a.len[i]; k++) a.mark[a.iw[k]] = tg;
                i32 prevj = i;
                for (i32 j = a.hnext[i]; j >= 0; j = a.hnext[j]) {
                    bool same = a.st[j] == ST_VAR && a.len[j] == a.len[i] && a.elen[j] == a.elen[i] &&
                                a.hval[j] == a.hval[i];
                    if (same)
                        for (i64 k = a.pe[j] + 1; k < a.pe[j] + a.len[j]; k++)
                            if (a.mark[a.iw[k]] != tg) { same = false; break; }
                    if (same) {
                        a.nv[i] += a.nv[j];
                        a.nv[j] = 0;
                        a.st[j] = ST_GONE;
                        a.len[j] = a.elen[j] = 0;
                        a.chain_append(i, j);
                        a.hnext[prevj] = a.hnext[j];     // unlink j from the bucket chain
                    } else {
                        prevj = j;
                    }
                }
            }
            a.hhead[b] = -1;
        }
        // the variables of L_p keep mark == tag only if still principal; restore marks clobbered by pass 3
        // (pass 3 stamps list entries, which are never members of L_p: L_p was pruned from all lists)
        // ---- finalize: degrees, degree lists, the new element
        nleft -= (i32)nvpiv;
        keep = 0;
        for (size_t q = 0; q < Lp.size(); q++) {
            const i32 i = Lp[q];
            if (a.st[i] != ST_VAR) continue;
            i64 d = (i64)a.degree[i] + degLp - a.nv[i];
            d = std::min<i64>(d, (i64)nleft - a.nv[i]);
            if (d < 0) d = 0;
            a.degree[i] = (i32)d;
            a.list_insert(i);
            if (d < mindeg) mindeg = (i32)d;
            Lp[keep++] = i;
        }
        Lp.resize(keep);
        a.esize[p] = degLp;
        if (Lp.empty()) a.st[p] = ST_DEAD; else a.evars[p] = Lp;
        a.len[p] = a.elen[p] = 0;
        emit(p);
        wflg += (i64)n + 1;
    }
    std::sort(dense_nodes.begin(), dense_nodes.end(), [&](i32 x, i32 y) {
        i32 dx = (i32)(G.ptr[x + 1] - G.ptr[x]), dy = (i32)(G.ptr[y + 1] - G.ptr[y]);
        return dx != dy ? dx < dy : x < y;
    });
    for (i32 d : dense_nodes) order.push_back(d);
    if ((i32)order.size() != n) throw std::logic_error("amd: ordering is not a permutation");
    return order;
}

namespace {
void nd_rec(i64 nx, i64 ny, i64 x0, i64 x1, i64 y0, i64 y1, i64 z0, i64 z1, i64 leaf, std::vector<i32>& out) {
    const i64 dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
    if (dx <= 0 || dy <= 0 || dz <= 0) return;
    auto emit_box = [&](i64 a0, i64 a1, i64 b0, i64 b1, i64 c0, i64 c1) {
        for (i64 z = c0; z < c1; z++)
            for (i64 y = b0; y < b1; y++)
                for (i64 x = a0; x < a1; x++) out.push_back((i32)(x + nx * (y + ny * z)));
    };
    const i64 longest = std::max(dx, std::max(dy, dz));
    if (dx * dy * dz <= leaf || longest <= 2) { emit_box(x0, x1, y0, y1, z0, z1); return; }
    if (dz == longest) {
        i64 m = z0 + dz / 2;
        nd_rec(nx, ny, x0, x1, y0, y1, z0, m, leaf, out);
        nd_rec(nx, ny, x0, x1, y0, y1, m + 1, z1, leaf, out);
        emit_box(x0, x1, y0, y1, m, m + 1);
    } else if (dy == longest) {
        i64 m = y0 + dy / 2;
        nd_rec(nx, ny, x0, x1, y0, m, z0, z1, leaf, out);
        nd_rec(nx, ny, x0, x1, m + 1, y1, z0, z1, leaf, out);
        emit_box(x0, x1, m, m + 1, z0, z1);
    } else {
        i64 m = x0 + dx / 2;
        nd_rec(nx, ny, x0, m, y0, y1, z0, z1, leaf, out);
        nd_rec(nx, ny, m + 1, x1, y0, y1, z0, z1, leaf, out);
        emit_box(m, m + 1, y0, y1, z0, z1);
    }
}
}  // namespace

std::vector<i32> grid_nd(i64 nx, i64 ny, i64 nz, i64 leaf) {
    std::vector<i32> out;
    out.reserve(nx * ny * nz);
    if (leaf < 1) leaf = 64;
    nd_rec(nx, ny, 0, nx, 0, ny, 0, nz, leaf, out);
    return out;
}

}  // namespace b200s
