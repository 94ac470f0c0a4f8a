"""CPU tests of the multi-GPU Cholesky mapping (kvxopt_b200/dist.py): ownership, exchange and gather plans are pure
functions of the symbolic plan."""
import ctypes as C

import numpy as np
import pytest

from conftest import lap3d, lower_ccs
from kvxopt_b200 import _lib as L, dist as D

fn = L.fn


def handle_for(nx, ny, nz):
    A = lower_ccs(lap3d(nx, ny, nz))
    n = A.shape[0]
    perm = np.zeros(n, np.int64)
    assert fn["b200s_grid_nd_perm"](nx, ny, nz, 32, L.ptr_i64(perm)) == 0
    cp, ri = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    F = L.vp()
    assert fn["b200s_chol_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), b"L", L.ptr_i64(perm), None, C.byref(F)) == 0
    return F


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_subtree_to_subcube_mapping(world):
    F = handle_for(20, 20, 20)
    lay = D.front_layout(F)
    owner = D.ownership(lay, world)
    ns = len(owner)
    assert owner.min() >= 0 and owner.max() < world
    if world > 1:
        assert len(set(owner.tolist())) == world           # every GPU gets work
    parent = lay["parent"]
    # once a front and its parent have different owners, the front's whole subtree has one owner (a cut edge);
    # a rank's fronts below a cut are never handed back up
    cut_children = [s for s in range(ns) if parent[s] >= 0 and owner[parent[s]] != owner[s]]
    xplan = D.exchange_plan(lay, owner)
    moved = [m[0] for lv in xplan for m in lv]
    assert sorted(moved) == sorted(s for s in cut_children if lay["usize"][s] > 0)
    for l, lv in enumerate(xplan):
        for s, src, dst in lv:
            assert lay["level"][parent[s]] == l and src == owner[s] and dst == owner[parent[s]] and src != dst
    assert len(moved) <= 4 * world                          # only the cut edges communicate
    # work balance of the subtree part: no rank above ~2.5x the mean total work
    w = D.front_work(lay)
    share = np.array([w[owner == r].sum() for r in range(world)])
    assert share.max() <= 2.5 * share.sum() / world + w.max()
    # panel gather covers exactly the fronts not owned by the root rank
    runs = D.gather_plan(lay, owner)
    covered = np.zeros(ns, dtype=bool)
    for rk, s0, s1 in runs:
        assert np.all(owner[s0:s1 + 1] == rk) and rk != 0
        covered[s0:s1 + 1] = True
    assert np.array_equal(covered, owner != 0)
    fn["b200s_chol_free"](F)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_shared_schur_complement_plan(world):
    """split_plan / SplitTables / split_moves: the column tiles of every shared Schur complement are dealt exactly once, to
    ranks of the front's subtree group, the owner works in place and helpers in disjoint scratch ranges, and every helper's
    slab is sent to the owner of the parent front before the parent's level"""
    F = handle_for(24, 24, 24)
    lay = D.front_layout(F)
    owner, g0, g1 = D.ownership(lay, world, with_groups=True)
    assert np.array_equal(owner, D.ownership(lay, world))
    assert np.all((g0 <= owner) & (owner < g1))
    splan = D.split_plan(lay, owner, g0, g1, min_flops=1e5, min_rows=100)
    if world >= 4:          # (two ranks: only the root front has a group of two, and it has no Schur complement)
        assert splan, "the top separators of a 24^3 grid are large enough for these thresholds"
    for s, parts in splan.items():
        mu = int(lay["nr"][s] - lay["nc"][s]) + int(lay["nc"][s] & 1)
        ncj = (mu + D.BTN - 1) // D.BTN
        assert parts[0][0] == owner[s] and len({r for r, _, _ in parts}) == len(parts) >= 2
        assert parts[0][1] == 0 and parts[-1][2] == ncj
        for (r, lo, hi), nxt in zip(parts, parts[1:] + [None]):
            assert g0[s] <= r < g1[s] and lo < hi
            if nxt is not None:
                assert nxt[1] == hi
        # about the same number of tiles each
        T = (mu + D.BT - 1) // D.BT
        tiles = [sum(T - (cj >> 1) for cj in range(lo, hi)) for _, lo, hi in parts]
        assert max(tiles) <= 2 * (sum(tiles) / len(tiles)) + T
    pm, sm = D.split_moves(lay, owner, splan)
    for l, lv in enumerate(pm):
        for s, src, dst in lv:
            assert lay["level"][s] == l and src == owner[s] and dst != src
    for l, lv in enumerate(sm):
        for s, helper, dst, lo, hi in lv:
            p = lay["parent"][s]
            assert lay["level"][p] == l and dst == owner[p] and helper != owner[s]
    for r in range(world):
        st = D.SplitTables(lay, owner, splan, r)
        used = []
        for s, parts in splan.items():
            mine = [q for q in parts if q[0] == r]
            assert st.own[s] == (1 if mine else 0)
            if mine:
                _, lo, hi = mine[0]
                assert (st.lo[s], st.hi[s]) == (lo, hi)
                if r == owner[s]:
                    assert st.base[s] == np.iinfo(np.int64).min
                else:
                    off, cnt = D.slab_range(lay, s, lo, hi)
                    assert st.base[s] >= 0 and st.scratch_off[s] == (st.base[s], cnt, off)
                    assert off + cnt <= lay["usize"][s]
                    used.append((int(st.base[s]), int(st.base[s]) + cnt))
        used.sort()
        assert all(a[1] <= b[0] for a, b in zip(used, used[1:])) and (not used or used[-1][1] == st.scratch_size)
        # fronts that are not shared keep the default: owned fronts complete and in place
        for s in range(len(owner)):
            if s not in splan:
                assert st.own[s] == (owner[s] == r) and st.lo[s] == 0 and st.hi[s] == 0x7fffffff
    # flops accounting: the shares of all ranks add up to the total
    tot = sum(D._work_share(lay, owner, splan, r) for r in range(world))
    assert abs(tot - D.front_work(lay).sum()) <= 1e-9 * tot
    fn["b200s_chol_free"](F)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_solve_moves(world):
    """solve_moves: update vectors travel exactly along the cut edges before the parent's forward level; after a top front's
    backward level its solution entries reach every other rank that owns a descendant; the final gather covers exactly the
    columns not owned by rank 0"""
    F = handle_for(20, 20, 20)
    lay = D.front_layout(F)
    owner, g0, g1 = D.ownership(lay, world, with_groups=True)
    fwd, bwd, gather = D.solve_moves(lay, owner, g0, g1)
    parent, level = lay["parent"], lay["level"]
    ns = len(owner)
    cut = sorted(s for s in range(ns) if parent[s] >= 0 and owner[parent[s]] != owner[s] and lay["nr"][s] > lay["nc"][s])
    assert sorted(m[0] for lv in fwd for m in lv) == cut
    for l, lv in enumerate(fwd):
        for s, src, dst in lv:
            assert level[parent[s]] == l and src == owner[s] and dst == owner[parent[s]]
    # every (ancestor f, descendant owner r != owner[f]) pair is served by a backward move of f, scheduled at level[f]
    need = set()
    for s in range(ns):
        a = parent[s]
        while a >= 0:
            if owner[a] != owner[s]:
                need.add((int(a), int(owner[s])))
            a = parent[a]
    have = {(m[0], m[2]) for lv in bwd for m in lv}
    assert need <= have
    for l, lv in enumerate(bwd):
        for f, src, dst in lv:
            assert level[f] == l and src == owner[f] and dst != src
    covered = np.zeros(lay["n"], dtype=bool)
    for rk, c0, c1 in gather:
        assert rk != 0 and not covered[c0:c1].any()
        covered[c0:c1] = True
    mine0 = np.zeros(lay["n"], dtype=bool)
    for s in range(ns):
        if owner[s] == 0:
            mine0[lay["col0"][s]: lay["col0"][s] + lay["nc"][s]] = True
    assert np.array_equal(covered, ~mine0)
    fn["b200s_chol_free"](F)
