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
