"""Multi-GPU sparse Cholesky: subtree-to-subcube mapping of the supernodal elimination tree over the GPUs of one box.

No counterpart in the reference (single process, no GPU); this is SURVEY section 8(e), BASELINE config 4.
Every process (one per GPU, `torchrun`) analyses the same matrix and therefore holds the identical plan.  The tree is
cut top-down: the root front goes to the first rank of the group, its children are split over the two halves of the
group by subtree work, and so on until a group is a single rank, which then owns the whole subtree.  The
factorization is stepped level by level through the C ABI (b200s_chol_factor_begin/level/end); before a level runs,
the update matrices of children that live on another GPU are moved GPU-to-GPU with NCCL send/recv straight between
the W buffers (identical layout on every rank).  Afterwards the panels are gathered on rank 0, which serves the
solves.  Round-1 limitation: a front is never split over several GPUs, so the top log2(G) levels are sequential.

`ownership()` and `exchange_plan()` are pure functions of the plan and are tested on the CPU; `VirtualRanks` runs the
same protocol with several handles on ONE GPU (device-to-device copies instead of NCCL) for single-GPU testing.
"""
import ctypes as C

import numpy as np

from . import _lib as L

fn = L.fn


def bind_to_gpu_numa(local_rank):
    """Pin the calling process to the CPU cores next to GPU `local_rank` (NVML's ideal CPU affinity), so that pinned
    host buffers allocated afterwards are first-touched on the GPU's own NUMA node.  With one process per GPU on an
    8-GPU box, the 961 MB-per-step uploads of the batched KLU path otherwise cross the socket interconnect.  Returns
    the number of cores bound, 0 when NVML or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def front_layout(h):
    inf = L.CholInfo()
    fn["b200s_chol_info"](h, C.byref(inf))
    ns = inf.nsuper
    a = {k: np.zeros(max(ns, 1), dtype=np.int64) for k in ("parent", "level", "nc", "nr", "loff", "lsize", "uoff", "usize")}
    st = fn["b200s_chol_front_layout"](h, *[L.ptr_i64(a[k]) for k in ("parent", "level", "nc", "nr", "loff", "lsize", "uoff", "usize")])
    assert st == 0
    lay = {k: v[:ns] for k, v in a.items()}
    rp, c0 = np.zeros(max(ns, 1), dtype=np.int64), np.zeros(max(ns, 1), dtype=np.int64)
    assert fn["b200s_chol_front_layout2"](h, L.ptr_i64(rp), L.ptr_i64(c0)) == 0
    lay["rowptr"], lay["col0"] = rp[:ns], c0[:ns]       # offsets into the solve work vectors / the permuted solution
    lay["nlevels"] = int(inf.nlevels)
    lay["n"] = int(inf.n)
    return lay


def front_work(lay):
    c = lay["nc"].astype(np.float64)
    r = (lay["nr"] - lay["nc"]).astype(np.float64)
    return c ** 3 / 3.0 + c * c * r + c * r * r


def ownership(lay, world, with_groups=False):
    """subtree-to-subcube: owner[s] in [0, world) for every front; with_groups: also the rank group [g0[s], g1[s]) that works
    on the subtree of s (the owner is its first rank; a group of one below the cut)"""
    ns = len(lay["parent"])
    parent = lay["parent"]
    work = front_work(lay)
    sub = work.copy()
    for s in range(ns):                       # fronts are postordered: children before parents
        if parent[s] >= 0:
            sub[parent[s]] += sub[s]
    kids = [[] for _ in range(ns)]
    roots = []
    for s in range(ns):
        (kids[parent[s]] if parent[s] >= 0 else roots).append(s)
    owner = np.zeros(ns, dtype=np.int64)
    g0 = np.zeros(ns, dtype=np.int64)
    g1 = np.ones(ns, dtype=np.int64)
    first = np.zeros(ns, dtype=np.int64)      # first descendant of s in postorder (subtree = [first[s], s])
    for s in range(ns):
        first[s] = first[kids[s][0]] if kids[s] else s

    stack = [(roots, 0, world)]
    while stack:
        nodes, r0, r1 = stack.pop()
        if not nodes:
            continue
        if r1 - r0 == 1:
            for s in nodes:
                owner[first[s]:s + 1] = r0
                g0[first[s]:s + 1] = r0
                g1[first[s]:s + 1] = r1
            continue
        if len(nodes) == 1:
            s = nodes[0]
            owner[s] = r0
            g0[s], g1[s] = r0, r1
            stack.append((kids[s], r0, r1))
            continue
        # split the set of subtrees into two bins of similar work, the rank group into two halves
        order = sorted(nodes, key=lambda q: -sub[q])
        bins, load = ([], []), [0.0, 0.0]
        for q in order:
            b = 0 if load[0] <= load[1] else 1
            bins[b].append(q)
            load[b] += sub[q]
        mid = r0 + max(1, min(r1 - r0 - 1, int(round((r1 - r0) * load[0] / max(load[0] + load[1], 1e-300)))))
        stack.append((bins[0], r0, mid))
        stack.append((bins[1], mid, r1))
    return (owner, g0, g1) if with_groups else owner


BT, BTN = 128, 64        # row / column tile of the Schur-complement kernel (k_update in csrc/chol_gpu.cu)
# cholmod.engine_options["max_merge_cols"] for factor objects that are going to be distributed: relaxed amalgamation would
# merge a top-level separator into its parent for free (same structure), and the merged front's work is all panel work on
# one GPU; as fronts of their own, the child's Schur complement (most of its flops) is shared by its rank group
DIST_MAX_MERGE_COLS = 2048


def split_plan(lay, owner, g0, g1, min_flops=2.0e10, min_rows=1024):
    """Fronts whose Schur complement C = -L21 L21' is shared by the ranks of their subtree group (they would idle while the
    owner works through the top separators).  For every such front s: its participants (owner first) and their column-tile
    ranges [lo, hi) (tiles of BTN columns, contiguous, about the same number of 128 x 64 tiles each).
    Returns {s: [(rank, lo, hi), ...]} -- a pure function of the plan, identical on every rank."""
    plan = {}
    nc, nr = lay["nc"], lay["nr"]
    for s in range(len(owner)):
        m = int(nr[s] - nc[s])
        g = int(g1[s] - g0[s])
        if g < 2 or m < min_rows or float(nc[s]) * m * m < min_flops:
            continue
        mu = m + int(nc[s] & 1)
        T = (mu + BT - 1) // BT
        ncj = (mu + BTN - 1) // BTN
        tiles = np.array([T - (cj >> 1) for cj in range(ncj)], dtype=np.float64)
        cum = np.concatenate([[0.0], np.cumsum(tiles)])
        ranks = [int(owner[s])] + [r for r in range(int(g0[s]), int(g1[s])) if r != owner[s]]
        cuts = [0]
        for q in range(1, g):
            cuts.append(int(np.searchsorted(cum, cum[-1] * q / g)))
        cuts.append(ncj)
        cuts = [min(max(c, 0), ncj) for c in cuts]
        for q in range(1, len(cuts)):
            cuts[q] = max(cuts[q], cuts[q - 1])
        parts = [(ranks[q], cuts[q], cuts[q + 1]) for q in range(g) if cuts[q + 1] > cuts[q]]
        if len(parts) >= 2:
            plan[s] = parts
    return plan


def slab_range(lay, s, lo, hi):
    """(offset from uoff[s], count) in doubles of the column tiles [lo, hi) of the update matrix of front s"""
    mu = int(lay["nr"][s] - lay["nc"][s]) + int(lay["nc"][s] & 1)
    ldu = (mu + 1) & ~1
    c0, c1 = lo * BTN, min(hi * BTN, mu)
    return c0 * ldu, (c1 - c0) * ldu


class SplitTables:
    """what one rank needs to take part in the shared Schur complements: the arrays for b200s_chol_set_syrk_split, the size
    and layout of its scratch buffer, and the transfer lists (panels before phase 2 of a level, slabs before the parent's level)"""

    def __init__(self, lay, owner, splan, rank):
        ns = len(owner)
        self.own = (owner == rank).astype(np.uint8)          # default: the owned fronts, complete, in place
        self.lo = np.zeros(ns, dtype=np.int32)
        self.hi = np.full(ns, 0x7fffffff, dtype=np.int32)
        self.base = np.full(ns, np.iinfo(np.int64).min, dtype=np.int64)
        self.scratch_size = 0
        self.scratch_off = {}
        for s, parts in splan.items():
            self.own[s] = 0
            for r, lo, hi in parts:
                if r != rank:
                    continue
                self.own[s] = 1
                self.lo[s], self.hi[s] = lo, hi
                if r != owner[s]:
                    off, cnt = slab_range(lay, s, lo, hi)
                    self.base[s] = self.scratch_size
                    self.scratch_off[s] = (self.scratch_size, cnt, off)
                    self.scratch_size += cnt
        self.any = bool(splan)


def split_moves(lay, owner, splan):
    """panel_moves[l]: (s, owner, helper) -- the factored panel of s goes to every helper between the two phases of level l;
    slab_moves[l]: (s, helper, dst, lo, hi) -- the helper's slab of s is added to the update matrix of s on dst = the owner of
    the parent of s before level l = level[parent] runs."""
    nl = lay["nlevels"]
    panel_moves = [[] for _ in range(nl)]
    slab_moves = [[] for _ in range(nl)]
    for s, parts in sorted(splan.items()):
        p = lay["parent"][s]
        for r, lo, hi in parts:
            if r == owner[s]:
                continue
            panel_moves[lay["level"][s]].append((s, int(owner[s]), r))
            if p >= 0:
                slab_moves[lay["level"][p]].append((s, r, int(owner[p]), lo, hi))
    return panel_moves, slab_moves


def exchange_plan(lay, owner):
    """per level l: list of (child, src_rank, dst_rank) transfers that must complete before level l runs"""
    parent, level = lay["parent"], lay["level"]
    plan = [[] for _ in range(lay["nlevels"])]
    for s in range(len(parent)):
        p = parent[s]
        if p >= 0 and owner[p] != owner[s] and lay["usize"][s] > 0:
            plan[level[p]].append((s, int(owner[s]), int(owner[p])))
    return plan


def gather_plan(lay, owner, root=0):
    """contiguous runs of fronts owned by the same rank != root: (rank, first front, last front) for the panel gather"""
    runs = []
    ns = len(owner)
    s = 0
    while s < ns:
        e = s
        while e + 1 < ns and owner[e + 1] == owner[s]:
            e += 1
        if owner[s] != root:
            runs.append((int(owner[s]), s, e))
        s = e + 1
    return runs


def solve_moves(lay, owner, g0, g1, root=0):
    """Transfers of the distributed triangular solves (pure function of the plan):
    fwd[l]:  (child s, src, dst) -- the update vector of s, T[rowptr[s] + nc[s] .. rowptr[s] + nr[s]), goes to the owner of its
             parent before forward level l = level[parent] runs (the cut edges of exchange_plan);
    bwd[l]:  (front f, src, dst) -- after backward level l = level[f] the solution entries X[col0[f] .. + nc[f]) go from the
             owner of f to every other rank of its subtree group (their fronts below need them in their row gathers);
    gather:  (rank, c0, c1) -- column ranges of the permuted solution that `root` collects at the end."""
    parent, level = lay["parent"], lay["level"]
    nl = lay["nlevels"]
    fwd = [[] for _ in range(nl)]
    bwd = [[] for _ in range(nl)]
    for s in range(len(owner)):
        p = parent[s]
        if p >= 0 and owner[p] != owner[s] and lay["nr"][s] > lay["nc"][s]:
            fwd[level[p]].append((s, int(owner[s]), int(owner[p])))
        for r in range(int(g0[s]), int(g1[s])):
            if r != owner[s]:
                bwd[level[s]].append((s, int(owner[s]), r))
    gather = []
    ns = len(owner)
    s = 0
    while s < ns:
        e = s
        while e + 1 < ns and owner[e + 1] == owner[s]:
            e += 1
        if owner[s] != root:
            gather.append((int(owner[s]), int(lay["col0"][s]), int(lay["col0"][e] + lay["nc"][e])))
        s = e + 1
    return fwd, bwd, gather


def _solve_views(h, lay):
    import torch
    Tp, Xp = L.vp(), L.vp()
    st = fn["b200s_chol_solve_buffers"](h, C.byref(Tp), C.byref(Xp))
    if st != 0:
        raise RuntimeError("solve buffers: %s (%s)" % (L.strerror(st), L.last_error()))
    dev = torch.device("cuda", torch.cuda.current_device())
    tsize = int((lay["rowptr"] + lay["nr"]).max()) if len(lay["nr"]) else 0
    Tt = torch.as_tensor(_DevArray(Tp.value, max(tsize, 1)), device=dev)
    Xt = torch.as_tensor(_DevArray(Xp.value, max(lay["n"], 1)), device=dev)
    return Tt, Xt


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _device_views(h, lay):
    import torch
    Lp, Wp = L.vp(), L.vp()
    st = fn["b200s_chol_device_buffers"](h, C.byref(Lp), C.byref(Wp))
    if st != 0:
        raise RuntimeError("device buffers: %s (%s)" % (L.strerror(st), L.last_error()))
    inf = L.CholInfo()
    fn["b200s_chol_info"](h, C.byref(inf))
    dev = torch.device("cuda", torch.cuda.current_device())
    Lt = torch.as_tensor(_DevArray(Lp.value, max(inf.factor_bytes // 8, 1)), device=dev)
    Wt = torch.as_tensor(_DevArray(Wp.value, max(inf.workspace_bytes // 8, 1)), device=dev)
    return Lt, Wt


def _check(st, what):
    if st not in (0, 1):
        raise RuntimeError("%s: %s (%s)" % (what, L.strerror(st), L.last_error()))
    return st


def _install_split(h, st):
    """hand the split tables of one rank to its factor handle; returns the (zeroed) scratch tensor or None"""
    import torch
    if not st.any:
        return None
    scratch = torch.zeros(max(1, st.scratch_size), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    _check(fn["b200s_chol_set_syrk_split"](h, st.own.tobytes(), st.lo.ctypes.data_as(C.POINTER(C.c_int32)),
                                           st.hi.ctypes.data_as(C.POINTER(C.c_int32)), L.ptr_i64(st.base),
                                           C.c_void_p(scratch.data_ptr())), "set_syrk_split")
    return scratch


def _work_share(lay, owner, splan, rank):
    """flops this rank executes: its fronts, minus the Schur-complement tiles it gives away, plus those it takes"""
    w = front_work(lay)
    total = float(w[owner == rank].sum())
    for s, parts in splan.items():
        c = float(lay["nc"][s]); m = float(lay["nr"][s] - lay["nc"][s])
        mu = int(m) + int(lay["nc"][s] & 1)
        T = (mu + BT - 1) // BT
        tiles = {r: sum(T - (cj >> 1) for cj in range(lo, hi)) for r, lo, hi in parts}
        alltiles = float(sum(tiles.values()))
        syrk = c * m * m
        if owner[s] == rank:
            total -= syrk
        if rank in tiles:
            total += syrk * tiles[rank] / alltiles
    return total


class DistCholesky:
    """one instance per rank; `F` is the capsule from kvxopt_b200.cholmod.symbolic on this rank's GPU"""

    def __init__(self, F, world, rank, group=None, split=True, split_args=None):
        from . import cholmod
        self.h, _ = cholmod._factor_handle(F)
        self.F = F
        self.world, self.rank, self.group = world, rank, group
        self.lay = front_layout(self.h)
        self.owner, g0, g1 = ownership(self.lay, world, with_groups=True)
        self.xplan = exchange_plan(self.lay, self.owner)
        self.gplan = gather_plan(self.lay, self.owner)
        self.sfwd, self.sbwd, self.sgather = solve_moves(self.lay, self.owner, g0, g1)
        self.minor = self.lay["n"]
        mine = (self.owner == rank).astype(np.uint8)
        _check(fn["b200s_chol_set_owned"](self.h, mine.tobytes()), "set_owned")
        self.Lt, self.Wt = _device_views(self.h, self.lay)
        # shared Schur complements of the top separators (split=False: every front entirely on its owner, as in round 1)
        self.splan = split_plan(self.lay, self.owner, g0, g1, **(split_args or {})) if split and world > 1 else {}
        self.st = SplitTables(self.lay, self.owner, self.splan, rank)
        self.panel_moves, self.slab_moves = split_moves(self.lay, self.owner, self.splan)
        self.split_levels = {int(self.lay["level"][s]) for s in self.splan}
        self.scratch = _install_split(self.h, self.st)
        self.work_share = float(_work_share(self.lay, self.owner, self.splan, rank))

    def factorize(self, values_ptr, on_device):
        """level-stepped factorization with NCCL exchange of the update matrices; returns (status, minor)"""
        import torch
        import torch.distributed as dist
        lay = self.lay
        if self.scratch is not None:
            self.scratch.zero_()
            torch.cuda.current_stream().synchronize()
        _check(fn["b200s_chol_factor_begin"](self.h, values_ptr, 1 if on_device else 0), "factor_begin")
        import os, time
        trace = [] if os.environ.get("B200S_DIST_TRACE") else None       # (level, what, ms since begin), device-synchronised
        t_begin = time.perf_counter()

        def mark(l, what):
            if trace is not None:
                fn["b200s_chol_sync"](self.h)
                trace.append((l, what, (time.perf_counter() - t_begin) * 1e3))
        self.trace = trace
        for l in range(lay["nlevels"]):
            mine = [m for m in self.xplan[l] if self.rank in (m[1], m[2])]
            slabs = [m for m in self.slab_moves[l] if self.rank in (m[1], m[2])]
            if mine or slabs:
                fn["b200s_chol_sync"](self.h)                 # my update matrices and slabs of earlier levels are complete
                ops, adds = [], []
                for s, src, dst in mine:
                    t = self.Wt[lay["uoff"][s]: lay["uoff"][s] + lay["usize"][s]]
                    if src == self.rank:
                        ops.append(dist.P2POp(dist.isend, t, dst, group=self.group))
                    else:
                        ops.append(dist.P2POp(dist.irecv, t, src, group=self.group))
                for s, helper, dst, lo, hi in slabs:           # slabs of shared Schur complements -> the parent's owner
                    off, cnt = slab_range(lay, s, lo, hi)
                    w0 = int(lay["uoff"][s]) + off
                    if helper == self.rank:
                        so = self.st.scratch_off[s][0]
                        t = self.scratch[so: so + cnt]
                        if dst == self.rank:
                            adds.append((w0, cnt, t))
                        else:
                            ops.append(dist.P2POp(dist.isend, t, dst, group=self.group))
                    else:
                        t = torch.empty(cnt, dtype=torch.float64, device=self.Wt.device)
                        ops.append(dist.P2POp(dist.irecv, t, helper, group=self.group))
                        adds.append((w0, cnt, t))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                for w0, cnt, t in adds:                        # after the update matrix itself has arrived
                    self.Wt[w0: w0 + cnt] += t
                torch.cuda.current_stream().synchronize()      # received data visible to the handle's stream
                mark(l, "exchange")
            if l in self.split_levels:
                _check(fn["b200s_chol_factor_level_phase"](self.h, l, 1), "factor_level (panels)")
                mark(l, "panels")
                pm = [m for m in self.panel_moves[l] if self.rank in (m[1], m[2])]
                if pm:
                    fn["b200s_chol_sync"](self.h)             # the panels of my shared fronts are factored
                    ops = []
                    for s, src, dst in pm:
                        t = self.Lt[lay["loff"][s]: lay["loff"][s] + lay["lsize"][s]]
                        ops.append(dist.P2POp(dist.isend if src == self.rank else dist.irecv, t, dst if src == self.rank else src,
                                              group=self.group))
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                    torch.cuda.current_stream().synchronize()
                    mark(l, "panel copies")
                _check(fn["b200s_chol_factor_level_phase"](self.h, l, 2), "factor_level (Schur complements)")
                mark(l, "schur")
            else:
                _check(fn["b200s_chol_factor_level"](self.h, l), "factor_level")
                if l >= lay["nlevels"] - 8:
                    mark(l, "level")
        minor = C.c_int64()
        st = _check(fn["b200s_chol_factor_end"](self.h, C.byref(minor)), "factor_end")
        m = torch.tensor([int(minor.value) if st == 1 else lay["n"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(m, op=dist.ReduceOp.MIN, group=self.group)
        self.minor = int(m.item())
        return (1 if self.minor < lay["n"] else 0), self.minor

    def solve(self, b):
        """Distributed solve of A x = b with the panels where they were factored (no gather_factor needed): `b` is a CUDA
        float64 tensor of length n, meaningful on rank 0 (it is broadcast); returns x as a CUDA tensor on rank 0 (None
        elsewhere).  Forward sweep leaves to root with the update vectors of the cut edges sent up, backward sweep root to
        leaves with the solution entries of the top fronts sent down their subtree groups."""
        import torch
        import torch.distributed as dist
        lay, h, rank = self.lay, self.h, self.rank
        if self.minor < lay["n"]:
            raise ArithmeticError("singular matrix")
        if not hasattr(self, "Tt"):
            _check(fn["b200s_chol_set_numeric"](h, 1, lay["n"]), "set_numeric")     # this rank's panels are numeric
            self.Tt, self.Xt = _solve_views(h, lay)
        bb = b if rank == 0 else torch.empty(lay["n"], dtype=torch.float64, device=self.Xt.device)
        dist.broadcast(bb, src=0, group=self.group)
        torch.cuda.current_stream().synchronize()
        _check(fn["b200s_chol_solve_dist_begin"](h, C.c_void_p(bb.data_ptr())), "solve_begin")

        def exchange(moves, view, rng):
            mine = [m for m in moves if rank in (m[1], m[2])]
            if not mine:
                return
            fn["b200s_chol_sync"](h)
            ops = []
            for s_, src, dst in mine:
                a, e = rng(s_)
                t = view[a:e]
                ops.append(dist.P2POp(dist.isend if src == rank else dist.irecv, t, dst if src == rank else src, group=self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            torch.cuda.current_stream().synchronize()
        upd = lambda s_: (int(lay["rowptr"][s_] + lay["nc"][s_]), int(lay["rowptr"][s_] + lay["nr"][s_]))
        cols = lambda s_: (int(lay["col0"][s_]), int(lay["col0"][s_] + lay["nc"][s_]))
        for l in range(lay["nlevels"]):
            exchange(self.sfwd[l], self.Tt, upd)
            _check(fn["b200s_chol_solve_dist_level"](h, 0, l), "solve_level (forward)")
        for l in range(lay["nlevels"] - 1, -1, -1):
            _check(fn["b200s_chol_solve_dist_level"](h, 1, l), "solve_level (backward)")
            exchange(self.sbwd[l], self.Xt, cols)
        fn["b200s_chol_sync"](h)
        ops = []
        for rk, c0, c1 in self.sgather:
            if rank == rk:
                ops.append(dist.P2POp(dist.isend, self.Xt[c0:c1], 0, group=self.group))
            elif rank == 0:
                ops.append(dist.P2POp(dist.irecv, self.Xt[c0:c1], rk, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        torch.cuda.current_stream().synchronize()
        if rank != 0:
            return None
        x = torch.empty(lay["n"], dtype=torch.float64, device=self.Xt.device)
        _check(fn["b200s_chol_solve_dist_end"](h, C.c_void_p(x.data_ptr())), "solve_end")
        return x

    def gather_factor(self, root=0):
        """collect all panels on `root` (contiguous runs of fronts per owner) and mark its factor numeric"""
        import torch
        import torch.distributed as dist
        lay = self.lay
        ops = []
        for rk, s0, s1 in self.gplan:
            a, b = lay["loff"][s0], lay["loff"][s1] + lay["lsize"][s1]
            t = self.Lt[a:b]
            if self.rank == rk:
                ops.append(dist.P2POp(dist.isend, t, root, group=self.group))
            elif self.rank == root:
                ops.append(dist.P2POp(dist.irecv, t, rk, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        torch.cuda.current_stream().synchronize()
        if self.rank == root:
            _check(fn["b200s_chol_set_numeric"](self.h, 1 if self.minor >= lay["n"] else 0, self.minor), "set_numeric")


class VirtualRanks:
    """The same protocol with `world` handles on ONE GPU (device-to-device copies instead of NCCL): lets the ownership
    split, the exchange plan and the level-stepped kernels be verified on a single-GPU box."""

    def __init__(self, capsules, split=True, split_args=None):
        from . import cholmod
        self.hs = [cholmod._factor_handle(F)[0] for F in capsules]
        self.world = len(self.hs)
        self.lay = front_layout(self.hs[0])
        self.owner, g0, g1 = ownership(self.lay, self.world, with_groups=True)
        self.xplan = exchange_plan(self.lay, self.owner)
        self.gplan = gather_plan(self.lay, self.owner)
        self.splan = split_plan(self.lay, self.owner, g0, g1, **(split_args or {})) if split and self.world > 1 else {}
        self.panel_moves, self.slab_moves = split_moves(self.lay, self.owner, self.splan)
        self.split_levels = {int(self.lay["level"][s]) for s in self.splan}
        self.sfwd, self.sbwd, self.sgather = solve_moves(self.lay, self.owner, g0, g1)
        self.sviews = None
        self.views, self.sts, self.scratch = [], [], []
        for r, h in enumerate(self.hs):
            _check(fn["b200s_chol_set_owned"](h, (self.owner == r).astype(np.uint8).tobytes()), "set_owned")
            self.views.append(_device_views(h, self.lay))
            self.sts.append(SplitTables(self.lay, self.owner, self.splan, r))
            self.scratch.append(_install_split(h, self.sts[-1]))

    def factorize(self, values):
        import torch
        lay = self.lay
        v = np.ascontiguousarray(values, dtype=np.float64)
        for sc in self.scratch:
            if sc is not None:
                sc.zero_()
        torch.cuda.synchronize()
        for h in self.hs:
            _check(fn["b200s_chol_factor_begin"](h, v.ctypes.data_as(C.c_void_p), 0), "factor_begin")
        for l in range(lay["nlevels"]):
            if self.xplan[l] or self.slab_moves[l]:
                for h in self.hs:
                    fn["b200s_chol_sync"](h)
                for s, src, dst in self.xplan[l]:
                    a, b = lay["uoff"][s], lay["uoff"][s] + lay["usize"][s]
                    self.views[dst][1][a:b].copy_(self.views[src][1][a:b])
                for s, helper, dst, lo, hi in self.slab_moves[l]:
                    off, cnt = slab_range(lay, s, lo, hi)
                    w0 = int(lay["uoff"][s]) + off
                    so = self.sts[helper].scratch_off[s][0]
                    self.views[dst][1][w0: w0 + cnt] += self.scratch[helper][so: so + cnt]
                torch.cuda.synchronize()
            if l in self.split_levels:
                for h in self.hs:
                    _check(fn["b200s_chol_factor_level_phase"](h, l, 1), "factor_level (panels)")
                for h in self.hs:
                    fn["b200s_chol_sync"](h)
                for s, src, dst in self.panel_moves[l]:
                    a, b = lay["loff"][s], lay["loff"][s] + lay["lsize"][s]
                    self.views[dst][0][a:b].copy_(self.views[src][0][a:b])
                torch.cuda.synchronize()
                for h in self.hs:
                    _check(fn["b200s_chol_factor_level_phase"](h, l, 2), "factor_level (Schur complements)")
            else:
                for h in self.hs:
                    _check(fn["b200s_chol_factor_level"](h, l), "factor_level")
        minors = []
        for h in self.hs:
            m = C.c_int64()
            st = _check(fn["b200s_chol_factor_end"](h, C.byref(m)), "factor_end")
            minors.append(int(m.value) if st == 1 else lay["n"])
        self.minor = min(minors)
        for rk, s0, s1 in self.gplan:
            a, b = lay["loff"][s0], lay["loff"][s1] + lay["lsize"][s1]
            self.views[0][0][a:b].copy_(self.views[rk][0][a:b])
        torch.cuda.synchronize()
        _check(fn["b200s_chol_set_numeric"](self.hs[0], 1 if self.minor >= lay["n"] else 0, self.minor), "set_numeric")
        return self.minor

    def solve(self, b):
        """the distributed solve protocol of DistCholesky.solve with device-to-device copies (every handle keeps its own panels:
        call it right after factorize(); the panel gather is not needed for it)"""
        import torch
        lay = self.lay
        dev = torch.device("cuda", torch.cuda.current_device())
        bb = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64).reshape(-1), device=dev)
        if self.sviews is None:
            for h in self.hs:
                _check(fn["b200s_chol_set_numeric"](h, 1, lay["n"]), "set_numeric")
            self.sviews = [_solve_views(h, lay) for h in self.hs]
        for h in self.hs:
            _check(fn["b200s_chol_solve_dist_begin"](h, C.c_void_p(bb.data_ptr())), "solve_begin")

        def sync_all():
            for h in self.hs:
                fn["b200s_chol_sync"](h)
        for l in range(lay["nlevels"]):
            if self.sfwd[l]:
                sync_all()
                for s_, src, dst in self.sfwd[l]:
                    a, e = int(lay["rowptr"][s_] + lay["nc"][s_]), int(lay["rowptr"][s_] + lay["nr"][s_])
                    self.sviews[dst][0][a:e].copy_(self.sviews[src][0][a:e])
                torch.cuda.synchronize()
            for h in self.hs:
                _check(fn["b200s_chol_solve_dist_level"](h, 0, l), "solve_level (forward)")
        for l in range(lay["nlevels"] - 1, -1, -1):
            for h in self.hs:
                _check(fn["b200s_chol_solve_dist_level"](h, 1, l), "solve_level (backward)")
            if self.sbwd[l]:
                sync_all()
                for s_, src, dst in self.sbwd[l]:
                    a, e = int(lay["col0"][s_]), int(lay["col0"][s_] + lay["nc"][s_])
                    self.sviews[dst][1][a:e].copy_(self.sviews[src][1][a:e])
                torch.cuda.synchronize()
        sync_all()
        for rk, c0, c1 in self.sgather:
            self.sviews[0][1][c0:c1].copy_(self.sviews[rk][1][c0:c1])
        torch.cuda.synchronize()
        x = torch.empty(lay["n"], dtype=torch.float64, device=dev)
        _check(fn["b200s_chol_solve_dist_end"](self.hs[0], C.c_void_p(x.data_ptr())), "solve_end")
        return x.cpu().numpy()
