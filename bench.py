#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 sparse direct-solve hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload klu|chol|all]

Headline workload (BASELINE.json configs[1]): batched KLU numeric refactorization of same-pattern value
perturbations of the ACTIVSg2000 power-flow Jacobian, 4096 matrices per GPU (weak scaling: every rank owns its
own 4096 matrices, no data-path collective).  metric = refactors/s.
  value : whole-job refactors/s with the value arrays already resident in HBM (b200s_klu_refactor_batch_dev),
          device-event time, max over ranks.
  e2e   : the same through the public API kvxopt_b200.klu.refactor_batch with HOST (pinned) buffers: H2D of the
          values and D2H of the per-matrix status inside the timed region.
At N=1 the same line also carries `cholesky` (BASELINE configs[3]: 100^3 7-point Laplacian, nested dissection,
supernodal Cholesky factor+solve ms and FP64 TFLOP/s, with the tensor-pipe roofline of the DMMA update kernel).
--impl reference times the CPU restatement of the reference's KLU path (oracle/, klu_refactor semantics) on all
host cores of this box; SuiteSparse itself is not installable here.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def load_activsg():
    z = np.load(os.path.join(GOLD, "ACTIVSg2000.npz"))
    n = int(z["n"])
    A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n))
    A.sort_indices()
    return A


def perturbed_values(base, batch, rank, out=None):
    """a_k (1 + 1e-3 u_k), u ~ U(-1,1), generator seeded by (rank, block) so every rank owns different matrices"""
    nnz = base.size
    if out is None:
        out = np.empty((batch, nnz), dtype=np.float64)
    blk = 256
    for b0 in range(0, batch, blk):
        rng = np.random.default_rng([20261018, rank, b0])
        b1 = min(batch, b0 + blk)
        u = rng.uniform(-1.0, 1.0, size=(b1 - b0, nnz))
        np.multiply(u, 1e-3, out=u)
        u += 1.0
        np.multiply(u, base[None, :], out=out[b0:b1])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    hbm, src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm, src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    fp64 = 36.9
    try:
        fp64 = float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))["dmma_m8n8k4_w8_tflops"])
    except Exception:
        pass
    return hbm, src, fp64


def traffic_from_profiles(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's KLU path on all host cores
# ------------------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(cp, ri, vx, P0, Q):
    from oracle import KluOracle
    _W["o"] = KluOracle(len(cp) - 1, cp, ri, vx, P0=P0, Q=Q)
    _W["base"] = vx


def _worker_run(args):
    seed, count = args
    o, base = _W["o"], _W["base"]
    rng = np.random.default_rng(seed)
    vals = [base * (1 + 1e-3 * rng.uniform(-1, 1, base.size)) for _ in range(count)]
    t0 = time.perf_counter()
    for v in vals:
        o.refactor(v)
    return time.perf_counter() - t0


def host_pattern(A):
    """pivot order of the product's host analysis (no GPU involved): gives the CPU port the same ordering"""
    from kvxopt_b200 import _lib as L
    fn = L.fn
    n = A.shape[0]
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
    N = L.vp(); assert fn["b200s_klu_pivot_host"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)) == 0
    P = np.zeros(n, np.int64); Q = np.zeros(n, np.int64)
    fn["b200s_klu_extract"](N, None, None, None, None, None, None, None, None, None, L.ptr_i64(P), L.ptr_i64(Q), None, None)
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf))
    d = inf.asdict()
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)
    return cp, ri, vx, P, Q, d


def cpu_refactor_rate(A, cores, per_worker, steps, warmup):
    import multiprocessing as mp
    cp, ri, vx, P, Q, d = host_pattern(A)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(cp, ri, vx, P, Q)) as pool:
        for w in range(warmup):
            pool.map(_worker_run, [(1000 + w * cores + c, 2) for c in range(cores)])
        times = []
        for s in range(steps):
            t0 = time.perf_counter()
            pool.map(_worker_run, [(s * cores + c, per_worker) for c in range(cores)])
            times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    return cores * per_worker * steps / total, total / steps * 1e3, d


def run_reference(args, rank, world):
    if rank != 0:
        return
    A = load_activsg()
    cores = os.cpu_count() or 1
    per_worker = 24
    rate, ms_step, d = cpu_refactor_rate(A, cores, per_worker, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "batched KLU refactors/sec", "value": rate, "unit": "refactors/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "klu_refactor_batch ACTIVSg2000 (n=4000, nnz=29336) same-pattern perturbations 1e-3",
                   "batch_per_step": cores * per_worker, "ordering": "same BTF+AMD ordering and pivot order as the GPU arm"},
        "cpu_baseline": {"value": rate, "unit": "refactors/s", "cores": cores, "kind": "port",
                         "sample": "%d refactorizations per step (%d per core), oracle/klu_oracle.c klu_refactor "
                                   "restatement; SuiteSparse KLU is not installable in this image" % (cores * per_worker, per_worker)},
        "e2e": {"value": rate, "unit": "refactors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# Cholesky (configs[3]) -- reported inside the N=1 line
# ------------------------------------------------------------------------------------------------------------
def lap3d_lower(nx):
    """lower triangle of the 7-point Laplacian on an nx^3 grid (diag 6, off-diag -1), x fastest, as CCS"""
    n = nx ** 3
    idx = np.arange(n, dtype=np.int64)
    x = idx % nx; y = (idx // nx) % nx; z = idx // (nx * nx)
    rows = [idx]; cols = [idx]; vals = [np.full(n, 6.0)]
    for mask, off in ((x + 1 < nx, 1), (y + 1 < nx, nx), (z + 1 < nx, nx * nx)):
        rows.append(idx[mask] + off); cols.append(idx[mask]); vals.append(np.full(int(mask.sum()), -1.0))
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    A.sort_indices()
    return A


def bench_cholesky(nx, steps, fp64_peak):
    import torch
    from kvxopt_b200 import _lib as L, cholmod
    fn = L.fn
    Al = lap3d_lower(nx)
    n = Al.shape[0]
    perm = np.zeros(n, np.int64)
    fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
    t0 = time.perf_counter()
    F = cholmod.symbolic(Al, p=perm)
    t_analyze = time.perf_counter() - t0
    h, _ = cholmod._factor_handle(F)
    vals_host = torch.from_numpy(Al.data.copy()).pin_memory()
    vals_dev = vals_host.cuda()
    torch.cuda.synchronize()
    minor = C.c_int64()
    inf = L.CholInfo()
    # warm-up (allocates L and the update workspace), then timed device-resident factorizations
    assert fn["b200s_chol_factorize_dev"](h, vals_dev.data_ptr(), C.byref(minor)) == 0, L.last_error()
    ms_f = []
    for _ in range(steps):
        assert fn["b200s_chol_factorize_dev"](h, vals_dev.data_ptr(), C.byref(minor)) == 0
        fn["b200s_chol_info"](h, C.byref(inf))
        ms_f.append(inf.ms_total)
    B = np.random.default_rng(0).standard_normal((n, 1))
    Bd = torch.from_numpy(B[:, 0].copy()).cuda()
    ms_s = []
    for _ in range(max(2, steps)):
        Xd = Bd.clone()
        torch.cuda.synchronize()
        assert fn["b200s_chol_solve_dev"](h, 0, Xd.data_ptr(), 1, n) == 0
        fn["b200s_chol_info"](h, C.byref(inf))
        ms_s.append(inf.ms_solve)
    x = Xd.cpu().numpy()
    A = (Al + sp.tril(Al, -1).T).tocsr()
    berr = float(np.linalg.norm(A @ x - B[:, 0]) / (12.0 * np.linalg.norm(x) + np.linalg.norm(B)))
    # end to end through the public API with host buffers (H2D of values and RHS, D2H of the solution)
    Xh = np.asfortranarray(B.copy())
    t0 = time.perf_counter()
    cholmod.numeric(Al, F)
    cholmod.solve(F, Xh)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # one profiled factorization: per-kernel-class device time (events around every launch)
    fn["b200s_chol_set_profiling"](h, 1)
    assert fn["b200s_chol_factorize_dev"](h, vals_dev.data_ptr(), C.byref(minor)) == 0
    fn["b200s_chol_info"](h, C.byref(inf))
    fn["b200s_chol_set_profiling"](h, 0)
    d = inf.asdict()
    best_f, best_s = float(np.min(ms_f)), float(np.min(ms_s))
    upd_tf = d["flops_update"] / (d["ms_dense_update"] * 1e-3) / 1e12 if d["ms_dense_update"] > 0 else None
    out = {
        "workload": "7-point Laplacian %d^3, geometric nested dissection (leaf 64), supernodal LL^T, 1 RHS" % nx,
        "n": n, "nnz_L": d["nnz_L"], "nsuper": d["nsuper"], "levels": d["nlevels"], "max_front": [d["max_front_rows"], d["max_front_cols"]],
        "flops": d["flops"], "analyze_ms_host": t_analyze * 1e3,
        "factor_ms": best_f, "solve_ms": best_s, "factor_plus_solve_ms": best_f + best_s,
        "factor_tflops": d["flops"] / (best_f * 1e-3) / 1e12,
        "e2e_factor_plus_solve_ms_host_buffers": e2e_ms, "backward_error": berr,
        "kernel_ms_profiled": {"extend_add": d["ms_extend"], "small_fronts": d["ms_potrf"], "panel": d["ms_trsm"],
                               "dmma_update": d["ms_dense_update"]},
        "roofline": {"bound": "tensor", "kernel": "k_update (FP64 DMMA m8n8k4)", "achieved": upd_tf, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": (upd_tf / fp64_peak) if upd_tf else None,
                     "peak_source": "measured: tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak.json)",
                     "traffic": traffic_from_profiles("k_update")},
    }
    del F
    # the signed instantiation of the same kernels (cholmod.options['supernodal'] = 0: LDL' without pivoting, the mode
    # kkt.ldl factors quasi-definite KKT systems with) on the same matrix
    o = L.CholOpts()
    fn["b200s_chol_default_opts"](C.byref(o))
    o.supernodal = 0
    h2 = C.c_void_p()
    cp_, ri_ = np.ascontiguousarray(Al.indptr, dtype=np.int64), np.ascontiguousarray(Al.indices, dtype=np.int64)
    assert fn["b200s_chol_analyze"](n, L.ptr_i64(cp_), L.ptr_i64(ri_), b"L", L.ptr_i64(perm), C.byref(o), C.byref(h2)) == 0
    ms_l = []
    for _ in range(1 + max(2, steps // 2)):
        assert fn["b200s_chol_factorize_dev"](h2, vals_dev.data_ptr(), C.byref(minor)) == 0, L.last_error()
        fn["b200s_chol_info"](h2, C.byref(inf))
        ms_l.append(inf.ms_total)
    Xd = Bd.clone()
    torch.cuda.synchronize()
    assert fn["b200s_chol_solve_dev"](h2, 0, Xd.data_ptr(), 1, n) == 0
    x2 = Xd.cpu().numpy()
    fn["b200s_chol_free"](h2)
    out["ldl_mode"] = {"what": "same matrix, supernodal = 0 (signed kernels: P A P' = L D L' without pivoting)",
                       "factor_ms": float(np.min(ms_l[1:])), "factor_tflops": d["flops"] / (float(np.min(ms_l[1:])) * 1e-3) / 1e12,
                       "backward_error": float(np.linalg.norm(A @ x2 - B[:, 0]) / (12.0 * np.linalg.norm(x2) + np.linalg.norm(B)))}
    return out


def cpu_cholesky_sample(nx):
    """bounded CPU sample of the same workload class: the oracle's supernodal left-looking LL^T (OpenBLAS) on nx^3"""
    from kvxopt_b200 import _lib as L
    from oracle import CholOracle, lib as olib
    Al = lap3d_lower(nx)
    n = Al.shape[0]
    perm = np.zeros(n, np.int64)
    L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
    threads = olib().oracle_blas_threads(os.cpu_count() or 1)
    O = CholOracle(n, Al.indptr, Al.indices, "L", perm)
    t0 = time.perf_counter(); O.factorize(Al.data); tf = time.perf_counter() - t0
    b = np.random.default_rng(0).standard_normal(n)
    t0 = time.perf_counter(); O.solve(b); ts = time.perf_counter() - t0
    return {"workload": "same generator at %d^3" % nx, "factor_ms": tf * 1e3, "solve_ms": ts * 1e3, "flops": O.flops,
            "factor_tflops": O.flops / tf / 1e12, "cores": threads, "kind": "port"}


def bench_cholesky_multi(nx, steps, rank, world):
    """BASELINE configs[3] over N GPUs: subtree-to-subcube (kvxopt_b200/dist.py), update matrices over NCCL/NVLink"""
    import torch
    import torch.distributed as dist
    from kvxopt_b200 import _lib as L, cholmod, dist as D
    Al = lap3d_lower(nx)
    n = Al.shape[0]
    perm = np.zeros(n, np.int64)
    L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
    F = cholmod.symbolic(Al, p=perm)
    dc = D.DistCholesky(F, world, rank)
    vals = torch.from_numpy(Al.data.copy()).cuda()
    times, gtimes = [], []
    for rep in range(steps + 1):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        st, minor = dc.factorize(vals.data_ptr(), True)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        dc.gather_factor(0)
        torch.cuda.synchronize(); dist.barrier()
        t2 = time.perf_counter()
        if rep > 0:
            times.append((t1 - t0) * 1e3); gtimes.append((t2 - t1) * 1e3)
    out = None
    if rank == 0:
        b = np.random.default_rng(0).standard_normal((n, 1)); x = np.asfortranarray(b.copy())
        cholmod.solve(F, x)
        d = cholmod.factor_info(F)
        A = (Al + sp.tril(Al, -1).T).tocsr()
        berr = float(np.linalg.norm(A @ x - b) / (12.0 * np.linalg.norm(x) + np.linalg.norm(b)))
        w = D.front_work(dc.lay)
        out = {"workload": "7-point Laplacian %d^3, nested dissection, subtree-to-subcube over %d GPUs" % (nx, world),
               "factor_ms": float(np.min(times)), "gather_panels_ms": float(np.min(gtimes)),
               "factor_tflops": d["flops"] / (float(np.min(times)) * 1e-3) / 1e12, "solve_ms_rank0": d["ms_solve"],
               "backward_error": berr, "timing": "host clock between barrier+synchronize pairs, max over ranks by construction",
               "work_share_per_rank": [round(float(w[dc.owner == r].sum() / w.sum()), 3) for r in range(world)],
               "nccl_transfers": int(sum(len(l) for l in dc.xplan)),
               "nccl_bytes": int(sum(int(dc.lay["usize"][m[0]]) for l in dc.xplan for m in l) * 8),
               "limitation": "fronts are not split across GPUs in round 1: the top log2(N) levels run on one GPU each"}
    del dc, F
    return out


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "klu", "chol"])
    ap.add_argument("--batch", type=int, default=4096, help="matrices per GPU")
    ap.add_argument("--chol-grid", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    # one process per GPU: stay on the cores (and the NUMA node) next to this rank's GPU before any pinned buffer exists
    numa_cores = 0
    if world > 1 and not os.environ.get("B200S_NO_NUMA_BIND"):
        from kvxopt_b200.dist import bind_to_gpu_numa
        numa_cores = bind_to_gpu_numa(local)
    import torch
    import torch.distributed as dist
    from kvxopt_b200 import _lib as L, klu
    fn = L.fn
    if L.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the numeric path has no CPU fallback")
    torch.cuda.set_device(local)
    assert fn["b200s_set_device"](local) == 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, hbm_src, fp64_peak = measured_peaks()
    A = load_activsg()
    n, nnz = A.shape[0], A.nnz
    Fs = klu.symbolic(A)
    Fn = klu.numeric(A, Fs)
    hn = klu._capsule_ptr(Fn, klu._NAME_NUM, "", "F")
    batch = args.batch
    host_vals = torch.empty((batch, nnz), dtype=torch.float64, pin_memory=True)
    perturbed_values(A.data, batch, rank, out=host_vals.numpy())
    dev_vals = host_vals.cuda()
    status = np.zeros(batch, dtype=np.int32)
    inf = L.KluInfo()
    # ---- warm-up
    for _ in range(args.warmup):
        assert fn["b200s_klu_refactor_batch_dev"](hn, dev_vals.data_ptr(), batch, nnz, None) == 0, L.last_error()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    # ---- timed: device-resident inputs.  Inputs (961 MB per step) are far larger than the 126 MB L2.
    dev_ms, ker_ms, dense_ms, nlaunch = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert fn["b200s_klu_refactor_batch_dev"](hn, dev_vals.data_ptr(), batch, nnz, None) == 0
        fn["b200s_klu_info"](hn, C.byref(inf))
        dev_ms += inf.ms_refactor
        ker_ms += inf.ms_kernel
        dense_ms += inf.ms_dense
        nlaunch += inf.launches
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else None
    dev_ms_max = max_over_ranks(dev_ms)
    wall_ms_max = max_over_ranks(wall_ms)
    value = world * batch * args.steps / (dev_ms_max * 1e-3)
    # ---- timed: end to end through the public API with host (pinned) buffers.  (i) one synchronous call per step
    # (upload, kernels, status download strictly one after the other); (ii) the streaming form of the same API --
    # refactor_batch_begin / refactor_batch_end with two batches in flight, so that the upload of step i+1 overlaps
    # the kernels of step i.  Every step's H2D (961 MB) and D2H (status) are inside the timed region in both.
    host_vals2 = torch.empty((batch, nnz), dtype=torch.float64, pin_memory=True)
    perturbed_values(A.data, batch, rank + 1000, out=host_vals2.numpy())
    hbufs = [host_vals.numpy(), host_vals2.numpy()]
    klu.refactor_batch(Fn, hbufs[0])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        st = klu.refactor_batch(Fn, hbufs[i % 2])
    barrier()
    e2e_sync_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    assert not st.any()
    klu.refactor_batch_begin(Fn, hbufs[0]); klu.refactor_batch_begin(Fn, hbufs[1])
    klu.refactor_batch_end(Fn); klu.refactor_batch_end(Fn)
    barrier()
    t0 = time.perf_counter()
    klu.refactor_batch_begin(Fn, hbufs[0])
    for i in range(1, args.steps):
        klu.refactor_batch_begin(Fn, hbufs[i % 2])
        st = klu.refactor_batch_end(Fn)
        assert not st.any()
    st = klu.refactor_batch_end(Fn)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    assert not st.any()
    e2e_value = world * batch * args.steps / (e2e_ms * 1e-3)
    e2e_sync_value = world * batch * args.steps / (e2e_sync_ms * 1e-3)
    klu.refactor_batch(Fn, hbufs[0])          # the factors the spot check below solves with
    d = klu.factor_info(Fn)
    # parity spot check of the timed configuration (one matrix of the batch against SuperLU)
    import scipy.sparse.linalg as spla
    Bb = np.random.default_rng(5).standard_normal((batch, 1, n))
    Xb = Bb.copy()
    klu.solve_batch(Fn, Xb)
    bsel = batch - 1
    Ab = sp.csc_matrix((host_vals.numpy()[bsel], A.indices, A.indptr), shape=(n, n))
    xref = spla.splu(Ab).solve(Bb[bsel, 0])
    spot = float(np.linalg.norm(Xb[bsel, 0] - xref) / np.linalg.norm(xref))

    bytes_per = d["bytes_per_refactor"]
    ker_avg_ms = ker_ms / args.steps
    achieved = bytes_per * batch / (ker_avg_ms * 1e-3) / 1e9
    line = {
        "metric": "batched KLU refactors/sec", "value": value, "unit": "refactors/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "klu_refactor_batch ACTIVSg2000 (n=4000, nnz=29336) same-pattern perturbations 1e-3",
                   "batch_per_gpu": batch, "global_batch": batch * world, "l2": "inputs_larger_than_l2 (961 MB values per GPU per step)",
                   "nnz_L": d["nnz_L"], "nnz_U": d["nnz_U"], "levels": d["nlevels"], "flops_per_refactor": d["flops"],
                   "parallelism": "independent matrices sharded by rank, no collective",
                   "host_cores_bound_per_rank": numa_cores},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "e2e": {"value": e2e_value, "unit": "refactors/s", "h2d_bytes_per_step": int(batch * nnz * 8),
                "d2h_bytes_per_step": int(batch * 4), "ms_per_step": e2e_ms / args.steps,
                "api": "kvxopt_b200.klu.refactor_batch_begin(Fn, values[batch, nnz]) / refactor_batch_end(Fn) -> status[batch], "
                       "two batches in flight: the pinned-host upload of step i+1 overlaps the kernels of step i; every "
                       "step's H2D and D2H are inside the timed region",
                "synchronous_call": {"value": e2e_sync_value, "ms_per_step": e2e_sync_ms / args.steps,
                                     "api": "kvxopt_b200.klu.refactor_batch(Fn, values) -> status, one blocking call per step"}},
        "gpu_launches": int(nlaunch),
        "roofline": {"bound": "hbm", "kernel": "k_klu_refactor_wave", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)",
                     "algorithmic_bytes_per_refactor": bytes_per, "kernel_ms_per_launch": ker_avg_ms,
                     "dense_block_ms_per_step": dense_ms / args.steps,
                     "traffic": traffic_from_profiles("k_klu_refactor_wave")},
        "parity_spot_check_rel_vs_superlu": spot,
        "clocks": clk,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = 1
        rate, _, _ = cpu_refactor_rate(A, cores, 48, 3, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "refactors/s", "cores": cores, "kind": "port",
                                "sample": "3 x 48 refactorizations on one core, oracle/klu_oracle.c (klu_refactor restatement, same ordering)"}
    if world == 1 and args.workload in ("all", "chol"):
        try:
            line["cholesky"] = bench_cholesky(args.chol_grid, max(2, min(args.steps, 3)), fp64_peak)
            if not args.no_cpu_baseline:
                line["cholesky"]["cpu_baseline"] = cpu_cholesky_sample(48)
        except Exception as e:  # the headline line must still be printed
            line["cholesky"] = {"error": repr(e)}
    if world > 1 and args.workload in ("all", "chol"):
        try:
            res = bench_cholesky_multi(args.chol_grid, 2, rank, world)
            if rank == 0:
                line["cholesky_subtree"] = res
        except Exception as e:
            if rank == 0:
                line["cholesky_subtree"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
