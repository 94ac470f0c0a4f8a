#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 sparse direct-solve hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload all|klu|chol|configs]

Headline workload (BASELINE.json configs[1]): batched KLU numeric refactorization of same-pattern value
perturbations of the ACTIVSg2000 power-flow Jacobian, 4096 matrices per GPU (weak scaling: every rank owns its
own 4096 matrices, no data-path collective).  metric = refactors/s.
  value : whole-job refactors/s with the value arrays already resident in HBM (b200s_klu_refactor_batch_dev),
          device-event time, max over ranks.
  e2e   : the same through the public API kvxopt_b200.klu.refactor_batch_begin/_end with HOST (pinned) buffers:
          H2D of the values and D2H of the per-matrix status inside the timed region.
At N>1 the line also carries `strong` (the fixed batch of 4096 of configs[1] split across the ranks) and
`cholesky_subtree` (configs[3] over N GPUs).  At N=1 it carries one object per remaining BASELINE config:
  config1  cholmod.linsolve on bcsstk24               config3  solvers.lp on boeing2 ('chol' KKT solver on the device)
  cholesky 100^3 Laplacian supernodal Cholesky        config5  200k-variable QP through coneqp, IPM iterations/s
each with its CPU baseline timed in the same run (cores stated).  configs 3 and 5 drive the UNMODIFIED reference
interior-point code (the caller, oracle/_ref probe build) with the B200 KKT solvers plugged in through the reference's
own kktsolver= API; they run in a child process because the reference's misc_solvers.scale is not safe under
multi-threaded OpenBLAS at m >= 4e5 rows (OPENBLAS_NUM_THREADS=1 must be set before the BLAS loads).

--impl reference times the CPU restatement of the reference's KLU path (oracle/klu_oracle.c) on all host cores of this
box, with an ordering computed WITHOUT the product library (SuperLU's minimum degree on A'+A through scipy, then the
oracle's own threshold-pivoting factorization), and reports both klu_factor (what reference src/C/klu.c:337 calls per
matrix) and klu_refactor rates; its `cholesky` object times the oracle's supernodal Cholesky (OpenBLAS, all cores) on
the 64^3 Laplacian that the GPU arm's `cholesky.sample_64` runs as well.  SuiteSparse itself is not installable here
(DESIGN.md section 4).  libb200sparse.so is never loaded by the reference arm.
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
CHOL_SAMPLE_GRID = 64          # the grid both arms factor so that a like-for-like Cholesky ratio exists


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n = int(z["n"])
    A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n))
    A.sort_indices()
    return A


def load_activsg():
    return load_golden("ACTIVSg2000")


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
    """SM clock and throttle reasons sampled DURING the timed regions, in-process through NVML every 5 ms (a
    `nvidia-smi -lms` child needs ~100 ms to deliver its first sample and saw nothing of a 150 ms region in round 1).
    start()/pause() bracket each timed region; samples taken outside are dropped."""

    def __init__(self, index):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz, self.err = None, None
        self._on, self._stop = False, False
        self.h = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:      # reported, never fatal
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        masks = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop:
            if self._on:
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    if self._on:
                        self.samples.append(float(mhz))
                        for name, m in masks.items():
                            if r & m:
                                self.reasons.add(name)
                except Exception as e:
                    self.err = repr(e)
            time.sleep(0.005)

    def start(self):
        self._on = True

    def pause(self):
        self._on = False

    def stop(self):
        self._on, self._stop = False, True
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "how": "NVML in-process, 5 ms period, timed regions only"}


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


# ------------------------------------------------------------------------------------------------------------
# CPU arms (oracle/, test infrastructure): nothing below this banner and above the next one touches libb200sparse.so
# ------------------------------------------------------------------------------------------------------------
_W = {}


def independent_klu_ordering(A):
    """fill-reducing ordering for the CPU arm computed without the product: SuperLU's multiple-minimum-degree ordering of
    A'+A (scipy.sparse.linalg.splu, permc_spec='MMD_AT_PLUS_A') -- the same class as KLU's AMD on A+A'; applied
    symmetrically (P0 = Q), the oracle's factorization then does its own threshold partial pivoting (tol 1e-3)"""
    import scipy.sparse.linalg as spla
    lu = spla.splu(A.tocsc(), permc_spec="MMD_AT_PLUS_A")
    return np.argsort(lu.perm_c).astype(np.int64)


def _worker_init(cp, ri, vx, q):
    from oracle import KluOracle
    _W["args"] = (len(cp) - 1, cp, ri)
    _W["q"] = q
    _W["o"] = KluOracle(len(cp) - 1, cp, ri, vx, P0=q, Q=q)
    _W["base"] = vx


def _worker_run(args):
    from oracle import KluOracle
    seed, count, mode = args
    o, base = _W["o"], _W["base"]
    rng = np.random.default_rng(seed)
    vals = [base * (1 + 1e-3 * rng.uniform(-1, 1, base.size)) for _ in range(count)]
    t0 = time.perf_counter()
    if mode == "refactor":
        for v in vals:
            o.refactor(v)
    else:       # klu_l_factor: pattern discovery + pivot search per matrix, what klu.numeric (klu.c:337) does
        n, cp, ri = _W["args"]
        for v in vals:
            KluOracle(n, cp, ri, v, P0=_W["q"], Q=_W["q"])
    return time.perf_counter() - t0


def cpu_klu_rates(A, cores, per_worker, steps, warmup):
    import multiprocessing as mp
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    q = independent_klu_ordering(A)
    from oracle import KluOracle
    o = KluOracle(A.shape[0], cp, ri, vx, P0=q, Q=q)
    stats = {"nnz_L": o.nnz_L, "nnz_U": o.nnz_U, "flops": o.flops}
    ctx = mp.get_context("fork")
    res = {}
    with ctx.Pool(cores, initializer=_worker_init, initargs=(cp, ri, vx, q)) as pool:
        for mode, count in (("refactor", per_worker), ("factor", max(per_worker // 4, 2))):
            for w in range(warmup):
                pool.map(_worker_run, [(1000 + w * cores + c, 2, mode) for c in range(cores)])
            times = []
            for s in range(steps):
                t0 = time.perf_counter()
                pool.map(_worker_run, [(s * cores + c, count, mode) for c in range(cores)])
                times.append(time.perf_counter() - t0)
            total = float(np.sum(times))
            res[mode] = (cores * count * steps / total, total / steps * 1e3, cores * count)
    return res, stats


def cpu_cholesky_sample(nx, reps):
    """the oracle's supernodal left-looking LL^T (OpenBLAS, all cores) on the nx^3 Laplacian with a nested-dissection
    ordering computed in oracle/grid_nd.py (independent of the product)"""
    from oracle import CholOracle, lib as olib
    from oracle.grid_nd import grid_nd_perm
    Al = lap3d_lower(nx)
    n = Al.shape[0]
    perm = grid_nd_perm(nx, nx, nx, 64)
    threads = olib().oracle_blas_threads(os.cpu_count() or 1)
    t0 = time.perf_counter(); O = CholOracle(n, Al.indptr, Al.indices, "L", perm); ta = time.perf_counter() - t0
    b = np.random.default_rng(0).standard_normal(n)
    tf, ts = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); O.factorize(Al.data); tf.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); x = O.solve(b); ts.append(time.perf_counter() - t0)
    A = (Al + sp.tril(Al, -1).T).tocsr()
    berr = float(np.linalg.norm(A @ x - b) / (12.0 * np.linalg.norm(x) + np.linalg.norm(b)))
    return {"workload": "7-point Laplacian %d^3, nested dissection (oracle/grid_nd.py), supernodal LL^T, 1 RHS" % nx,
            "n": n, "nnz_L": O.nnzL, "flops": O.flops, "analyze_ms_host": ta * 1e3, "factor_ms": min(tf) * 1e3, "solve_ms": min(ts) * 1e3,
            "factor_plus_solve_ms": (min(tf) + min(ts)) * 1e3, "factor_tflops": O.flops / min(tf) / 1e12, "backward_error": berr,
            "cores": threads, "kind": "port", "reps": reps,
            "sample": "oracle/chol_oracle.c (left-looking supernodal LL^T, OpenBLAS dsyrk/dgemm/dpotrf/dtrsm on %d threads)" % threads}


def run_reference(args, rank, world):
    if rank != 0:
        return
    assert "kvxopt_b200" not in sys.modules
    A = load_activsg()
    cores = os.cpu_count() or 1
    per_worker = 24
    line = {"impl": "reference", "metric": "batched KLU refactors/sec", "unit": "refactors/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gpu_launches": 0}
    if args.workload in ("all", "klu", "configs"):
        res, stats = cpu_klu_rates(A, cores, per_worker, args.steps, max(args.warmup, 1))
        rate, ms_step, per_step = res["refactor"]
        frate, fms, fper = res["factor"]
        line.update({
            "value": rate, "ms_per_step": ms_step,
            "config": {"workload": "klu_refactor_batch ACTIVSg2000 (n=4000, nnz=29336) same-pattern perturbations 1e-3",
                       "batch_per_step": per_step,
                       "ordering": "independent of the product: SuperLU MMD on A'+A (scipy), oracle's own threshold pivoting; "
                                   "nnz(L)=%d nnz(U)=%d flops=%.0f" % (stats["nnz_L"], stats["nnz_U"], stats["flops"])},
            "cpu_baseline": {"value": rate, "unit": "refactors/s", "cores": cores, "kind": "port",
                             "sample": "%d refactorizations per step (%d per core), oracle/klu_oracle.c klu_refactor restatement; "
                                       "SuiteSparse KLU is not installable in this image" % (per_step, per_worker)},
            "klu_factor": {"value": frate, "unit": "factorizations/s", "ms_per_step": fms, "batch_per_step": fper,
                           "what": "klu_l_factor restatement (pattern discovery + threshold pivot search per matrix): what the reference's "
                                   "klu.numeric calls for every matrix (src/C/klu.c:337); the reference never calls klu_refactor, so "
                                   "`value` (refactor, no pivot search) is the KINDER baseline for the GPU arm"},
            "e2e": {"value": rate, "unit": "refactors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    if args.workload in ("all", "chol"):
        try:
            line["cholesky"] = cpu_cholesky_sample(args.chol_sample_grid, max(1, min(args.steps, 2)))
        except Exception as e:
            line["cholesky"] = {"error": repr(e)}
        if "value" not in line:
            c = line["cholesky"]
            line.update({"metric": "sparse Cholesky factor+solve ms", "unit": "ms", "higher_is_better": False,
                         "value": c.get("factor_plus_solve_ms"), "ms_per_step": c.get("factor_plus_solve_ms"), "config": {"workload": c.get("workload")},
                         "cpu_baseline": {"value": c.get("factor_plus_solve_ms"), "unit": "ms", "cores": c.get("cores"), "kind": "port", "sample": c.get("sample")},
                         "e2e": {"value": c.get("factor_plus_solve_ms"), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    line["native_so_loaded"] = sorted({os.path.relpath(l.split()[-1], ROOT) for l in open("/proc/self/maps") if ROOT in l and ".so" in l})
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def bench_cholesky(nx, steps, fp64_peak, full=True):
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
    # the same solve with the launch-per-step sweeps (what the persistent level kernels replace): time and bitwise comparison
    cholmod.set_solve_sweeps(F, 0)
    ms_s0 = []
    for _ in range(max(2, steps)):
        Xd0 = Bd.clone()
        torch.cuda.synchronize()
        assert fn["b200s_chol_solve_dev"](h, 0, Xd0.data_ptr(), 1, n) == 0
        fn["b200s_chol_info"](h, C.byref(inf))
        ms_s0.append(inf.ms_solve)
    same_bits = bool(np.array_equal(x, Xd0.cpu().numpy()))
    cholmod.set_solve_sweeps(F, -1)
    A = (Al + sp.tril(Al, -1).T).tocsr()
    berr = float(np.linalg.norm(A @ x - B[:, 0]) / (12.0 * np.linalg.norm(x) + np.linalg.norm(B)))
    # end to end through the public API with host buffers (H2D of values and RHS, D2H of the solution)
    # (one untimed pass first: the host-buffer entry points stage through their own buffers and record their own solve graph,
    #  and the sweep-mode switch above rebuilt the solve schedules; then the best of two timed passes, tools/prof_chol_e2e.py)
    e2e_all = []
    for _ in range(3):
        Xh = np.asfortranarray(B.copy())
        t0 = time.perf_counter()
        cholmod.numeric(Al, F)
        cholmod.solve(F, Xh)
        e2e_all.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = float(min(e2e_all[1:]))
    d = cholmod.factor_info(F)
    best_f, best_s = float(np.min(ms_f)), float(np.min(ms_s))
    out = {
        "workload": "7-point Laplacian %d^3, geometric nested dissection (leaf 64), supernodal LL^T, 1 RHS" % nx,
        "n": n, "nnz_L": d["nnz_L"], "nsuper": d["nsuper"], "levels": d["nlevels"], "max_front": [d["max_front_rows"], d["max_front_cols"]],
        "flops": d["flops"], "analyze_ms_host": t_analyze * 1e3,
        "factor_ms": best_f, "solve_ms": best_s, "factor_plus_solve_ms": best_f + best_s,
        "factor_tflops": d["flops"] / (best_f * 1e-3) / 1e12,
        "solve_gbs": (16.0 * d["nnz_L"] + 16.0 * n) / (best_s * 1e-3) / 1e9,
        "e2e_factor_plus_solve_ms_host_buffers": e2e_ms, "e2e_first_pass_ms_host_buffers": float(e2e_all[0]), "backward_error": berr,
        "solve_ms_launch_per_step": float(np.min(ms_s0)), "solve_bitwise_equal_both_paths": same_bits,
    }
    hbm, hbm_src, _ = measured_peaks()
    out["solve_roofline"] = {"bound": "hbm", "kernel": "k_fwd_persist / k_bwd_persist / k_bwd_rect + the launch-per-step sweeps of the wide levels",
                             "achieved": out["solve_gbs"], "peak": hbm, "peak_source": hbm_src, "unit": "GB/s", "frac": out["solve_gbs"] / hbm,
                             "algorithmic_bytes": 16.0 * d["nnz_L"] + 16.0 * n,
                             "note": "L is read once per sweep (2 x 8 x nnz(L)); the chain of the root front (117 dependent block steps) and one "
                                     "64 KB slice in flight per SM in the streaming levels bound it, DESIGN.md 2.4"}
    if not full:
        del F
        return out
    # one profiled factorization: per-kernel-class device time (events around every launch)
    fn["b200s_chol_set_profiling"](h, 1)
    assert fn["b200s_chol_factorize_dev"](h, vals_dev.data_ptr(), C.byref(minor)) == 0
    fn["b200s_chol_info"](h, C.byref(inf))
    fn["b200s_chol_set_profiling"](h, 0)
    d = inf.asdict()
    upd_tf = d["flops_update"] / (d["ms_dense_update"] * 1e-3) / 1e12 if d["ms_dense_update"] > 0 else None
    out["kernel_ms_profiled"] = {"extend_add": d["ms_extend"], "small_fronts": d["ms_potrf"], "panel": d["ms_trsm"],
                                 "dmma_update": d["ms_dense_update"]}
    out["roofline"] = {"bound": "tensor", "kernel": "k_update (FP64 DMMA m8n8k4)", "achieved": upd_tf, "peak": fp64_peak,
                       "unit": "TFLOP/s", "frac": (upd_tf / fp64_peak) if upd_tf else None,
                       "peak_source": "builder-measured: tools/fp64_peak.cu (DMMA m8n8k4 and DFMA issue-rate microbenchmark, "
                                      "profiles/r01_fp64_peak.json); MEASURED_PEAKS.json has no FP64 entry",
                       "traffic": traffic_from_profiles("k_update")}
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


def bench_cholesky_multi(nx, steps, rank, world):
    """BASELINE configs[3] over N GPUs: subtree-to-subcube (kvxopt_b200/dist.py), update matrices over NCCL/NVLink"""
    import torch
    import torch.distributed as dist
    from kvxopt_b200 import _lib as L, cholmod, dist as D
    Al = lap3d_lower(nx)
    n = Al.shape[0]
    perm = np.zeros(n, np.int64)
    L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
    # top-level separators are kept as fronts of their own (a merged front has no Schur complement left to share)
    cholmod.engine_options["max_merge_cols"] = D.DIST_MAX_MERGE_COLS
    F = cholmod.symbolic(Al, p=perm)
    cholmod.engine_options.pop("max_merge_cols", None)
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
    # distributed triangular solves (panels stay on their GPUs): timed like the factorization
    bvec = torch.from_numpy(np.random.default_rng(0).standard_normal(n)).cuda()
    stimes = []
    for rep in range(steps + 2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        xdist = dc.solve(bvec)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        if rep > 0:
            stimes.append((time.perf_counter() - t0) * 1e3)
    out = None
    if rank == 0:
        A_ = (Al + sp.tril(Al, -1).T).tocsr()
        xd = xdist.cpu().numpy(); bh = bvec.cpu().numpy()
        berr_dist = float(np.linalg.norm(A_ @ xd - bh) / (12.0 * np.linalg.norm(xd) + np.linalg.norm(bh)))
        b = np.random.default_rng(0).standard_normal((n, 1)); x = np.asfortranarray(b.copy())
        cholmod.solve(F, x)
        d = cholmod.factor_info(F)
        A = (Al + sp.tril(Al, -1).T).tocsr()
        berr = float(np.linalg.norm(A @ x - b) / (12.0 * np.linalg.norm(x) + np.linalg.norm(b)))
        w = D.front_work(dc.lay)
        fm, gm = float(np.min(times)), float(np.min(gtimes))
        out = {"workload": "7-point Laplacian %d^3, nested dissection, subtree-to-subcube over %d GPUs" % (nx, world),
               "factor_ms": fm, "gather_panels_ms": gm, "solve_ms_rank0": d["ms_solve"],
               "factor_gather_solve_ms": fm + gm + d["ms_solve"],
               "solve_distributed_ms": float(np.min(stimes)), "factor_plus_distributed_solve_ms": fm + float(np.min(stimes)),
               "backward_error_distributed_solve": berr_dist,
               "factor_tflops": d["flops"] / (fm * 1e-3) / 1e12,
               "backward_error": berr, "timing": "host clock between barrier+synchronize pairs, max over ranks by construction",
               "work_share_per_rank": [round(float(D._work_share(dc.lay, dc.owner, dc.splan, r) / w.sum()), 3) for r in range(world)],
               "work_share_per_rank_fronts_only": [round(float(w[dc.owner == r].sum() / w.sum()), 3) for r in range(world)],
               "shared_fronts": {int(s): [[int(r), int(lo), int(hi)] for r, lo, hi in parts] for s, parts in dc.splan.items()},
               "nccl_transfers": int(sum(len(l) for l in dc.xplan) + sum(len(l) for l in dc.panel_moves) + sum(len(l) for l in dc.slab_moves)),
               "nccl_bytes": int((sum(int(dc.lay["usize"][m[0]]) for l in dc.xplan for m in l)
                                  + sum(int(dc.lay["lsize"][m[0]]) for l in dc.panel_moves for m in l)
                                  + sum(D.slab_range(dc.lay, m[0], m[3], m[4])[1] for l in dc.slab_moves for m in l)) * 8),
               "limitation": "the Schur complements of the top separators are shared by the ranks of their subtree group; their panels "
                             "(potrf + trsm + in-panel updates) and the root front still run on one GPU each; solve_distributed_ms = triangular solves with the panels where "
                             "they were factored (update vectors up the cut edges, solution entries down the subtree groups); "
                             "solve_ms_rank0 = the single-GPU solve after gathering all panels on rank 0"}
    del dc, F
    return out


def bench_config1(steps):
    """BASELINE configs[0]: cholmod.linsolve on bcsstk24 (lower triangle as stored), random RHS, through the public API with
    host buffers; CPU baseline = the oracle with its own (SuperLU MMD on A+A') ordering, timed in the same run."""
    from kvxopt_b200 import cholmod
    Al = sp.tril(load_golden("bcsstk24")).tocsc(); Al.sort_indices()
    A = (Al + sp.tril(Al, -1).T).tocsc()
    n = A.shape[0]
    B = np.random.default_rng(0).standard_normal((n, 1))

    def best(f, reps):
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); r = f(); ts.append((time.perf_counter() - t) * 1e3)
        return min(ts), r

    reps = max(8, steps)
    # headline = COLD calls: the engine keeps the symbolic analysis of the last few patterns (kvxopt re-analyses the same pattern
    # in every linsolve call and every IPM iteration); a benchmark that repeats one matrix would only time that cache, so the
    # pattern is changed between repetitions (one explicit zero moved), and the warm (cache-hit) time is reported next to it
    cholmod.linsolve(Al, np.asfortranarray(B.copy()))

    def cold_variant(k):
        M = Al.tolil(copy=True)
        M[n - 1, k] = 1e-300 if M[n - 1, k] == 0 else M[n - 1, k]
        M = M.tocsc(); M.sort_indices()
        return M
    variants = [cold_variant(k) for k in range(1, reps + 1)]
    ts = []
    for M in variants:
        Xc = np.asfortranarray(B.copy())
        t = time.perf_counter(); cholmod.linsolve(M, Xc); ts.append((time.perf_counter() - t) * 1e3)
    ms_lin = min(ts)
    ms_warm, _ = best(lambda: cholmod.linsolve(Al, np.asfortranarray(B.copy())), reps)
    ts = []
    for M in [cold_variant(k) for k in range(reps + 1, 2 * reps + 1)]:
        t = time.perf_counter(); F = cholmod.symbolic(M); ts.append((time.perf_counter() - t) * 1e3)
    ms_sym = min(ts)
    F = cholmod.symbolic(Al)
    ms_num, _ = best(lambda: cholmod.numeric(Al, F), reps)
    X = np.asfortranarray(B.copy())
    ms_sol, _ = best(lambda: cholmod.solve(F, X), reps)
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    anorm = abs(A).sum(axis=0).max()
    berr = float(np.linalg.norm(A @ X - B) / (anorm * np.linalg.norm(X) + np.linalg.norm(B)))
    info = cholmod.factor_info(F)
    # CPU arm, independent ordering
    import scipy.sparse.linalg as spla
    from oracle import CholOracle, lib as olib
    threads = olib().oracle_blas_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    pat = sp.csc_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape) + sp.identity(n) * (2.0 * A.getnnz(axis=0).max())
    perm = np.argsort(spla.splu(pat.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0).perm_c).astype(np.int64)
    ms_ord = (time.perf_counter() - t0) * 1e3
    ms_oan, O = best(lambda: CholOracle(n, Al.indptr, Al.indices, "L", perm), 2)
    ms_ofa, _ = best(lambda: O.factorize(Al.data), 3)
    ms_oso, Xo = best(lambda: O.solve(B), 3)
    return {"workload": "cholmod.linsolve(A, B) on bcsstk24 (n=3562, lower triangle as stored, 81736 entries), B = n x 1 normal seed 0",
            "n": n, "nnz_L": info["nnz_L"], "flops": info["flops"],
            "linsolve_ms": ms_lin, "linsolve_ms_warm_pattern_cache": ms_warm,
            "split_ms": {"symbolic_host (cold)": ms_sym, "numeric": ms_num, "solve": ms_sol},
            "e2e": {"value": ms_lin, "unit": "ms", "h2d_bytes_per_step": int(Al.nnz * 8 + n * 8), "d2h_bytes_per_step": int(n * 8),
                    "api": "kvxopt_b200.cholmod.linsolve (host buffers in, solution out)"},
            "backward_error": berr, "rel_diff_vs_oracle": float(np.linalg.norm(X - Xo) / np.linalg.norm(Xo)),
            "roofline": {"bound": "latency", "note": "3.2e7 flops and 4.5 MB: launch/transfer latency and the host symbolic analysis dominate",
                         "numeric_gflops": info["flops"] / (ms_num * 1e-3) / 1e9},
            "cpu_baseline": {"value": ms_oan + ms_ofa + ms_oso, "unit": "ms", "cores": threads, "kind": "port",
                             "split_ms": {"ordering (SuperLU MMD via scipy, not counted)": ms_ord, "symbolic": ms_oan, "numeric": ms_ofa, "solve": ms_oso},
                             "sample": "whole workload: oracle/chol_oracle.c analyze + factorize + solve, own ordering"}}


def bench_strong(fn, L, hn, dev_vals, batch, nnz, world, steps, barrier, max_over_ranks):
    """configs[1] as written: ONE batch of `batch` matrices split across the ranks (strong scaling)"""
    per = batch // world
    inf = L.KluInfo()
    for _ in range(3):
        assert fn["b200s_klu_refactor_batch_dev"](hn, dev_vals.data_ptr(), per, nnz, None) == 0
    barrier()
    ms = 0.0
    for _ in range(steps):
        assert fn["b200s_klu_refactor_batch_dev"](hn, dev_vals.data_ptr(), per, nnz, None) == 0
        fn["b200s_klu_info"](hn, C.byref(inf))
        ms += inf.ms_refactor
    barrier()
    ms = max_over_ranks(ms)
    return {"scaling": "strong", "global_batch": per * world, "batch_per_gpu": per, "value": per * world * steps / (ms * 1e-3),
            "unit": "refactors/s", "ms_per_step": ms / steps,
            "note": "one CTA serves 32 matrices and its time is the latency of the column-dependency chain, so a smaller per-GPU batch "
                    "runs fewer CTAs for about the same time (DESIGN.md section 3.2)"}


def bench_config2_single(steps):
    """BASELINE configs[1], first part: ONE ACTIVSg2000 Jacobian through klu.linsolve / symbolic / numeric / solve with host
    buffers; CPU arm = oracle/klu_oracle.c (its own BTF-free ordering: SuperLU MMD, own pivoting), timed in the same run."""
    from kvxopt_b200 import klu
    from oracle import KluOracle
    import scipy.sparse.linalg as spla
    A = load_activsg()
    n = A.shape[0]
    B = np.asfortranarray(np.random.default_rng(0).standard_normal((n, 1)))
    reps = max(5, steps)

    def best(f):
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); r = f(); ts.append((time.perf_counter() - t) * 1e3)
        return min(ts), r
    klu.linsolve(A, B.copy(order="F"))
    ms_lin, _ = best(lambda: klu.linsolve(A, B.copy(order="F")))
    ms_sym, Fs = best(lambda: klu.symbolic(A))
    ms_num, Fn = best(lambda: klu.numeric(A, Fs))
    X = B.copy(order="F")
    ms_sol, _ = best(lambda: klu.solve(A, Fs, Fn, X))
    X = B.copy(order="F"); klu.solve(A, Fs, Fn, X)
    res = float(np.abs(A @ X - B).max())
    xs = spla.splu(A.tocsc()).solve(B[:, 0])
    t0 = time.perf_counter()
    Q = independent_klu_ordering(A)
    ms_ord = (time.perf_counter() - t0) * 1e3
    ms_ofac, O = best(lambda: KluOracle(n, A.indptr, A.indices, A.data, P0=Q, Q=Q))
    ms_osol, xo = best(lambda: O.solve(B[:, 0]))
    return {"workload": "klu.linsolve(A, b) on ONE ACTIVSg2000 Jacobian (n=4000, nnz=29336), host buffers in, solution out",
            "linsolve_ms": ms_lin,
            "split_ms": {"symbolic_host (BTF + AMD)": ms_sym, "numeric (host pivot search, its values loaded to the device)": ms_num,
                         "solve (one matrix: k_klu_solve_one, one CTA per right-hand side; repeat calls timed)": ms_sol},
            "e2e": {"value": ms_lin, "unit": "ms", "h2d_bytes_per_step": int(8 * (A.nnz + 2 * n)), "d2h_bytes_per_step": int(8 * n),
                    "api": "kvxopt_b200.klu.linsolve"},
            "max_residual": res, "rel_diff_vs_superlu": float(np.linalg.norm(X[:, 0] - xs) / np.linalg.norm(xs)),
            "roofline": {"bound": "latency", "note": "1.9e6 flops: the host analysis (2.3 ms) and pivot search (2.8 ms) and the ~8200 dependent column operations of one solve (one warp, work vector in shared memory) bound it; "
                                                     "the engine is built for the batched case (headline)"},
            "cpu_baseline": {"value": ms_ofac + ms_osol, "unit": "ms", "cores": 1, "kind": "port",
                             "split_ms": {"ordering (SuperLU MMD via scipy, not counted)": ms_ord, "factor": ms_ofac, "solve": ms_osol},
                             "rel_diff_vs_superlu": float(np.linalg.norm(xo - xs) / np.linalg.norm(xs)),
                             "sample": "whole workload: oracle/klu_oracle.c pivoting factorization + solve, own ordering"}}


def ipm_child():
    """configs 3 and 5 (run as `bench.py --ipm-child`, OPENBLAS_NUM_THREADS=1 in the environment): the unmodified reference
    IPM (oracle/_ref, the CALLER of the path) with the B200 KKT solvers plugged in through kktsolver=; prints one JSON line"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(0, GOLD)
    out = {}
    import kvxopt
    from kvxopt_b200 import cholmod as gcholmod, klu as gklu, kkt as gkkt, _lib as L
    gcholmod.install(kvxopt); gklu.install(kvxopt)
    from kvxopt import matrix, spmatrix, solvers
    solvers.options["show_progress"] = False
    gcholmod.linsolve(sp.identity(2, format="csc"), np.ones((2, 1), order="F"))      # CUDA context before any timed region
    cores = os.cpu_count() or 1

    def tosp(M):
        M = sp.coo_matrix(M)
        return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)

    # ---- config 3: solvers.lp on boeing2 through conelp
    try:
        z = np.load(os.path.join(GOLD, "boeing2_lp.npz"))
        G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
        Aeq = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
        c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
        Gd, Ad, Gs, As = matrix(G.toarray()), matrix(Aeq.toarray()), tosp(G), tosp(Aeq)
        dims = {"l": G.shape[0], "q": [], "s": []}

        def run(Gm, Am, mk):
            best, sol = 1e30, None
            for _ in range(4):
                k = mk()
                t = time.perf_counter(); sol = solvers.conelp(c, Gm, h, dims, Am, b, **({} if k is None else {"kktsolver": k}))
                best = min(best, (time.perf_counter() - t) * 1e3)
            return {"ms": best, "iterations": sol["iterations"], "objective": sol["primal objective"], "status": sol["status"]}

        arms = {"device_chol (kvxopt_b200.kkt.chol, dense 'chol' counterpart, misc.py:1213-1349)": run(Gd, Ad, lambda: gkkt.chol(Gd, dims, Ad)),
                "device_chol2 (kvxopt_b200.kkt.chol2, device-resident sparse reduced system)": run(Gs, As, lambda: gkkt.chol2(Gs, dims, As)),
                "module_chol2 (reference misc.kkt_chol2 on kvxopt_b200.cholmod)": run(Gs, As, lambda: None)}
        ref = {"reference_chol (LAPACK, dense)": run(Gd, Ad, lambda: "chol"), "reference_chol2 (LAPACK, dense G)": run(Gd, Ad, lambda: "chol2")}
        first = arms["device_chol (kvxopt_b200.kkt.chol, dense 'chol' counterpart, misc.py:1213-1349)"]
        rc = ref["reference_chol (LAPACK, dense)"]
        out["config3"] = {"workload": "solvers.lp on boeing2 (n=143, G 352x143, A 4x143) via conelp, full IPM, KKT factorization on the GPU each iteration",
                          "ms": first["ms"], "iterations": first["iterations"], "objective": first["objective"], "arms": arms,
                          "parity": {"iterations_equal_reference": first["iterations"] == rc["iterations"],
                                     "objective_rel_diff": abs(first["objective"] - rc["objective"]) / abs(rc["objective"])},
                          "e2e": {"value": first["ms"], "unit": "ms", "api": "solvers.conelp(..., kktsolver=kvxopt_b200.kkt.chol(G, dims, A)): host vectors in/out every KKT solve"},
                          "roofline": {"bound": "latency", "note": "143 x 143 dense system: 30 factorizations + ~90 solves of microseconds of work each; "
                                                                    "per-call launch + PCIe round trips bound it"},
                          "cpu_baseline": {"value": rc["ms"], "unit": "ms", "cores": 1, "kind": "reference",
                                           "sample": "whole workload: the reference's own 'chol' kktsolver (LAPACK via oracle/_ref, OPENBLAS_NUM_THREADS=1)",
                                           "arms": ref}}
    except Exception as e:
        out["config3"] = {"error": repr(e)}

    # ---- config 5: 200k-variable QP through coneqp
    try:
        from generators import qp_instance
        nx, ny, nrand = 500, 400, 5000
        P, q, G, h = qp_instance(nx, ny, nrand)
        Pk, Gk = tosp(sp.tril(P)), tosp(G)
        t0 = time.perf_counter()
        ks = gkkt.qp_kktsolver(Pk, Gk)
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=ks)
        wall = time.perf_counter() - t0
        inf = ks.info()
        _, _, fp64_peak = measured_peaks()
        tf = inf["flops"] / (inf["ms_factor"] * 1e-3) / 1e12 if inf.get("ms_factor") else None
        c5 = {"workload": "synthetic sparse QP n=200000 (500x400 grid Laplacian + 1e-2 I), m=405000 (box + 5000 random 3-nnz rows), coneqp",
              "iterations": sol["iterations"], "objective": sol["primal objective"], "status": sol["status"],
              "ipm_iterations_per_s": sol["iterations"] / wall, "wall_s": wall, "setup_s_host_analysis": t_setup,
              "per_iteration_ms": {"assemble": inf.get("ms_assemble"), "factor": inf.get("ms_factor"), "kkt_solve": inf.get("ms_solve")},
              "nnz_L": inf.get("nnz_L"), "flops_per_factorization": inf.get("flops"),
              "e2e": {"value": sol["iterations"] / wall, "unit": "IPM iterations/s",
                      "api": "solvers.qp(P, q, G, h, kktsolver=kvxopt_b200.kkt.qp_kktsolver(P, G)): W['di'] and x, y, z cross PCIe every call"},
              "roofline": {"bound": "tensor", "kernel": "k_update (FP64 DMMA) inside the factorization of S", "achieved": tf, "peak": fp64_peak,
                           "unit": "TFLOP/s", "frac": tf / fp64_peak if tf else None,
                           "note": "whole-factorization rate (assembly, small fronts and panels included), last iteration"},
              "parity": {"golden_objective": -43988.5773845128, "golden_iterations": 10,
                         "objective_rel_diff": abs(sol["primal objective"] + 43988.5773845128) / 43988.5773845128}}
        # CPU sample: 2 IPM iterations of the reference's own kkt_chol2 with the oracle behind kvxopt.cholmod
        try:
            from oracle import cholmod_cpu, lib as olib
            threads = os.cpu_count() or 1
            on = cholmod_cpu.numeric

            def numeric_mt(*a, **k):       # BLAS threads only inside the factorization (scale() is not thread-safe here)
                olib().oracle_blas_threads(threads)
                try:
                    return on(*a, **k)
                finally:
                    olib().oracle_blas_threads(1)
            cholmod_cpu.numeric = numeric_mt
            sys.modules["kvxopt.cholmod"] = cholmod_cpu; kvxopt.cholmod = cholmod_cpu
            import kvxopt.misc as misc
            misc.cholmod = cholmod_cpu
            solvers.options["maxiters"] = 2
            t0 = time.perf_counter()
            s2 = solvers.qp(Pk, matrix(q), Gk, matrix(h))
            w2 = time.perf_counter() - t0
            c5["cpu_baseline"] = {"value": 2.0 / w2, "unit": "IPM iterations/s", "cores": threads, "kind": "port",
                                  "sample": "2 iterations (maxiters=2; %.1f s incl. the initial KKT solve) of the reference's coneqp + misc.kkt_chol2 "
                                            "with oracle/chol_oracle.c behind kvxopt.cholmod (fill-reducing ordering from the engine's host AMD -- "
                                            "integer work, no GPU --, OpenBLAS %d threads inside the factorization)" % (w2, threads)}
        except Exception as e:
            c5["cpu_baseline"] = {"error": repr(e)}
        out["config5"] = c5
    except Exception as e:
        out["config5"] = {"error": repr(e)}
    print("IPMCHILD " + json.dumps(out), flush=True)


def run_ipm_child(timeout=600):
    env = dict(os.environ)
    env["OPENBLAS_NUM_THREADS"] = "1"
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--ipm-child"], env=env, capture_output=True, text=True, timeout=timeout)
        for ln in reversed(p.stdout.splitlines()):
            if ln.startswith("IPMCHILD "):
                return json.loads(ln[len("IPMCHILD "):])
        return {"config3": {"error": "child rc=%d: %s" % (p.returncode, p.stderr[-400:])}, "config5": {"error": "child failed"}}
    except Exception as e:
        return {"config3": {"error": repr(e)}, "config5": {"error": repr(e)}}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "klu", "chol", "configs"])
    ap.add_argument("--batch", type=int, default=4096, help="matrices per GPU")
    ap.add_argument("--chol-grid", type=int, default=100)
    ap.add_argument("--chol-sample-grid", type=int, default=CHOL_SAMPLE_GRID)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ipm-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.ipm_child:
        ipm_child()
        return
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
    inf = L.KluInfo()
    clocks = ClockSampler(local) if rank == 0 else None
    # ---- warm-up
    for _ in range(args.warmup):
        assert fn["b200s_klu_refactor_batch_dev"](hn, dev_vals.data_ptr(), batch, nnz, None) == 0, L.last_error()
    barrier()
    if clocks:
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
    if clocks:
        clocks.pause()
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
    if clocks:
        clocks.start()
    t0 = time.perf_counter()
    klu.refactor_batch_begin(Fn, hbufs[0])
    for i in range(1, args.steps):
        klu.refactor_batch_begin(Fn, hbufs[i % 2])
        st = klu.refactor_batch_end(Fn)
        assert not st.any()
    st = klu.refactor_batch_end(Fn)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    clk = clocks.stop() if clocks else None
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
    del Bb, Xb

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
                "h2d_gbs": batch * nnz * 8 / (e2e_ms / args.steps * 1e-3) / 1e9,
                "api": "kvxopt_b200.klu.refactor_batch_begin(Fn, values[batch, nnz]) / refactor_batch_end(Fn) -> status[batch], "
                       "two batches in flight: the pinned-host upload of step i+1 overlaps the kernels of step i; every "
                       "step's H2D and D2H are inside the timed region",
                "synchronous_call": {"value": e2e_sync_value, "ms_per_step": e2e_sync_ms / args.steps,
                                     "api": "kvxopt_b200.klu.refactor_batch(Fn, values) -> status, one blocking call per step"}},
        "gpu_launches": int(nlaunch),
        "roofline": {"bound": "hbm", "kernel": "k_klu_early (one launch per wide dependency level) + k_klu_refactor_wave", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)",
                     "algorithmic_bytes_per_refactor": bytes_per, "kernel_ms_per_launch": ker_avg_ms,
                     "dense_block_ms_per_step": dense_ms / args.steps,
                     "traffic": (traffic_from_profiles("k_klu_refactor_wave") or 0) + (traffic_from_profiles("k_klu_early") or 0) or None,
                     "traffic_source": "profiles/r02c_ncu_klu_wave_early.txt: dram bytes of the wave kernel (8.72 GB) + the six early launches (1.64 GB), one batch of 4096"},
        "parity_spot_check_rel_vs_superlu": spot,
        "clocks": clk,
    }
    if world > 1 and batch % world == 0:
        try:
            line["strong"] = bench_strong(fn, L, hn, dev_vals, batch, nnz, world, args.steps, barrier, max_over_ranks)
        except Exception as e:
            line["strong"] = {"error": repr(e)}
    del dev_vals, host_vals, host_vals2, hbufs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res, stats = cpu_klu_rates(A, 1, 48, 3, 1)
        line["cpu_baseline"] = {"value": res["refactor"][0], "unit": "refactors/s", "cores": 1, "kind": "port",
                                "klu_factor_per_s": res["factor"][0],
                                "sample": "3 x 48 refactorizations on one core, oracle/klu_oracle.c (klu_refactor restatement; own ordering: "
                                          "SuperLU MMD on A'+A, own pivoting; nnz(L)=%d)" % stats["nnz_L"]}
    if world == 1 and args.workload in ("all", "chol"):
        try:
            line["cholesky"] = bench_cholesky(args.chol_grid, max(2, min(args.steps, 3)), fp64_peak)
            line["cholesky"]["sample_64"] = bench_cholesky(args.chol_sample_grid, 3, fp64_peak, full=False)
            if not args.no_cpu_baseline:
                line["cholesky"]["cpu_baseline"] = cpu_cholesky_sample(args.chol_sample_grid, 1)
                line["cholesky"]["cpu_baseline"]["note"] = "bounded sample: the %d^3 grid of cholesky.sample_64 (the 100^3 case is ~60 s per CPU factorization)" % args.chol_sample_grid
        except Exception as e:  # the headline line must still be printed
            line["cholesky"] = {"error": repr(e)}
    if world == 1 and args.workload in ("all", "configs"):
        try:
            line["config1"] = bench_config1(args.steps)
        except Exception as e:
            line["config1"] = {"error": repr(e)}
        try:
            line["config2_single"] = bench_config2_single(args.steps)
        except Exception as e:
            line["config2_single"] = {"error": repr(e)}
        line.update(run_ipm_child())
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
