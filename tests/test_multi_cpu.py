"""world_size-2 gloo test (CPU) of the multi-GPU path of the batched KLU workload: every rank
analyses the same pattern and must arrive at the identical static plan; the matrices are sharded by rank with no
data-path collective; whole-job accounting = sum over ranks, timing = max over ranks."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def host_pattern(A):
    """pivot order of the product's HOST analysis (no GPU involved)"""
    import ctypes as C
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


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import KluOracle
    A = bench.load_activsg()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_multi_cpu import host_pattern
    cp, ri, vx, P, Q, info = host_pattern(A)          # host analysis + pivot search, no GPU
    h = hashlib.sha256(P.tobytes() + Q.tobytes() + np.int64(info["nnz_L"]).tobytes()).digest()
    t = torch.tensor(list(h), dtype=torch.uint8)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same_plan = all(bool((g == gathered[0]).all()) for g in gathered)
    per_rank = 6
    vals = bench.perturbed_values(A.data, per_rank, rank)
    sig = torch.tensor([float(vals.sum())], dtype=torch.float64)
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    O = KluOracle(A.shape[0], cp, ri, vx, P0=P, Q=Q)
    worst = 0.0
    b = np.random.default_rng(rank).standard_normal(A.shape[0])
    import scipy.sparse as sp
    for k in range(per_rank):
        O.refactor(vals[k])
        x = O.solve(b)
        Ak = sp.csc_matrix((vals[k], A.indices, A.indptr), shape=A.shape)
        worst = max(worst, float(np.abs(Ak @ x - b).max()))
    done = torch.tensor([per_rank], dtype=torch.int64)
    dist.all_reduce(done, op=dist.ReduceOp.SUM)              # whole-job units
    tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)              # timing rule: max over ranks
    if rank == 0:
        out.put(dict(same_plan=same_plan, distinct_shards=float(sigs[0]) != float(sigs[1]), total=int(done.item()),
                     tmax=float(tmax.item()), worst=worst))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["same_plan"], "ranks must derive the identical pivot order / plan from the same pattern"
    assert res["distinct_shards"], "every rank owns different matrices"
    assert res["total"] == 12 and res["tmax"] == 2.0
    assert res["worst"] < 1e-7


def test_reference_arm_prints_contract_line():
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "klu"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["klu_factor"]["value"] > 0
    # the reference arm is independent of the product: libb200sparse.so is never mapped into its process
    assert not any("b200sparse" in s for s in line["native_so_loaded"]), line["native_so_loaded"]
