"""torchrun --nproc-per-node N tools/dist_chol_check.py [nx [max_merge_cols [split]]]: NCCL subtree-to-subcube factorization of
the nx^3 Laplacian (shared Schur complements of the top separators unless split = 0), checked against the residual; prints
timing per rank count."""
import ctypes as C, os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod, dist as D
import scipy.sparse as sp

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); L.fn["b200s_set_device"](local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
cap = int(sys.argv[2]) if len(sys.argv) > 2 else D.DIST_MAX_MERGE_COLS
split = (int(sys.argv[3]) != 0) if len(sys.argv) > 3 else True
cholmod.engine_options["max_merge_cols"] = cap
F = cholmod.symbolic(Al, p=perm)
dc = D.DistCholesky(F, world, rank, split=split,
                    split_args=dict(min_flops=float(sys.argv[4]), min_rows=int(sys.argv[5])) if len(sys.argv) > 5 else None)
vals = torch.from_numpy(Al.data.copy()).cuda()
torch.cuda.synchronize(); dist.barrier()
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    st, minor = dc.factorize(vals.data_ptr(), True)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    dc.gather_factor(0)
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    if rank == 0:
        print("world %d nx %d rep %d: factor %.1f ms  gather %.1f ms  status %d" % (world, nx, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st), flush=True)
# distributed triangular solves straight after the factorization (the panels stay where they were factored)
bvec = torch.from_numpy(np.random.default_rng(0).standard_normal(n)).cuda()
ts = []
for rep in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    xdist = dc.solve(bvec)
    torch.cuda.synchronize(); dist.barrier(); ts.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    A_ = (Al + sp.tril(Al, -1).T).tocsr()
    xd = xdist.cpu().numpy(); bh = bvec.cpu().numpy()
    berr_d = np.linalg.norm(A_ @ xd - bh) / (12 * np.linalg.norm(xd) + np.linalg.norm(bh))
    print("distributed solve: %.1f ms (first %.1f), backward error %.2e" % (min(ts[1:]), ts[0], berr_d), flush=True)
    assert berr_d < 1e-12
if os.environ.get("B200S_DIST_TRACE"):
    for r in range(world):
        dist.barrier()
        if r == rank and rank in (0, 1, 4):
            print("rank %d trace: %s" % (rank, " | ".join("L%d %s %.1f" % t for t in dc.trace)), flush=True)
if rank == 0:
    info = cholmod.factor_info(F)
    b = np.random.default_rng(0).standard_normal((n, 1)); x = np.asfortranarray(b.copy())
    cholmod.solve(F, x)
    A = (Al + sp.tril(Al, -1).T).tocsr()
    berr = np.linalg.norm(A @ x - b) / (12 * np.linalg.norm(x) + np.linalg.norm(b))
    shares = D.front_work(dc.lay); tot = shares.sum()
    print("backward error %.2e | flops %.3g | work share per rank: %s | transfers: %d update matrices, %.1f MB" % (
        berr, info["flops"], [round(float(D._work_share(dc.lay, dc.owner, dc.splan, r) / tot), 3) for r in range(world)],
        sum(len(l) for l in dc.xplan), sum(dc.lay["usize"][m[0]] for l in dc.xplan for m in l) * 8 / 1e6), flush=True)
    print("max_merge_cols %d, shared fronts: %s" % (cap, {int(k): [(int(r), int(lo), int(hi)) for r, lo, hi in v] for k, v in dc.splan.items()}), flush=True)
    top = sorted(range(len(shares)), key=lambda q: -shares[q])[:8]
    print("largest fronts (id, nc, nr, owner, level, share):", [(int(q), int(dc.lay["nc"][q]), int(dc.lay["nr"][q]), int(dc.owner[q]), int(dc.lay["level"][q]), round(float(shares[q] / tot), 3)) for q in top], flush=True)
    assert berr < 1e-12
dist.barrier(); dist.destroy_process_group()
