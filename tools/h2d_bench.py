"""Host-to-device copy floor of the batched KLU path: every rank uploads the 961 MB value block of one step (4096 x 29336
doubles) from pinned host memory, all ranks at once, with 1 or 2 copy streams per GPU -- the bound of bench.py's `e2e` at N
GPUs.  torchrun --nproc-per-node N tools/h2d_bench.py   (or plain python for N = 1).  Prints one JSON line on rank 0."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bound = 0
if os.environ.get("H2D_BIND_NUMA", "1") == "1":
    from kvxopt_b200.dist import bind_to_gpu_numa
    bound = bind_to_gpu_numa(local)
nbytes = 4096 * 29336 * 8
host = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
host.fill_(1.0)
dev = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
out = {"bytes_per_copy": nbytes, "n_gpus": world, "cores_bound_per_rank": bound}


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for nstreams in (1, 2):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    part = host.numel() // nstreams
    ms = []
    for rep in range(6):
        barrier()
        t0 = time.perf_counter()
        for q, s in enumerate(streams):
            with torch.cuda.stream(s):
                a, b = q * part, (q + 1) * part if q + 1 < nstreams else host.numel()
                dev[a:b].copy_(host[a:b], non_blocking=True)
        barrier()
        ms.append((time.perf_counter() - t0) * 1e3)
    best = min(ms[1:])
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["streams_%d" % nstreams] = {"ms_max_over_ranks": float(t.item()), "gbs_per_gpu": nbytes / (float(t.item()) * 1e-3) / 1e9,
                                    "gbs_aggregate": world * nbytes / (float(t.item()) * 1e-3) / 1e9}
if rank == 0:
    import subprocess
    try:
        out["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:12]
    except Exception:
        pass
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
