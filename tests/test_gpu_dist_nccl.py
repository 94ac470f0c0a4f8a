"""The real NCCL path of the multi-GPU Cholesky (kvxopt_b200/dist.py: DistCholesky under torchrun, one process per GPU):
runs when the box has at least two GPUs (the single-GPU boxes of the regular test run exercise the same protocol through
VirtualRanks in test_gpu_dist.py).  The launched tool asserts a backward error <= 1e-12 on the gathered factor."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("split", [1, 0])
def test_nccl_subtree_to_subcube_with_shared_schur_complements(split):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU over NCCL); VirtualRanks covers the protocol on one GPU")
    world = 4 if ngpu >= 4 else 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    # 40^3 Laplacian; max_merge_cols = 512 keeps the top separators (1600 and 800 columns) apart so that at 4 ranks the
    # level-1 fronts have Schur complements to share
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + split), os.path.join(ROOT, "tools", "dist_chol_check.py"), "40", "512", str(split), "1e8", "256"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "backward error" in r.stdout
    if split and world >= 4:
        assert "shared fronts: {}" not in r.stdout
