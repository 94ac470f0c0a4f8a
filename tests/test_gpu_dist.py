"""Multi-GPU subtree-to-subcube protocol (kvxopt_b200/dist.py).  On a single GPU the ranks are emulated as several
factor handles on the same device (device-to-device copies instead of NCCL); with >= 2 GPUs the real NCCL path is
exercised by tools/dist_chol_check.py under torchrun (run by hand / by bench.py --gpus N)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import lap3d, lower_ccs, rand_spd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", ["lap3d", "rand"])
def test_virtual_ranks_match_single_gpu_bitwise(world, case):
    from kvxopt_b200 import _lib as L, cholmod, dist as D
    assert L.device_count() > 0
    if case == "lap3d":
        nx = 22
        A = lap3d(nx, nx, nx)
        perm = np.zeros(A.shape[0], np.int64)
        assert L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm)) == 0
    else:
        A = rand_spd(3000, 0.002, 5)
        perm = None
    Al = lower_ccs(A)
    n = A.shape[0]
    B = np.random.default_rng(0).standard_normal((n, 2))
    Fs = cholmod.symbolic(Al, p=perm)
    cholmod.numeric(Al, Fs)
    Xs = np.asfortranarray(B.copy()); cholmod.solve(Fs, Xs)
    caps = [cholmod.symbolic(Al, p=perm) for _ in range(world)]
    vr = D.VirtualRanks(caps)
    assert len(set(vr.owner.tolist())) == world
    minor = vr.factorize(Al.data)
    assert minor == n
    Xd = np.asfortranarray(B.copy()); cholmod.solve(caps[0], Xd)
    # same kernels, same summation order per front => identical bits
    assert np.array_equal(Xs, Xd)
    berr = (np.linalg.norm(A @ Xd - B, axis=0) / (spla.norm(A, 1) * np.linalg.norm(Xd, axis=0) + np.linalg.norm(B, axis=0))).max()
    assert berr <= 1e-12
    # not positive definite: the smallest failing column over all ranks is reported
    Abad = A.tolil(); Abad[n // 2, n // 2] = -1.0
    Albad = lower_ccs(Abad.tocsc())
    with pytest.raises(ArithmeticError) as e:
        cholmod.numeric(Albad, Fs)
    assert vr.factorize(Albad.data) == e.value.args[0]
