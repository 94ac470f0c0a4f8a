"""Multi-GPU subtree-to-subcube protocol (kvxopt_b200/dist.py).  On a single GPU the ranks are emulated as several
factor handles on the same device (device-to-device copies instead of NCCL); with >= 2 GPUs the real NCCL path is
exercised by tools/dist_chol_check.py under torchrun (run by hand / by bench.py --gpus N)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import lap3d, lower_ccs, rand_spd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", ["lap3d", "rand"])
def test_virtual_ranks_match_single_gpu_bitwise(world, case, split):
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
    # split: the Schur complements of the top separators are shared by the ranks of their subtree group (thresholds lowered
    # so that these small matrices have shared fronts at all)
    vr = D.VirtualRanks(caps, split=split, split_args=dict(min_flops=1e5, min_rows=100))
    assert len(set(vr.owner.tolist())) == world
    if split and case == "lap3d" and world >= 4:
        assert len(vr.splan) >= 1 and any(len(p) >= 2 for p in vr.splan.values())
    minor = vr.factorize(Al.data)
    assert minor == n
    Xd = np.asfortranarray(B.copy()); cholmod.solve(caps[0], Xd)
    if not vr.splan:
        # same kernels, same summation order per front => identical bits
        assert np.array_equal(Xs, Xd)
    else:
        # a helper's slab is E + (0 - L21 L21') instead of E - L21 L21' accumulated in one pass: one more rounding
        assert np.linalg.norm(Xs - Xd) <= 1e-12 * np.linalg.norm(Xs)
        minor2 = vr.factorize(Al.data)                     # a second factorization reuses (re-zeroed) scratch buffers
        Xd2 = np.asfortranarray(B.copy()); cholmod.solve(caps[0], Xd2)
        assert minor2 == n and np.array_equal(Xd, Xd2)
    berr = (np.linalg.norm(A @ Xd - B, axis=0) / (spla.norm(A, 1) * np.linalg.norm(Xd, axis=0) + np.linalg.norm(B, axis=0))).max()
    assert berr <= 1e-12
    # distributed triangular solves: every rank sweeps over its own fronts, update vectors go up the cut edges, solution
    # entries of the top fronts go down their subtree groups; same kernels and summation order as the single-GPU solve
    for col in range(2):
        xv = vr.solve(B[:, col])
        if not vr.splan:
            assert np.array_equal(xv, Xs[:, col])
        else:
            assert np.linalg.norm(xv - Xs[:, col]) <= 1e-12 * np.linalg.norm(Xs[:, col])
    assert any(vr.sfwd) and any(vr.sbwd) and vr.sgather
    # not positive definite: the smallest failing column over all ranks is reported
    Abad = A.tolil(); Abad[n // 2, n // 2] = -1.0
    Albad = lower_ccs(Abad.tocsc())
    with pytest.raises(ArithmeticError) as e:
        cholmod.numeric(Albad, Fs)
    assert vr.factorize(Albad.data) == e.value.args[0]
