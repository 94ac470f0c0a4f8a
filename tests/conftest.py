import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip: nothing here skips gpu tests.
    pass


def load_matrix(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    n = int(z["n"])
    return sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n))


def sym_from_lower(Al):
    return (Al + sp.tril(Al, -1).T).tocsc()


def lap3d(nx, ny, nz):
    def T1(n):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    A = sp.kron(Iz, sp.kron(Iy, T1(nx))) + sp.kron(Iz, sp.kron(T1(ny), Ix))
    if nz > 1:
        A = A + sp.kron(T1(nz), sp.kron(Iy, Ix))
    return A.tocsc()


def rand_spd(n, dens, seed):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=dens, random_state=rng, format="csc")
    return (M + M.T + sp.identity(n) * (n * dens * 2 + 1)).tocsc()


def lower_ccs(A):
    Al = sp.tril(A).tocsc()
    Al.sort_indices()
    return Al


@pytest.fixture(scope="session")
def kvx():
    """the reference's own matrix types + IPM (probe build oracle/_ref) with the B200 modules plugged in"""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "kvxopt", "__init__.py")):
        pytest.fail("oracle/_ref/kvxopt is missing: run oracle/build_ref.sh in the build container")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import kvxopt
    from kvxopt_b200 import cholmod, klu
    cholmod.install(kvxopt)
    klu.install(kvxopt)
    kvxopt.solvers.options["show_progress"] = False
    return kvxopt
