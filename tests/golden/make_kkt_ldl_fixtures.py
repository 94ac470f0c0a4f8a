"""Golden vectors of the UNMODIFIED reference's dense 3 x 3 KKT solver `misc.kkt_ldl` (src/python/misc.py:1055-1130,
LAPACK sytrf/sytrs with Bunch-Kaufman pivoting), run from the probe build oracle/_ref in the build container:

  kkt_ldl_ref.npz   boeing2 LP data (G 352 x 143, A 4 x 143, H = None) and the mini config-5 QP (H = P, no equalities),
                    each with a seeded scaling W['d'] = exp(N(0,1)) and right-hand sides N(0,1); stored are the inputs
                    and the reference's (ux, uy, W uz).

They pin oracle/ldl_oracle.py (tests/test_oracle.py) and the sparse no-pivoting LDL' of the B200 engine
(tests/test_gpu_ldl.py) to the reference's own answers for the quasi-definite KKT systems the 'ldl' solver factors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402  (sets up oracle/_ref on sys.path and the cholmod stub)
from kvxopt import matrix, misc  # noqa: E402
from generators import qp_instance  # noqa: E402
import scipy.sparse as sp  # noqa: E402


def run(G, A, H, seed):
    m, n = G.shape
    p = A.shape[0]
    rng = np.random.default_rng(seed)
    d = np.exp(rng.standard_normal(m))
    bx, by, bz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    W = {"d": matrix(d), "di": matrix(1.0 / d), "dnl": matrix(0.0, (0, 1)), "dnli": matrix(0.0, (0, 1)),
         "r": [], "rti": [], "beta": [], "v": []}
    dims = {"l": m, "q": [], "s": []}
    factor = misc.kkt_ldl(matrix(G), dims, matrix(A) if p else matrix(0.0, (0, n)))
    solve = factor(W, matrix(H) if H is not None else None)
    x, y, z = matrix(bx), matrix(by) if p else matrix(0.0, (0, 1)), matrix(bz)
    solve(x, y, z)
    return dict(d=d, bx=bx, by=by, bz=bz, ux=np.array(x).reshape(-1), uy=np.array(y).reshape(-1), uz=np.array(z).reshape(-1))


if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "boeing2_lp.npz"))
    G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"])).toarray()
    A = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"])).toarray()
    out = {"lp_" + k: v for k, v in run(G, A, None, 11).items()}
    P, q, Gq, h = qp_instance(50, 40, 50)
    out.update({"qp_" + k: v for k, v in run(Gq.toarray(), np.zeros((0, P.shape[0])), P.toarray(), 12).items()})
    np.savez_compressed(os.path.join(HERE, "kkt_ldl_ref.npz"), **out)
    print({k: v.shape for k, v in out.items()})
