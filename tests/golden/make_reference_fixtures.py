"""Golden vectors produced by the UNMODIFIED reference (probe build oracle/_ref, see oracle/build_ref.sh)
in the build container.  The GPU box has no /root/reference, hence these are committed as fixtures.

  boeing2_lp.npz   LP in matrix form from reference tests/boeing2.mps via the reference's own MPS reader
                   (modeling.py op.fromfile / _inmatrixform('sparse')), plus the iteration count and
                   objective of the reference IPM (solvers.lp -> conelp) with its dense LAPACK KKT solvers.
  posv_<name>.npz  X = A^-1 B for the reference test matrices (lower triangle as stored = a symmetric
                   matrix), B = standard normal n x 3 (numpy default_rng(0)), computed by the reference's
                   lapack.posv on the dense matrix -- the pin for cholmod.linsolve parity.
  qp_mini.npz      small instance of the BASELINE config-5 QP generator solved by the reference coneqp
                   with the dense 'chol' KKT solver: iteration count and objective.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, HERE)

import kvxopt  # noqa: E402
from kvxopt import matrix, spmatrix, solvers, lapack  # noqa: E402
from kvxopt.modeling import op  # noqa: E402

solvers.options["show_progress"] = False


class _Stub:   # kvxopt.misc imports kvxopt.cholmod at module level (misc.py:21); the dense solvers never call it
    options = {}


sys.modules["kvxopt.cholmod"] = _Stub()
kvxopt.cholmod = _Stub()


def ccs_arrays(S):
    cp, ri, vx = S.CCS
    return (np.array(cp, dtype=np.int64).reshape(-1), np.array(ri, dtype=np.int64).reshape(-1),
            np.array(vx, dtype=np.float64).reshape(-1))


def boeing2():
    import io
    import contextlib
    p = op()
    with contextlib.redirect_stdout(io.StringIO()):
        p.fromfile("/root/reference/tests/boeing2.mps")
        t = p._inmatrixform("sparse")
    lp1 = t[0]
    x = lp1.variables()[0]
    c = matrix(lp1.objective._linear._coeff[x], tc="d")
    G = lp1._inequalities[0]._f._linear._coeff[x]
    h = -lp1._inequalities[0]._f._constant
    A = lp1._equalities[0]._f._linear._coeff[x]
    b = -lp1._equalities[0]._f._constant
    out = {}
    for name, kkt in (("chol", "chol"), ("chol2", "chol2"), ("ldl", "ldl")):
        # dense G, A  => the reference's dense LAPACK KKT solvers
        sol = solvers.lp(c[:], matrix(G), h, matrix(A), b, kktsolver=kkt)
        out["iters_" + name] = sol["iterations"]
        out["pobj_" + name] = sol["primal objective"]
        out["status_" + name] = sol["status"]
        if name == "chol2":
            out["x_ref"] = np.array(sol["x"]).reshape(-1)
    Gp, Gi, Gx = ccs_arrays(G)
    Ap, Ai, Ax = ccs_arrays(A)
    np.savez_compressed(os.path.join(HERE, "boeing2_lp.npz"), c=np.array(c).reshape(-1), h=np.array(h).reshape(-1),
                        b=np.array(b).reshape(-1), G_size=np.array(G.size), Gp=Gp, Gi=Gi, Gx=Gx, A_size=np.array(A.size),
                        Ap=Ap, Ai=Ai, Ax=Ax, **out)
    print("boeing2", G.size, A.size, {k: v for k, v in out.items() if k != "x_ref"})


def posv(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    n = int(z["n"])
    Al = sp.csc_matrix((z["values"], z["rowind"], z["colptr"]), shape=(n, n))
    Ad = (Al + sp.tril(Al, -1).T).toarray()
    B = np.random.default_rng(0).standard_normal((n, 3))
    Am = matrix(Ad)
    X = matrix(B)
    lapack.posv(Am, X)          # reference LAPACK path (dpotrf + dpotrs), lower triangle
    X = np.array(X)
    r = np.linalg.norm(Ad @ X - B) / (np.linalg.norm(Ad) * np.linalg.norm(X))
    np.savez_compressed(os.path.join(HERE, "posv_" + name + ".npz"), X=X)
    print("posv", name, n, "relative residual of reference solution %.2e" % r)


from generators import qp_instance  # noqa: E402


def to_spmatrix(M):
    M = M.tocoo()
    return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)


def qp_mini():
    P, q, G, h = qp_instance(50, 40, 50)
    sol = solvers.qp(matrix(P.toarray()), matrix(q), matrix(G.toarray()), matrix(h), kktsolver="chol")
    np.savez_compressed(os.path.join(HERE, "qp_mini.npz"), nx=50, ny=40, nrand=50, iters=sol["iterations"],
                        pobj=sol["primal objective"], x=np.array(sol["x"]).reshape(-1))
    print("qp_mini", sol["status"], sol["iterations"], sol["primal objective"])


if __name__ == "__main__":
    boeing2()
    for nm in ("bcsstk13", "bcsstk24"):
        posv(nm)
    qp_mini()
