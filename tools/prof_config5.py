"""cProfile of BASELINE config 5 (200k-variable QP through the reference coneqp with kvxopt_b200.kkt.qp_kktsolver): host time
per call site -- how much of an IPM iteration is the KKT plug-in and how much the reference's own Python."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, scipy.sparse as sp
import kvxopt
from kvxopt_b200 import cholmod as gcholmod, klu as gklu, kkt as gkkt
gcholmod.install(kvxopt); gklu.install(kvxopt)
from kvxopt import matrix, spmatrix, solvers
from generators import qp_instance
solvers.options["show_progress"] = False
gcholmod.linsolve(sp.identity(2, format="csc"), np.ones((2, 1), order="F"))
P, q, G, h = qp_instance(500, 400, 5000)
def tosp(M):
    M = sp.coo_matrix(M); return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)
Pk, Gk = tosp(sp.tril(P)), tosp(G)
ks = gkkt.qp_kktsolver(Pk, Gk)
qm, hm = matrix(q), matrix(h)
pr = cProfile.Profile()
t = time.perf_counter(); pr.enable(); sol = solvers.qp(Pk, qm, Gk, hm, kktsolver=ks); pr.disable(); dt = time.perf_counter() - t
print("wall %.3f s, %d iterations, %.2f it/s, objective %.10f" % (dt, sol["iterations"], sol["iterations"] / dt, sol["primal objective"]))
print(ks.info())
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
