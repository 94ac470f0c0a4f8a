"""cProfile of BASELINE config 3 (boeing2 LP through the reference conelp with kvxopt_b200.kkt.chol): where the host time goes."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np, scipy.sparse as sp
import kvxopt
from kvxopt_b200 import cholmod as gcholmod, klu as gklu, kkt as gkkt
gcholmod.install(kvxopt); gklu.install(kvxopt)
from kvxopt import matrix, solvers
solvers.options["show_progress"] = False
gcholmod.linsolve(sp.identity(2, format="csc"), np.ones((2, 1), order="F"))
z = np.load(os.path.join(ROOT, "tests", "golden", "boeing2_lp.npz"))
G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
A = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
Gd, Ad = matrix(G.toarray()), matrix(A.toarray())
dims = {"l": G.shape[0], "q": [], "s": []}
for name, mk in (("kkt.chol", lambda: gkkt.chol(Gd, dims, Ad)), ("reference chol", lambda: "chol")):
    for _ in range(2):
        t = time.perf_counter(); sol = solvers.conelp(c, Gd, h, dims, Ad, b, kktsolver=mk()); dt = (time.perf_counter() - t) * 1e3
    print(name, "%.2f ms" % dt, sol["iterations"], flush=True)
pr = cProfile.Profile(); k = gkkt.chol(Gd, dims, Ad)
pr.enable(); solvers.conelp(c, Gd, h, dims, Ad, b, kktsolver=k); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
