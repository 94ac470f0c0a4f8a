"""BASELINE config 5: synthetic sparse QP through the reference coneqp (probe build oracle/_ref) with the B200 cholmod
module as kvxopt.cholmod.  usage: run_qp.py nx ny nrand [kkt]  (config 5: 500 400 5000; small: 500 400 2000)
With a 4th argument `kkt` the device-side KKT solver (kvxopt_b200.kkt) is plugged in through the kktsolver callable."""
import os, sys, time, functools, faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(int(os.environ.get("QP_DUMP_AFTER", "100000")), exit=True)
print = functools.partial(print, flush=True)
import numpy as np, scipy.sparse as sp
# the reference's own host code (misc_solvers.scale -> BLAS) crashes inside multi-threaded OpenBLAS at m >= 4e5 on this image
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import kvxopt
from kvxopt_b200 import cholmod, klu
mode = sys.argv[4] if len(sys.argv) > 4 else "chol2"
if mode == "cpu":
    # CPU arm: the oracle restatement behind kvxopt.cholmod (kind "port"; SuiteSparse is not installable here)
    from oracle import cholmod_cpu as cholmod
    sys.modules["kvxopt.cholmod"] = cholmod; kvxopt.cholmod = cholmod
    os.environ["OPENBLAS_NUM_THREADS"] = os.environ.get("QP_CPU_THREADS", "16")
else:
    cholmod.install(kvxopt)
    # the CUDA context (1-2 s, once per process) is created before the timed region
    import scipy.sparse as _sp
    cholmod.linsolve(_sp.identity(2, format="csc"), np.ones((2, 1), order="F"))
from kvxopt import matrix, spmatrix, solvers
from generators import qp_instance
nx, ny, nrand = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (100, 80, 200)
P, q, G, h = qp_instance(nx, ny, nrand)
def tosp(M):
    M = sp.coo_matrix(M); return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)
t_num = [0.0]; t_sol = [0.0]; n_num = [0]; n_sol = [0]; t_sym = [0.0]
on, osv, osy = cholmod.numeric, cholmod.solve, cholmod.symbolic
def numeric(*a, **k):
    t = time.perf_counter(); r = on(*a, **k); dt = time.perf_counter() - t; t_num[0] += dt; n_num[0] += 1
    if dt > 0.05: print("  [numeric call %d took %.3f s, n=%d]" % (n_num[0], dt, a[0].size[0]), flush=True)
    return r
def solve(*a, **k):
    t = time.perf_counter(); r = osv(*a, **k); dt = time.perf_counter() - t; t_sol[0] += dt; n_sol[0] += 1
    if dt > 0.05: print("  [solve call %d took %.3f s]" % (n_sol[0], dt), flush=True)
    return r
def symbolic(*a, **k):
    t = time.perf_counter(); r = osy(*a, **k); t_sym[0] += time.perf_counter() - t
    if mode != "cpu" and cholmod.factor_info(r)["n"] > 0: print("symbolic: %s" % {k2: v for k2, v in cholmod.factor_info(r).items() if k2 in ("n", "nnz_L", "flops", "nsuper", "max_front_rows", "ms_analyze")}, flush=True)
    return r
cholmod.numeric, cholmod.solve, cholmod.symbolic = numeric, solve, symbolic
solvers.options["show_progress"] = True
t0 = time.perf_counter()
use_kkt = mode == "kkt"
if use_kkt:
    from kvxopt_b200 import kkt
    Pk, Gk = tosp(sp.tril(P)), tosp(G)
    t0 = time.perf_counter()
    ks = kkt.qp_kktsolver(Pk, Gk)
    sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=ks)
    print("device KKT solver:", ks.info())
elif mode in ("ldl", "ldlreg"):
    # sparse LDL' of the 3 x 3 KKT system (counterpart of the reference's dense 'ldl'); ldlreg: kktreg = 1e-9, free ordering
    from kvxopt_b200 import kkt
    Pk, Gk = tosp(sp.tril(P)), tosp(G)
    t0 = time.perf_counter()
    f3 = kkt.ldl(Gk, {"l": Gk.size[0], "q": [], "s": []}, spmatrix([], [], [], (0, Pk.size[0])), kktreg=1e-9 if mode == "ldlreg" else None)
    sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=lambda W: f3(W, Pk))
    print("sparse LDL' KKT solver:", f3.info())
elif mode == "ldl2":
    # sparse signed LDL' of the reduced 2 x 2 system (counterpart of the reference's dense 'ldl2')
    from kvxopt_b200 import kkt
    Pk, Gk = tosp(sp.tril(P)), tosp(G)
    t0 = time.perf_counter()
    f2 = kkt.ldl2(Gk, {"l": Gk.size[0], "q": [], "s": []}, spmatrix([], [], [], (0, Pk.size[0])))
    sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=lambda W: f2(W, Pk))
    print("sparse LDL' (2 x 2) KKT solver:", f2.info())
else:
    sol = solvers.qp(tosp(sp.tril(P)), matrix(q), tosp(G), matrix(h))
wall = time.perf_counter() - t0
print("mode %s: status %s iterations %d objective %.10f wall %.2f s => %.2f IPM iterations/s%s" % (mode, sol["status"], sol["iterations"], sol["primal objective"], wall, sol["iterations"] / wall, "" if mode == "cpu" else " (CUDA context created before the timed region)"))
print("cholmod: symbolic %.2f s | numeric %d calls %.3f s | solve %d calls %.3f s | everything else (reference Python IPM, host S assembly) %.2f s" % (
    t_sym[0], n_num[0], t_num[0], n_sol[0], t_sol[0], wall - t_sym[0] - t_num[0] - t_sol[0]))
