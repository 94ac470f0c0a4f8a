"""BASELINE configs 1 and 3 (the small, latency-bound cases of SURVEY section 8d) on the B200 engine, with the CPU arms
timed beside them on this box's host cores: the oracle restatement behind the same interface (kind "port") and, for
config 3, the reference's own dense LAPACK KKT solvers ('chol', 'chol2' with dense G) from the probe build oracle/_ref.
Prints one JSON object.  usage: python tools/run_configs.py [out.json]"""
import json, os, sys, time
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_matrix, sym_from_lower, lower_ccs, GOLD
import kvxopt
from kvxopt_b200 import cholmod as gcholmod, klu as gklu, _lib as L
from oracle import CholOracle
from oracle import cholmod_cpu
from kvxopt import matrix, spmatrix, solvers
solvers.options["show_progress"] = False
out = {"host_cores": os.cpu_count()}


def best(f, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = f(); ts.append((time.perf_counter() - t) * 1e3)
    return min(ts), r


# ---- config 1: cholmod.linsolve on bcsstk24 (lower triangle as stored), random RHS
Al = lower_ccs(load_matrix("bcsstk24")); A = sym_from_lower(Al); n = A.shape[0]
rng = np.random.default_rng(0)
for nrhs in (1, 3):
    B = rng.standard_normal((n, nrhs))
    gcholmod.linsolve(Al, np.asfortranarray(B.copy()))                                   # context + first-call setup
    ms_lin, _ = best(lambda: gcholmod.linsolve(Al, np.asfortranarray(B.copy())))
    ms_sym, F = best(lambda: gcholmod.symbolic(Al))
    ms_num, _ = best(lambda: gcholmod.numeric(Al, F))
    X = np.asfortranarray(B.copy())
    ms_sol, _ = best(lambda: gcholmod.solve(F, X))
    X = np.asfortranarray(B.copy()); gcholmod.solve(F, X)
    res = float(np.linalg.norm(A @ X - B) / (abs(A).sum(axis=0).max() * np.linalg.norm(X) + np.linalg.norm(B)))
    info = gcholmod.factor_info(F)
    perm = gcholmod.factor_perm(F)
    O = CholOracle(n, Al.indptr, Al.indices, "L", perm)
    ms_onum, _ = best(lambda: O.factorize(Al.data), 3)
    ms_osol, Xo = best(lambda: O.solve(B), 3)
    out["config1_bcsstk24_nrhs%d" % nrhs] = {
        "n": n, "nnz_L": info["nnz_L"], "flops": info["flops"],
        "gpu_ms": {"linsolve_total": ms_lin, "symbolic_host": ms_sym, "numeric": ms_num, "solve": ms_sol},
        "cpu_port_ms": {"numeric": ms_onum, "solve": ms_osol, "kind": "port (oracle/chol_oracle.c, OpenBLAS, same ordering)"},
        "backward_error": res, "rel_diff_vs_oracle": float(np.linalg.norm(X - Xo) / np.linalg.norm(Xo))}

# ---- config 2, single-matrix part: klu.linsolve / symbolic / numeric / solve on ACTIVSg2000 (the batched part is bench.py)
from oracle import KluOracle
Aj = load_matrix("ACTIVSg2000"); Aj.sort_indices(); nj = Aj.shape[0]
Bj = np.asfortranarray(np.random.default_rng(0).standard_normal((nj, 1)))
gklu.linsolve(Aj, Bj.copy(order="F"))
ms_lin, _ = best(lambda: gklu.linsolve(Aj, Bj.copy(order="F")))
ms_sym, Fs = best(lambda: gklu.symbolic(Aj))
ms_num, Fn = best(lambda: gklu.numeric(Aj, Fs))
Xj = Bj.copy(order="F")
ms_sol, _ = best(lambda: gklu.solve(Aj, Fs, Fn, Xj))
Xj = Bj.copy(order="F"); gklu.solve(Aj, Fs, Fn, Xj)
resj = float(np.abs(Aj @ Xj - Bj).max())
_Lm, _Um, _P, _Q, _R, _Fm, _r = gklu.get_numeric(Aj, Fs, Fn)
_P0, _Qv = np.asarray(_P.tocsr().indices), np.asarray(_Q.tocsc().indices)
ms_ofac, Oj = best(lambda: KluOracle(nj, Aj.indptr, Aj.indices, Aj.data, P0=_P0, Q=_Qv), 3)
ms_osol, _ = best(lambda: Oj.solve(Bj[:, 0]), 3)
out["config2_ACTIVSg2000_single"] = {
    "n": nj, "nnz": int(Aj.nnz),
    "gpu_ms": {"linsolve_total": ms_lin, "symbolic_host": ms_sym, "numeric (host pivot search + plan + upload + device refactor)": ms_num, "solve": ms_sol},
    "cpu_port_ms": {"factor (pivoting Gilbert-Peierls, same ordering)": ms_ofac, "solve": ms_osol, "kind": "port (oracle/klu_oracle.c: BTF + AMD + Gilbert-Peierls)"},
    "max_residual": resj}

# ---- config 3: solvers.lp on boeing2
z = np.load(GOLD + "/boeing2_lp.npz")
G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
Aeq = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])


def tosp(M):
    M = sp.coo_matrix(M); return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)


def run_lp(Gm, Am, kkt=None):
    kw = {} if kkt is None else {"kktsolver": kkt}
    t = time.perf_counter(); sol = solvers.lp(c, Gm, h, Am, b, **kw); dt = (time.perf_counter() - t) * 1e3
    return dt, sol


arms = {}
gcholmod.install(kvxopt)
run_lp(tosp(G), tosp(Aeq))
dt = min(run_lp(tosp(G), tosp(Aeq))[0] for _ in range(3)); _, sol = run_lp(tosp(G), tosp(Aeq))
arms["gpu_chol2_sparse_cuda_cholmod"] = {"ms": dt, "iterations": sol["iterations"], "pobj": sol["primal objective"]}
from kvxopt_b200 import kkt as gkkt
Gk, Ak = tosp(G), tosp(Aeq)
run_lp(Gk, Ak, gkkt.lp_kktsolver(Gk, Ak))
dt = min(run_lp(Gk, Ak, gkkt.lp_kktsolver(Gk, Ak))[0] for _ in range(3)); _, sol = run_lp(Gk, Ak, gkkt.lp_kktsolver(Gk, Ak))
arms["gpu_device_kkt_solver"] = {"ms": dt, "iterations": sol["iterations"], "pobj": sol["primal objective"],
                                 "api": "solvers.lp(..., kktsolver=kvxopt_b200.kkt.lp_kktsolver(G, A))"}
sys.modules["kvxopt.cholmod"] = cholmod_cpu; kvxopt.cholmod = cholmod_cpu
import kvxopt.misc as misc
misc.cholmod = cholmod_cpu
dt = min(run_lp(tosp(G), tosp(Aeq))[0] for _ in range(3)); _, sol = run_lp(tosp(G), tosp(Aeq))
arms["cpu_chol2_sparse_oracle_cholmod"] = {"ms": dt, "iterations": sol["iterations"], "pobj": sol["primal objective"], "kind": "port"}
Gd, Ad = matrix(G.toarray()), matrix(Aeq.toarray())
for name in ("chol", "chol2"):
    dt = min(run_lp(Gd, Ad, name)[0] for _ in range(3)); _, sol = run_lp(Gd, Ad, name)
    arms["cpu_reference_dense_" + name] = {"ms": dt, "iterations": sol["iterations"], "pobj": sol["primal objective"], "kind": "reference (LAPACK)"}
out["config3_boeing2_lp"] = arms
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s)
