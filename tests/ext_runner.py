"""Runs in a fresh interpreter (tests/test_gpu_ext_modules.py): exercises the COMPILED kvxopt.cholmod / kvxopt.klu extension
modules -- the reference's own src/C/cholmod.c and src/C/klu.c built against include/suitesparse_shim and linked to
libb200sparse.so by tools/build_kvxopt_ext.sh -- with nothing monkey-patched into kvxopt.  Prints one JSON line."""
import importlib.util
import json
import math
import os
import sys
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np

out = {}
import kvxopt
from kvxopt import cholmod, klu, matrix, spmatrix, solvers, log, mul, div
solvers.options["show_progress"] = False
out["cholmod_file"] = os.path.basename(cholmod.__file__)
out["klu_file"] = os.path.basename(klu.__file__)
out["compiled"] = cholmod.__file__.endswith(".so") and klu.__file__.endswith(".so")

# doc/source/spsolvers.rst:555-563  cholmod.linsolve known answer
A = spmatrix([10, 3, 5, -2, 5, 2], [0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3])
X = matrix(range(8), (4, 2), "d")
cholmod.linsolve(A, X)
ref = np.array([[-0.146341463414634, 0.048780487804878], [1.333333333333333, 4.0], [0.487804878048781, 1.170731707317073],
                [2.833333333333333, 7.5]])
out["linsolve_err"] = float(np.abs(np.array(X) - ref).max())
# :580-585 splinsolve gives the inverse
Xs = cholmod.splinsolve(A, spmatrix(1.0, range(4), range(4)))
Afull = np.array(matrix(A)); Afull = np.tril(Afull) + np.tril(Afull, -1).T
out["splinsolve_err"] = float(np.abs(np.array(matrix(Xs)) @ Afull - np.eye(4)).max())
# :759-772 log det through diag (supernodal) and through sys=6 (LDL')
F = cholmod.symbolic(A)
cholmod.numeric(A, F)
out["logdet_diag"] = float(2.0 * sum(log(cholmod.diag(F))))
Lf = cholmod.getfactor(F)
out["getfactor_nnz"] = len(Lf)
B = matrix(range(8), (4, 2), "d")
cholmod.solve(F, B)
out["solve_err"] = float(np.abs(np.array(B) - ref).max())
Bs = cholmod.spsolve(F, spmatrix([1.0, 2.0], [0, 3], [0, 1], (4, 2)))
out["spsolve_err"] = float(np.abs(Afull @ np.array(matrix(Bs)) - np.array(matrix(spmatrix([1.0, 2.0], [0, 3], [0, 1], (4, 2))))).max())
cholmod.options["supernodal"] = 0
F0 = cholmod.symbolic(A)
cholmod.numeric(A, F0)
Di = matrix(1.0, (4, 1))
cholmod.solve(F0, Di, sys=6)
out["logdet_ldl"] = float(-sum(log(Di)))
try:
    cholmod.diag(F0)
    out["diag_ldl_refused"] = False
except ValueError:
    out["diag_ldl_refused"] = True
del cholmod.options["supernodal"]
# not positive definite: numeric raises ArithmeticError(k) (documented contract), the factor recovers afterwards
An = spmatrix([1.0, 2.0, 1.0, -5.0], [0, 1, 1, 2], [0, 0, 1, 2], (3, 3))
Fn = cholmod.symbolic(An)
try:
    cholmod.numeric(An, Fn)
    out["npd"] = "no error"
except ArithmeticError as e:
    out["npd"] = "ArithmeticError(%s)" % (e.args[0],)
try:
    cholmod.solve(Fn, matrix(1.0, (3, 1)))
    out["npd_solve"] = "no error"
except ArithmeticError as e:
    out["npd_solve"] = str(e)
# error contract of the wrapper itself
try:
    cholmod.solve(cholmod.symbolic(A), matrix(1.0, (4, 1)))
    out["symbolic_solve"] = "no error"
except ValueError as e:
    out["symbolic_solve"] = str(e)

# klu: doc/source/spsolvers.rst:333-345 and :420-439, det = 114 (tests/test_sparse_solvers.py:298-313)
V = [2, 3, 3, -1, 4, 4, -3, 1, 2, 2, 6, 1]
I = [0, 1, 0, 2, 4, 1, 2, 3, 4, 2, 1, 4]
J = [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4]
Ak = spmatrix(V, I, J)
Bk = matrix(range(5), tc="d")
klu.linsolve(Ak, Bk)
out["klu_linsolve_err"] = float(np.abs(np.array(Bk).ravel() - [0.052631578947368, -0.035087719298246, 3.0, 5.482456140350877,
                                                               -1.859649122807017]).max())
Fs = klu.symbolic(Ak)
Fk = klu.numeric(Ak, Fs)
out["klu_det"] = float(klu.get_det(Ak, Fs, Fk))
Lm, Um, P, Q, R, Fm, r = klu.get_numeric(Ak, Fs, Fk)
out["klu_identity"] = float(max(abs(R * P * Ak * Q - (Lm * Um + Fm))))
Bt = matrix(1.0, (5, 1))
klu.solve(Ak, Fs, Fk, Bt, trans="T")
out["klu_tsolve_res"] = float(max(abs(Ak.T * Bt - matrix(1.0, (5, 1)))))
try:
    klu.numeric(spmatrix([1.0, 1.0, 1.0, 1.0], [0, 1, 0, 1], [0, 0, 1, 1]), klu.symbolic(spmatrix([1.0, 1.0, 1.0, 1.0], [0, 1, 0, 1], [0, 0, 1, 1])))
    out["klu_singular"] = "no error"
except ArithmeticError as e:
    out["klu_singular"] = str(e)


# the reference's own test-suite, verbatim, against the compiled modules
def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(ROOT, "oracle", "_ref", "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(cls, method):
    res = unittest.TestResult()
    unittest.TestSuite([cls(method)]).run(res)
    return {"ok": res.testsRun == 1 and not res.failures and not res.errors and not res.skipped,
            "detail": [t[1][-300:] for t in res.failures + res.errors]}


ts = load("test_sparse_solvers")
# complex Hermitian matrix through the compiled cholmod wrapper ('z' spmatrix, cholmod.c:144,153,343-357)
Az = spmatrix([10, 3 + 1j, 5, -2 - 2j, 5, 2], [0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3], (4, 4), "z")
Azd = np.array(matrix(Az)); Azd = np.tril(Azd) + np.tril(Azd, -1).conj().T
bz = matrix([1 + 1j, 2, 3 - 2j, 4j])
xz = +bz
cholmod.linsolve(Az, xz)
out["cholmod_z_linsolve_err"] = float(np.abs(np.array(xz).ravel() - np.linalg.solve(Azd, np.array(bz).ravel())).max())
Fz = cholmod.symbolic(Az); cholmod.numeric(Az, Fz)
dz = cholmod.diag(Fz)
Lz = np.array(matrix(cholmod.getfactor(Fz)))
out["cholmod_z_diag_err"] = float(np.abs(np.array(dz).ravel() - np.diag(Lz)).max())
out["cholmod_z_typecode"] = dz.typecode
# the reference's own loops, real and complex ('z') cases alike
out["ref_tests"] = {m: run(ts.TestKLU, m) for m in ("test_lu", "test_linsolve", "test_solve", "test_get_det")}
te = load("test_examples")
for m in ("test_ch9_acent", "test_ch8_lp", "test_ch8_coneqp"):
    out["ref_tests"][m] = run(te.TestExamples, m)

# boeing2 through the reference's default sparse path: misc.kkt_chol2 -> the compiled kvxopt.cholmod
import scipy.sparse as sp
z = np.load(os.path.join(ROOT, "tests", "golden", "boeing2_lp.npz"))
G = sp.coo_matrix(sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"])))
Aeq = sp.coo_matrix(sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"])))
sol = solvers.lp(matrix(z["c"]), spmatrix(G.data.tolist(), G.row.tolist(), G.col.tolist(), G.shape), matrix(z["h"]),
                 spmatrix(Aeq.data.tolist(), Aeq.row.tolist(), Aeq.col.tolist(), Aeq.shape), matrix(z["b"]))
out["boeing2"] = {"status": sol["status"], "iterations": sol["iterations"], "objective": sol["primal objective"]}
out["native"] = sorted({os.path.basename(l.split()[-1]) for l in open("/proc/self/maps") if "b200sparse" in l or "/kvxopt/cholmod" in l or "/kvxopt/klu" in l})
print("EXTRUNNER " + json.dumps(out))
