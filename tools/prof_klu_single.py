"""Latency breakdown of the single-matrix KLU path (klu.symbolic / numeric / solve on ACTIVSg2000); the repeat solve with the
one-matrix kernel (k_klu_solve_one) and with the batched level kernel (B200S_KLU_SOLVE_ONE=0)."""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import klu
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ACTIVSg2000.npz"))
n = int(z["n"]); A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n)); A.sort_indices()
B = np.asfortranarray(np.ones((n, 1)))
klu.linsolve(A, B.copy(order="F"))
for flag in ("1", "0"):
    os.environ["B200S_KLU_SOLVE_ONE"] = flag
    for rep in range(2):
        t0 = time.perf_counter(); Fs = klu.symbolic(A); t1 = time.perf_counter()
        if rep: os.environ["B200S_DEBUG"] = "1"
        Fn = klu.numeric(A, Fs); t2 = time.perf_counter()
        os.environ.pop("B200S_DEBUG", None)
        X = B.copy(order="F"); klu.solve(A, Fs, Fn, X); t3 = time.perf_counter()
        reps = []
        for _ in range(20):
            X = B.copy(order="F"); t4 = time.perf_counter(); klu.solve(A, Fs, Fn, X); reps.append(time.perf_counter() - t4)
        t5 = time.perf_counter(); klu.linsolve(A, B.copy(order="F")); t6 = time.perf_counter()
        print("SOLVE_ONE=%s: symbolic %.2f ms numeric %.2f ms first solve %.2f ms repeat solve %.3f ms (device %.3f ms) | linsolve %.2f ms" % (
            flag, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, np.median(reps) * 1e3, klu.factor_info(Fn)["ms_solve"] if hasattr(klu, "factor_info") else float("nan"), (t6 - t5) * 1e3), flush=True)
