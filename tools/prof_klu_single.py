"""Latency breakdown of the single-matrix KLU path (klu.symbolic / numeric / solve on ACTIVSg2000)."""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import klu
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ACTIVSg2000.npz"))
n = int(z["n"]); A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n)); A.sort_indices()
B = np.asfortranarray(np.ones((n, 1)))
klu.linsolve(A, B.copy(order="F"))
for rep in range(2):
    t0 = time.perf_counter(); Fs = klu.symbolic(A); t1 = time.perf_counter()
    if rep: os.environ["B200S_DEBUG"] = "1"
    Fn = klu.numeric(A, Fs); t2 = time.perf_counter()
    os.environ.pop("B200S_DEBUG", None)
    X = B.copy(order="F"); klu.solve(A, Fs, Fn, X); t3 = time.perf_counter()
    print("symbolic %.2f ms numeric %.2f ms solve %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
