import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
from kvxopt_b200 import cholmod
for n in (143, 400, 1500):
    rng = np.random.default_rng(1)
    M = rng.standard_normal((n, n)) / np.sqrt(n)
    A = sp.csc_matrix(np.tril(M @ M.T + 2 * np.eye(n)))
    F = cholmod.symbolic(A, p=np.arange(n, dtype=np.int64)); cholmod.numeric(A, F)
    b = rng.standard_normal((n, 1))
    for mode in (0, 3):
        cholmod.set_solve_sweeps(F, mode)
        x = np.asfortranarray(b.copy()); cholmod.solve(F, x)
        t0 = time.perf_counter()
        for _ in range(300):
            x = np.asfortranarray(b.copy()); cholmod.solve(F, x)
        dt = (time.perf_counter() - t0) / 300 * 1e3
        print("n %d mode %d: %.4f ms per solve (host clock), ms_solve %.4f" % (n, mode, dt, cholmod.factor_info(F)["ms_solve"]), flush=True)
